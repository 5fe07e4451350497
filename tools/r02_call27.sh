#!/bin/bash
# round 2, GPU call 27: SBGEMMT (triangle walk in the tcgen05 kernels) -- its parity tests, the SBGEMM regression tests
# (the kernels gained an argument), smoke(), and SBGEMMT against SBGEMM timings at 4096 / 8192 / 16384
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sbgemmt_gpu.py tests/test_f_rows_gpu.py tests/test_round2_gpu.py tests/test_gemm_gpu.py tests/test_full_size_gpu.py -m gpu -x -q -k "sbgemm or sbgemmt or bf16 or sb or golden" > gpurun_out/r02_sbgemmt_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r02_sbgemmt_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"
tail -4 gpurun_out/r02_smoke.log
timeout 600 python tools/sbgemmt_time.py > gpurun_out/r02_sbgemmt_times.jsonl 2> gpurun_out/r02_sbgemmt_times.err; echo "times rc=$?"
cat gpurun_out/r02_sbgemmt_times.jsonl; tail -3 gpurun_out/r02_sbgemmt_times.err
