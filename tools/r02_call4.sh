#!/bin/bash
# round 2, GPU call 4: GPU test suite after the fixes, process-exit check, bench N=1 with the extras, reference arm (short)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
tail -5 gpurun_out/r02_bench_n1.err
cat gpurun_out/r02_bench_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"
cat gpurun_out/r02_bench_ref.json
