#!/bin/bash
# round 2, GPU call 37: GEMM3M on three real products -- its tests (the ctest-grid sweep stays on the 4-multiply kernel), then timing against ?GEMM
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gemm_gpu.py tests/test_ctest_dropin.py -m gpu -x -q -k "gemm3m or 3m" > gpurun_out/r02_gemm3m_pytest.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r02_gemm3m_pytest.log
timeout 120 python tools/gemm3m_time.py > gpurun_out/r02_gemm3m_vs_gemm.jsonl 2> gpurun_out/r02_gemm3m_time.err; echo "time rc=$?"
cat gpurun_out/r02_gemm3m_vs_gemm.jsonl; tail -3 gpurun_out/r02_gemm3m_time.err
