#!/bin/bash
# round 2, GPU call 5: the new f-row tests (gemmt / sbgemv / sbdot / 3M ctest / unstubbed test_sbgemm) and the round-2 kernel tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_f_rows_gpu.py tests/test_round2_gpu.py "tests/test_ctest_dropin.py::test_ctest_level3_gemm3m" tests/test_ctest_dropin.py::test_compare_sgemm_sbgemm -m gpu -q > gpurun_out/r02_pytest_frows.log 2>&1; echo "pytest rc=$?"
tail -60 gpurun_out/r02_pytest_frows.log
