#!/bin/bash
# round 2, GPU call 18: complex TRMM/TRSM on the register block kernel -- level-3 tests, ctest drivers (c, z), sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_level3_gpu.py "tests/test_ctest_dropin.py::test_ctest_level3_gemm[c]" "tests/test_ctest_dropin.py::test_ctest_level3_gemm[z]" -m gpu -q -x > gpurun_out/r02_pytest_call18.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02_pytest_call18.log
timeout 600 python bench.py --sweep-level3 --sweep-dtypes z,c --sizes 8192 2> gpurun_out/r02_level3_sweep.err | grep "trmm\|trsm" | cut -c1-160
