#!/bin/bash
# round 2, GPU call 21: ZSYRK / ZHERK / ZSYR2K with the compact 1:2 triangle enumeration; level-3 tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_level3_gpu.py tests/test_f_rows_gpu.py -m gpu -q -x > gpurun_out/r02_pytest_call21.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02_pytest_call21.log
timeout 600 python bench.py --sweep-level3 --sweep-dtypes z --sizes 8192 2> gpurun_out/r02_level3_sweep.err | grep "syr\|her" | cut -c1-160
