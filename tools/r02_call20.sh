#!/bin/bash
# round 2, GPU call 20: SYRK family on the TMA S/CGEMM kernel with the compact 2:1 triangle enumeration
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_round2_gpu.py tests/test_level3_gpu.py -m gpu -q -x -k "ssyrk or level3_all_flag or golden or scaling" > gpurun_out/r02_pytest_call20.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02_pytest_call20.log
timeout 600 python bench.py --sweep-level3 --sweep-dtypes s,c --sizes 8192 2> gpurun_out/r02_level3_sweep.err | grep "syr\|her" | cut -c1-160
