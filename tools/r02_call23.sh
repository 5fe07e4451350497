#!/bin/bash
# round 2, GPU call 23: library with the helper-warp S/CGEMM kernel: kernel tests, S and C sweep at 8192 over the op combinations
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_round2_gpu.py tests/test_gemm_gpu.py -m gpu -q -x -k "ws_tma or ssyrk or tile_aligned or golden or ragged or both_tile" > gpurun_out/r02_pytest_call23.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02_pytest_call23.log
timeout 600 python bench.py --sweep --sweep-dtypes s,c --sizes 8192 --all-ops 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['dtype'], d['n'], d['op'], round(d['tflops_real'], 1), d['kernel'])"
