#!/bin/bash
# round 2, GPU call 15: DGEMM 64x64 producer-warp kernel with two CTAs per SM (B200_DGEMM_MID=2) against one: correctness, mid sizes, DTRSM
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B200_DGEMM_MID=2 timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x -k "both_tile_sizes or golden or ragged" > gpurun_out/r02_pytest_call15.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02_pytest_call15.log
for mid in 1 2; do
  echo "== B200_DGEMM_MID=$mid"
  B200_DGEMM_MID=$mid timeout 300 python bench.py --sweep --sweep-dtypes d --sizes 512,1024,1536,2048,3072 --all-ops 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['n'], d['op'], round(d['tflops_real'], 1), d['kernel'])"
  B200_DGEMM_MID=$mid timeout 300 python bench.py --sweep-level3 --sweep-dtypes d --sizes 8192 2>/dev/null | grep "trmm\|trsm" | cut -c1-150
done
