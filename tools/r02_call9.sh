#!/bin/bash
# round 2, GPU call 9: SGEMM lab (B tile with duplicated scalars), TRMM/TRSM with the lane-group register kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/sgemm_lab.txt
tools/sgemm_lab.sh run 8192 11 12 13 14 15 > /dev/null 2>&1
grep "TFLOP\|WRONG\|launch\|run ->" gpurun_out/sgemm_lab.txt
timeout 900 python -m pytest tests/test_level3_gpu.py -m gpu -q -x -k "trxm or trsm or golden" > gpurun_out/r02_pytest_call9.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r02_pytest_call9.log
timeout 600 python bench.py --sweep-level3 --sweep-dtypes d,s --sizes 8192 2> gpurun_out/r02_level3_sweep.err | grep "trmm\|trsm" > gpurun_out/r02_level3_sweep_8192_trxm.jsonl
cut -c1-200 gpurun_out/r02_level3_sweep_8192_trxm.jsonl
