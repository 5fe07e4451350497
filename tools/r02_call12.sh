#!/bin/bash
# round 2, GPU call 12: TRMM/TRSM sweep with the single-wave 128-block kernel, then the WHOLE gpu test suite and the N=1 bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --sweep-level3 --sweep-dtypes d,s,z,c --sizes 8192 > gpurun_out/r02_level3_sweep_8192.jsonl 2> gpurun_out/r02_level3_sweep.err
cut -c1-200 gpurun_out/r02_level3_sweep_8192.jsonl | grep "trmm\|trsm"
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_bench_n1.err
cut -c1-1500 gpurun_out/r02_bench_n1.json
