#!/bin/bash
# round 2, GPU call 3: GPU test suite, config-1 harness (first call / steady state), S and C sweeps
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1
tail -15 gpurun_out/r02_pytest_gpu.log
B=oracle/_ref/bench
{
  echo "# benchmark/gemm.c (reference harness, Fortran ABI, malloc'd host buffers) linked against libopenblas_b200.so"
  for d in d s z c; do
    echo "## ${d}gemm 1024^3: OPENBLAS_LOOPS=1 (first call: CUDA init + allocations inside the timed loop), then 200"
    OPENBLAS_LOOPS=1 timeout 120 $B/${d}gemm.b200 1024 1024 1 | tail -1
    OPENBLAS_LOOPS=200 timeout 120 $B/${d}gemm.b200 1024 1024 1 | tail -1
  done
  OPENBLAS_LOOPS=1 timeout 120 $B/dgemm.b200 4096 4096 1 | tail -1
  OPENBLAS_LOOPS=20 timeout 120 $B/dgemm.b200 4096 4096 1 | tail -1
} > gpurun_out/r02_config1_harness.txt 2>&1
cat gpurun_out/r02_config1_harness.txt
timeout 600 python bench.py --sweep --sweep-dtypes s,c --sizes 1024,2048,4096,8192 --all-ops > gpurun_out/r02_sweep_sc.jsonl 2>gpurun_out/r02_sweep_sc.err
cut -c1-200 gpurun_out/r02_sweep_sc.jsonl
