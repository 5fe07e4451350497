#!/bin/bash
# TEST INFRASTRUCTURE (lives under tests/ because it executes binaries from oracle/_ref/): BASELINE.json
# configs through the reference's own benchmark/gemm.c harness -- once linked against libopenblas_b200.so,
# once against the reference itself on the host cores -- and through bench.py; outputs in gpurun_out/.
# The run recorded in profiles/r01_config1_reference_harness.txt came from this script.
set -u
mkdir -p gpurun_out
B=oracle/_ref/bench
T=$(python -c "from oracle import cpu; print(cpu.best_target())")
{
  echo "# config 1: benchmark/gemm.c (reference harness, Fortran ABI, malloc'd host buffers) DGEMM 1024^3 NN alpha=1 beta=0"
  echo "## linked against libopenblas_b200.so (GPU, staging inside the call)"
  OPENBLAS_LOOPS=20 timeout 120 $B/dgemm.b200 1024 1024 1
  OPENBLAS_LOOPS=5 timeout 120 $B/dgemm.b200 4096 4096 1
  OPENBLAS_LOOPS=20 timeout 120 $B/sgemm.b200 1024 1024 1
  OPENBLAS_LOOPS=20 timeout 120 $B/zgemm.b200 1024 1024 1
  OPENBLAS_LOOPS=20 timeout 120 $B/cgemm.b200 1024 1024 1
  OPENBLAS_LOOPS=20 timeout 120 $B/sbgemm.b200 1024 1024 1
  echo "## linked against the reference (oracle/_ref/$T), OPENBLAS_NUM_THREADS=$(nproc)"
  OPENBLAS_NUM_THREADS=$(nproc) OPENBLAS_LOOPS=20 timeout 120 $B/dgemm.$T 1024 1024 1
  OPENBLAS_NUM_THREADS=$(nproc) OPENBLAS_LOOPS=5 timeout 120 $B/dgemm.$T 4096 4096 1
  OPENBLAS_NUM_THREADS=$(nproc) OPENBLAS_LOOPS=20 timeout 120 $B/sgemm.$T 1024 1024 1
  OPENBLAS_NUM_THREADS=$(nproc) OPENBLAS_LOOPS=20 timeout 120 $B/zgemm.$T 1024 1024 1
  OPENBLAS_NUM_THREADS=$(nproc) OPENBLAS_LOOPS=20 timeout 120 $B/cgemm.$T 1024 1024 1
  OPENBLAS_NUM_THREADS=$(nproc) OPENBLAS_LOOPS=20 timeout 120 $B/sbgemm.$T 1024 1024 1
  OPENBLAS_NUM_THREADS=1 OPENBLAS_LOOPS=5 timeout 120 $B/dgemm.$T 1024 1024 1
} > gpurun_out/config1_reference_harness.txt 2>&1
timeout 900 python bench.py --sweep --sweep-dtypes d,s --sizes 1024,2048,4096,8192,12288,16384 --all-ops > gpurun_out/sweep_config2_ds.jsonl 2>/dev/null
timeout 600 python bench.py --sweep --sweep-dtypes sb --sizes 8192 --all-ops > gpurun_out/sweep_config3_sb.jsonl 2>/dev/null
timeout 600 python bench.py --sweep --sweep-dtypes z,c --sizes 4096,8192 --all-ops > gpurun_out/sweep_config5_zc.jsonl 2>/dev/null
cat gpurun_out/config1_reference_harness.txt
