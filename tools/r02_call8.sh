#!/bin/bash
# round 2, GPU call 8: level-3 tests (register-resident triangular block kernel, SYMM panels, GEMMT), level-3 sweep at 8192 (new vs round-1 block kernel), config-1 harness table
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_level3_gpu.py tests/test_f_rows_gpu.py "tests/test_ctest_dropin.py::test_ctest_level3_gemm" -m gpu -q -x > gpurun_out/r02_pytest_call8.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r02_pytest_call8.log
timeout 600 python bench.py --sweep-level3 --sweep-dtypes d,s --sizes 8192 > gpurun_out/r02_level3_sweep_8192.jsonl 2> gpurun_out/r02_level3_sweep.err
B200_TRI_KERNEL=smem timeout 600 python bench.py --sweep-level3 --sweep-dtypes d --sizes 8192 > gpurun_out/r02_level3_sweep_8192_round1_block_kernel.jsonl 2>> gpurun_out/r02_level3_sweep.err
cut -c1-220 gpurun_out/r02_level3_sweep_8192.jsonl; echo ---; cut -c1-220 gpurun_out/r02_level3_sweep_8192_round1_block_kernel.jsonl
B=oracle/_ref/bench
T=$(python -c "from oracle import cpu; print(cpu.best_target())")
{
  echo "# BASELINE config 1: benchmark/gemm.c (the reference's harness: Fortran ABI, malloc'd host buffers, no warm-up), NN 1024^3, alpha=1 beta=0"
  echo "# columns: harness output (MFlops over ALL loops, first call included) for this library and for the reference ($T build) on the box's host cores"
  nproc
  for d in d s z c; do
    for loops in 1 200 2000; do
      echo "## ${d}gemm 1024^3 OPENBLAS_LOOPS=$loops: libopenblas_b200.so"
      OPENBLAS_LOOPS=$loops timeout 300 $B/${d}gemm.b200 1024 1024 1 | tail -1
      echo "## ${d}gemm 1024^3 OPENBLAS_LOOPS=$loops: reference ($T), all host cores"
      OPENBLAS_LOOPS=$loops timeout 300 $B/${d}gemm.$T 1024 1024 1 | tail -1
    done
  done
  echo "## dgemm 4096^3 OPENBLAS_LOOPS=20: libopenblas_b200.so, then reference"
  OPENBLAS_LOOPS=20 timeout 300 $B/dgemm.b200 4096 4096 1 | tail -1
  OPENBLAS_LOOPS=20 timeout 300 $B/dgemm.$T 4096 4096 1 | tail -1
  echo "## steady state per call (B200_TRACE=1, calls 2..6 of dgemm 1024^3)"
  B200_TRACE=1 OPENBLAS_LOOPS=6 timeout 120 $B/dgemm.b200 1024 1024 1 2>&1 | grep trace
} > gpurun_out/r02_config1_harness_table.txt 2>&1
cat gpurun_out/r02_config1_harness_table.txt
