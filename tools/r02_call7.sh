#!/bin/bash
# round 2, GPU call 7: fixed tests + new SBGEMM coverage; ncu of the SGEMM TMA kernel; config-1 harness with the host-path trace
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_round2_gpu.py tests/test_f_rows_gpu.py::test_gemmt_error_exits_match_the_oracle_table "tests/test_ctest_dropin.py::test_ctest_level3_gemm3m" tests/test_gemm_gpu.py -m gpu -q -x > gpurun_out/r02_pytest_call7.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/r02_pytest_call7.log
for op in NN NT; do
  ta=0; tb=0; [ $op = NT ] && tb=1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:sgemm_ws -c 1 -o gpurun_out/r02_sgemm_ws_$op -f python tools/prof_one.py s 8192 $ta $tb > gpurun_out/r02_ncu_$op.log 2>&1
  ncu -i gpurun_out/r02_sgemm_ws_$op.ncu-rep --page raw --csv > gpurun_out/r02_sgemm_ws_${op}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r02_sgemm_ws_$op.ncu-rep --page source --csv > gpurun_out/r02_sgemm_ws_${op}_source.csv 2>/dev/null
  ncu -i gpurun_out/r02_sgemm_ws_$op.ncu-rep --page details > gpurun_out/r02_sgemm_ws_${op}_details.txt 2>/dev/null
done
ls -la gpurun_out/*.ncu-rep
B=oracle/_ref/bench
{
  echo "# benchmark/gemm.c linked against libopenblas_b200.so, B200_TRACE=1, dgemm 1024^3 x 6 calls and 2048^3 x 4"
  B200_TRACE=1 OPENBLAS_LOOPS=6 timeout 120 $B/dgemm.b200 1024 1024 1 2>&1 | tail -8
  B200_TRACE=1 OPENBLAS_LOOPS=4 timeout 120 $B/dgemm.b200 2048 2048 1 2>&1 | tail -6
} > gpurun_out/r02_config1_trace.txt 2>&1
cat gpurun_out/r02_config1_trace.txt
