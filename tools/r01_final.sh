#!/bin/bash
# round-1 closing measurements on one GPU: batch path, ncu traffic of the final kernels, bench + launch list
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -m gpu -k "batch or golden" 2>&1 | tail -3
for n in 32 64 128; do
  B200_BATCH_PACKED=1 timeout 120 python tools/batch_time.py $n 1000 5
  B200_BATCH_PACKED=0 timeout 120 python tools/batch_time.py $n 1000 5
done 2>&1 | tee gpurun_out/batch_time.txt
KERNELS="d:16384 sb:8192" TAG=_final2 bash tools/gpu_prof.sh
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1
tail -3 gpurun_out/launches.csv
