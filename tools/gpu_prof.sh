#!/bin/bash
# ncu captures of the top kernels (one GPU). Usage: KERNELS="d:4096 s:4096" [TAG=x] bash tools/gpu_prof.sh
set -u
mkdir -p gpurun_out
for spec in ${KERNELS:-d:4096}; do
  dt=${spec%%:*}; n=${spec##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm" -s 2 -c 1 -f -o gpurun_out/prof_${dt}_${n}${TAG:-} \
     python tools/prof_one.py $dt $n 0 0 3 > gpurun_out/prof_${dt}_${n}${TAG:-}.log 2>&1
  tail -2 gpurun_out/prof_${dt}_${n}${TAG:-}.log
done
