#!/bin/bash
# round 2, GPU call 1: inner-loop probe + SGEMM lab sweep + ncu of two lab configurations
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_call1_gpu.txt
timeout 120 tools/ffma_probe > gpurun_out/r02_ffma_probe.txt 2>&1
rm -f gpurun_out/sgemm_lab.txt
tools/sgemm_lab.sh run 8192 0 1 2 3 4 5 6 7 8 9 10 11 12 > /dev/null 2>&1
for c in 1 10; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:sgemm_ws -c 1 -o gpurun_out/r02_lab_cfg$c -f tools/sgemm_lab $c 4096 1 > gpurun_out/r02_lab_cfg$c.log 2>&1
done
tail -60 gpurun_out/sgemm_lab.txt
