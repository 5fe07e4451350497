#!/bin/bash
# round 2, GPU call 2: inner-loop probe + SGEMM lab sweep (second set) + ncu of two lab configurations
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_call2_gpu.txt
timeout 120 tools/ffma_probe > gpurun_out/r02_ffma_probe.txt 2>&1
rm -f gpurun_out/sgemm_lab.txt
tools/sgemm_lab.sh run 8192 0 1 2 3 4 5 6 7 8 9 10 > /dev/null 2>&1
for c in 1 2; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:sgemm_ws -c 1 -o gpurun_out/r02_lab_cfg$c -f tools/sgemm_lab $c 4096 1 > gpurun_out/r02_lab_cfg$c.log 2>&1
done
timeout 100 tools/sgemm_lab 2 16384 3 >> gpurun_out/sgemm_lab.txt 2>&1
cat gpurun_out/r02_ffma_probe.txt
grep TFLOP gpurun_out/sgemm_lab.txt
