#!/bin/bash
# round 2, GPU call 22: SGEMM lab -- helper warps for the staged-tile transposition
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/sgemm_lab.txt
tools/sgemm_lab.sh run 8192 11 12 13 14 > /dev/null 2>&1
grep "TFLOP\|WRONG\|launch\|run ->" gpurun_out/sgemm_lab.txt
timeout 100 tools/sgemm_lab 12 1000 2 1000 77 513 >> gpurun_out/sgemm_lab.txt 2>&1; tail -8 gpurun_out/sgemm_lab.txt | grep -v "^$" | cut -c1-160
