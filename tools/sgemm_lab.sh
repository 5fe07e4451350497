#!/bin/bash
# Builds tools/sgemm_lab (here, no GPU needed) or runs every configuration of it (on the GPU box).
#   tools/sgemm_lab.sh build
#   tools/sgemm_lab.sh run [n] [cfgs...]     -> gpurun_out/sgemm_lab.txt
set -u
cd "$(dirname "$0")/.."
if [ "${1:-}" = build ]; then
  make -C openblas_b200/csrc -j"$(nproc)" >/dev/null || exit 1
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -DB200_LAB_SPIN_LIMIT=400000000u \
       -Xptxas -v -o tools/sgemm_lab tools/sgemm_lab.cu build/csrc/sgemm_ffma.o -lcuda 2> build/sgemm_lab_ptxas.txt || { tail -30 build/sgemm_lab_ptxas.txt; exit 1; }
  grep -A1 "sgemm_ws_kernel" build/sgemm_lab_ptxas.txt | grep -E "registers|spill" | sort | uniq -c | sort -rn | head -40
  exit 0
fi
shift
n=${1:-8192}; shift || true
cfgs=${*:-0 1 2 3 4 5 6 7 8 9 10 11 12}
mkdir -p gpurun_out
for c in $cfgs; do
  timeout 300 tools/sgemm_lab "$c" "$n" 5
done 2>&1 | tee -a gpurun_out/sgemm_lab.txt
