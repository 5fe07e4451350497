#!/bin/bash
# forced-tile comparison for the DGEMM / SGEMM tile heuristic
for t in 64 128; do
  B200_DGEMM_TILE=$t B200_SGEMM_TILE=$t python bench.py --sweep --sweep-dtypes d,s --sizes 512,1024,1536,2048,3072,4096 2>/dev/null | sed "s/^/tile$t /"
done
