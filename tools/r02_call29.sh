#!/bin/bash
# round 2, GPU call 29: TRMM / TRSM with the lower recursion levels run as slices of B on several streams (tri_forked):
# parity tests with the slices on, then lanes x threshold sweep at 8192 for d, s, z, c
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_level3_gpu.py tests/test_full_size_gpu.py -m gpu -x -q -k "trsm or trmm or trxm or tri" > gpurun_out/r02_lanes_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02_lanes_pytest.log
: > gpurun_out/r02_tri_lanes_sweep.jsonl
for cfg in "1 1024" "2 1024" "4 512" "4 1024" "4 2048" "3 1024"; do
  set -- $cfg
  echo "{\"lanes\": $1, \"below\": $2}" >> gpurun_out/r02_tri_lanes_sweep.jsonl
  B200_TRI_LANES=$1 B200_TRI_LANE_BELOW=$2 timeout 300 python bench.py --sweep-level3 --sizes 8192 --sweep-dtypes d,s,z,c --sweep-routines trmm,trsm 2>/dev/null | grep routine >> gpurun_out/r02_tri_lanes_sweep.jsonl
done
python - <<'PY'
import json
cfg = None
for ln in open("gpurun_out/r02_tri_lanes_sweep.jsonl"):
    d = json.loads(ln)
    if "lanes" in d:
        cfg = (d["lanes"], d["below"]); print("== lanes", cfg)
    else:
        print("  %-22s %7.2f ms %6.1f TF  launches %d" % (d["routine"], d["ms"], d["tflops_useful"], d["launches_per_call"]))
PY
