#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -q -m gpu -x -k "all_ops or tile or conj or golden" 2>&1 | tail -3
timeout 300 ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,smsp__inst_executed_op_shared_ld.sum --csv --log-file gpurun_out/lds_probe_ncu.csv ./build/lds_probe > gpurun_out/lds_probe.txt 2>&1
python bench.py --sweep --sweep-dtypes s --sizes 1024,2048,4096,8192,16384 --all-ops 2>/dev/null > gpurun_out/lane_s.jsonl
python bench.py --sweep --sweep-dtypes c --sizes 4096,8192 --all-ops 2>/dev/null > gpurun_out/lane_c.jsonl
B200_SGEMM_CFG=1 python bench.py --sweep --sweep-dtypes s --sizes 8192 --all-ops 2>/dev/null > gpurun_out/lane_s_pw.jsonl
python - <<PY
import json
for f in ("lane_s","lane_c","lane_s_pw"):
    for l in open(f"gpurun_out/{f}.jsonl"):
        try: d=json.loads(l)
        except Exception: continue
        print(f,d["dtype"],d["n"],d.get("op"),round(d["ms"],4),round(d["tflops_real"],1),d["kernel"])
PY
