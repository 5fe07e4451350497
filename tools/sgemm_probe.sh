#!/bin/bash
# SGEMM experiments: correctness of the current build, then timing probes (results of probe runs are garbage by design)
timeout 600 python -m pytest tests/test_gemm_gpu.py -q -m gpu -x -k "golden or all_ops or tile" 2>&1 | tail -2
for p in 0 1 2; do
  B200_SGEMM_PROBE=$p python bench.py --sweep --sweep-dtypes s --sizes 8192 --all-ops 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('probe$p',d['dtype'],d['n'],d.get('op'),round(d['ms'],3),round(d['tflops_real'],1),d['kernel'])"
done
