#!/bin/bash
# round 2, GPU call 10: DTRSM launch list (per-kernel durations) with the register block kernel, TRMM/TRSM sweep and tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_dtrsm8192_launches.csv python tools/trsm_once.py 8192 > gpurun_out/r02_trsm_once.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02_dtrsm8192_launches.csv")) if len(r) > 5]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value")
rows = rows[1:]
half = rows[len(rows) // 2:]        # the second (measured) call
tot = collections.defaultdict(lambda: [0, 0.0])
for r in half:
    name = r[ik].split("(")[0][-60:]
    tot[name][0] += 1; tot[name][1] += float(r[iv].replace(",", "")) / 1e3
for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:5d} launches {us/1e3:9.3f} ms total {us/n:9.1f} us each  {k}")
PY
timeout 900 python -m pytest tests/test_level3_gpu.py -m gpu -q -x -k "trxm or trsm or golden" > gpurun_out/r02_pytest_call10.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02_pytest_call10.log
timeout 600 python bench.py --sweep-level3 --sweep-dtypes d,s --sizes 8192 2> gpurun_out/r02_level3_sweep.err | grep "trmm\|trsm" > gpurun_out/r02_level3_sweep_8192_trxm.jsonl
cut -c1-200 gpurun_out/r02_level3_sweep_8192_trxm.jsonl
