#!/bin/bash
# round 2, GPU call 14: new SUMMA C-ABI tests, ncu launch list of the bench command, ncu --set full of the DGEMM kernel at 16384^3 (DRAM traffic), bench N=1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_round2_gpu.py -m gpu -q -x -k "summa" > gpurun_out/r02_pytest_call14.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02_pytest_call14.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_n1.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r02_bench_under_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02_launches_bench_n1.csv")) if len(r) > 5]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value")
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    name = r[ik].split("(")[0][-70:]
    tot[name][0] += 1; tot[name][1] += float(r[iv].replace(",", "")) / 1e6
for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"{n:5d} launches {ms:10.3f} ms  {k}")
PY
timeout 600 ncu --set full --clock-control none -k regex:dgemm_dmma -c 1 -o gpurun_out/r02_dgemm_16384 -f python tools/prof_one.py d 16384 0 0 2 > gpurun_out/r02_ncu_d16384.log 2>&1
ncu -i gpurun_out/r02_dgemm_16384.ncu-rep --page raw --csv > gpurun_out/r02_dgemm_16384_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_dgemm_16384.ncu-rep --page details > gpurun_out/r02_dgemm_16384_details.txt 2>/dev/null
python - <<'PY'
import csv, json
rows = list(csv.reader(open("gpurun_out/r02_dgemm_16384_raw.csv")))
d = {h: (v, u) for h, u, v in zip(rows[0], rows[1], rows[2])}
def val(k):
    v, u = d[k]; x = float(v.replace(",", ""))
    return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
rec = {"dtype": "d", "m": 16384, "n": 16384, "k": 16384, "kernel": d["Kernel Name"][0] if "Kernel Name" in d else "dgemm_dmma",
       "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
       "gpu_time_ms": float(d["gpu__time_duration.sum"][0].replace(",", "")) * {"ms": 1.0, "us": 1e-3, "s": 1e3}.get(d["gpu__time_duration.sum"][1], 1.0),
       "tensor_pipe_dmma_pct": d.get("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", ("", ""))[0],
       "lts_hit_rate_pct": d.get("lts__t_sector_hit_rate.pct", ("", ""))[0],
       "source": "profiles/r02_dgemm_16384_ncu.json: ncu --set full --clock-control none, one launch (tools/r02_call14.sh)"}
json.dump(rec, open("gpurun_out/r02_dgemm_16384_ncu.json", "w"), indent=1)
print(rec)
PY
rm -f gpurun_out/r02_dgemm_16384.ncu-rep
