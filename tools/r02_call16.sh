#!/bin/bash
# round 2, GPU call 16: bench N=1 with the SBGEMM size sweep in extras
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"], d["roofline"].get("traffic_source"))
e = d["extra"]
print("sb8192", {k: e["sbgemm_8192"][k] for k in ("burst_tflops", "sustained_tflops", "other_ops_tflops", "frac_of_burst_peak", "sustained_frac_of_sustained_peak")})
for k, v in e["sbgemm_sweep"].items():
    print(k, {a: round(b) for a, b in v["tflops"].items()}, round(v["frac_of_burst_peak"], 3), v["kernel"])
for k, v in e["square_sweep_nn"].items():
    print(k, round(v["tflops"], 1), round(v["frac"], 3), v["kernel"])
print("sgemm", e["sgemm_16384"]["tflops_real_flops"])
PY
