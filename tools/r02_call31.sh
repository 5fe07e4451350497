#!/bin/bash
# round 2, GPU call 31 (after SBGEMMT, in-place conversions, trimmed tests, level-3 rows in the bench line): the whole GPU suite, smoke(), the N=1 bench line and the reference arm (short) as the driver will run them
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r02_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"
tail -8 gpurun_out/r02_smoke.log
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "launches", d["gpu_launches"], "parity", d["parity"]["worst_ratio"])
e = d["extra"]
for k in ("zgemm_8192", "cgemm_8192"):
    print(k, {a: round(b, 1) for a, b in e[k]["tflops_real_flops"].items()}, "parity", e[k]["parity_worst_ratio"])
for k, v in e["tall_skinny"].items():
    print(k, {a: round(b, 1) for a, b in v["tflops_real_flops"].items()}, "ms", round(v["measured_ms"], 1), "parity", v["parity_worst_ratio"])
for k, v in e["square_sweep_nn"].items():
    print(k, round(v["tflops"], 1), round(v["frac"], 3))
PY
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_n1.json").read().strip().splitlines()[-1])
for k, v in d["extra"]["level3_8192"].items():
    if isinstance(v, dict):
        print(k, round(v["ms"], 2), "ms", round(v.get("tflops_useful", v.get("tflops_triangle", 0)), 1), "TF")
PY
