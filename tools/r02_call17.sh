#!/bin/bash
# round 2, GPU call 17 (4 GPUs): the SUMMA driver on the 2x2 grid -- correctness and the N=4 bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
timeout 240 $TR tools/summa_c_check.py 3000 2500 2200 256 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -30 > gpurun_out/r02_summa_c_${N}gpu.log
grep -c " OK" gpurun_out/r02_summa_c_${N}gpu.log; grep "MISMATCH\|ALL OK\|Error\|error" gpurun_out/r02_summa_c_${N}gpu.log | head
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_bench_n$N.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_n4.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "parity", d["parity"]["worst_ratio"], d["parity"]["ok"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
PY
