#!/bin/bash
# round 2, GPU call 24 (2 GPUs): host-C column sweeps of the SUMMA driver -- correctness with the limit lowered, then the N=2 bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521"
{
echo "=== one process, halves from 64 columns"; B200_SUMMA_HOST_HALVES=64 timeout 120 python tools/summa_c_check.py 1000 900 800 128 2>&1 | tail -8
echo "=== 2 GPUs, halves from 64 columns"; B200_SUMMA_HOST_HALVES=64 timeout 170 $TR tools/summa_c_check.py 3000 2500 2200 256 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -14
} > gpurun_out/r02_summa_c_2gpu_halves.log 2>&1
cat gpurun_out/r02_summa_c_2gpu_halves.log | cut -c1-200
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench rc=$?"
tail -2 gpurun_out/r02_bench_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_n2.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "parity", d["parity"]["worst_ratio"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["max_abs_diff_vs_device_path_rank0"])
PY
B200_SUMMA_HOST_HALVES=0 timeout 600 $TR bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_n2_nohalves.json 2>/dev/null
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_n2_nohalves.json").read().strip().splitlines()[-1])
print("no halves: e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
PY
