#!/bin/bash
# round 2, GPU call 13 (8 GPUs): the library's SUMMA driver on the 2x4 grid -- correctness (ragged problem, all operand kinds), timing at 32768^3, bench N=8 (device + host-shard e2e + sampled parity), the NCCL transport beside it
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
{
echo "=== default (flags + copy-engine pulls), $N GPUs"; B200_SUMMA_HOST_HALVES=64 timeout 240 $TR tools/summa_c_check.py 3000 2500 2200 256 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -50
} > gpurun_out/r02_summa_c_${N}gpu.log 2>&1
grep -c " OK" gpurun_out/r02_summa_c_${N}gpu.log; grep "MISMATCH\|ALL OK\|Error\|error" gpurun_out/r02_summa_c_${N}gpu.log | head
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_bench_n$N.err; cut -c1-600 gpurun_out/r02_bench_n$N.json
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_bench_n$N.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "parity", d["parity"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["max_abs_diff_vs_device_path_rank0"], "launches", d["gpu_launches"])
PY
