#!/bin/bash
# round 2, GPU call 6 (2 GPUs): the library's SUMMA driver -- correctness (flags + copy-engine pulls, NCCL-sync, NCCL transport), timing, bench N=2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
{
echo "=== one process, 1x1 grid"; timeout 120 python tools/summa_c_check.py 1000 900 800 128 2>&1 | tail -8
echo "=== default (memops flags + pulls)"; timeout 170 $TR tools/summa_c_check.py 3000 2500 2200 256 --time 16384 32768 16384 4096 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -25
echo "=== B200_SUMMA_SYNC=nccl"; B200_SUMMA_SYNC=nccl timeout 170 $TR tools/summa_c_check.py 3000 2500 2200 256 --time 16384 32768 16384 4096 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -12
echo "=== B200_SUMMA_TRANSPORT=nccl"; B200_SUMMA_TRANSPORT=nccl timeout 170 $TR tools/summa_c_check.py 3000 2500 2200 256 --time 16384 32768 16384 4096 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -12
echo "=== 2x1 grid"; B200_SUMMA_GRID=2x1 timeout 170 $TR tools/summa_c_check.py 3000 2500 2200 256 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -12
} > gpurun_out/r02_summa_c_2gpu.log 2>&1
cat gpurun_out/r02_summa_c_2gpu.log
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_bench_n2.err; cat gpurun_out/r02_bench_n2.json
timeout 300 $TR bench.py --gpus 2 --steps 5 --warmup 3 --summa py --nb 2048 --no-e2e > gpurun_out/r02_bench_n2_py.json 2> gpurun_out/r02_bench_n2_py.err; echo "bench py rc=$?"
cut -c1-400 gpurun_out/r02_bench_n2_py.json
