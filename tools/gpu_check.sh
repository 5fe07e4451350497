#!/bin/bash
# One GPU-box session: host facts, optional peak microbenchmarks, smoke, GPU tests, sweep, bench.
# Outputs in gpurun_out/.  Every stage runs under its own timeout so a hung kernel cannot hold the box.
set -u
mkdir -p gpurun_out
{
  echo "== host"; nproc; lscpu | grep -E "Model name|Socket|Core|Thread" | cut -c1-200; free -g | head -2
  echo "== gpu"; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit,memory.total --format=csv
} > gpurun_out/host.txt 2>&1
if [ "${RUN_PEAKS:-0}" = "1" ]; then timeout 300 tools/peaks > gpurun_out/peaks.json 2> gpurun_out/peaks.err; fi
if [ "${RUN_SMOKE:-1}" = "1" ]; then
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
  tail -8 gpurun_out/smoke.log
fi
if [ "${RUN_TESTS:-1}" = "1" ]; then
  timeout ${PYTEST_TIMEOUT:-1500} python -m pytest tests -m gpu -q ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
  tail -15 gpurun_out/pytest_gpu.log
fi
if [ "${RUN_SWEEP:-0}" = "1" ]; then
  timeout 900 python bench.py --sweep ${SWEEP_ARGS:-} > gpurun_out/sweep.jsonl 2> gpurun_out/sweep.err; echo "sweep exit $?"; cat gpurun_out/sweep.jsonl; tail -3 gpurun_out/sweep.err
fi
if [ "${RUN_BENCH:-0}" = "1" ]; then
  timeout 900 python bench.py ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
fi
