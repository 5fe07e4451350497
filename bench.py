#!/usr/bin/env python3
"""bench.py -- headline benchmark of the GEMM hot path (BASELINE.json: "?GEMM TFLOP/s (2mnk/t)").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--dtype d|s|z|c|sb] [--sweep]

N = 1   one step = one DGEMM 16384^3, NN, column-major, alpha = 1, beta = 0 (BASELINE config 2,
        the size the ">= 80 % of FP64 peak" target is quoted on), operands resident in HBM, through
        the library's device entry point; `e2e` = the same GEMM through cblas_dgemm with pinned
        HOST buffers (H2D of A and B and D2H of C inside the timed region).
N > 1   weak scaling, 2*16384^3 flops per GPU: 1x2 -> 16384 x 32768 x 16384, 2x2 -> 32768 x 32768 x
        16384, 2x4 -> 32768^3 (BASELINE config 4), 2-D block-cyclic SUMMA (openblas_b200/summa.py),
        one process per GPU under torchrun, NCCL panel broadcasts overlapped with the local DGEMM.
--impl reference   the reference's own CPU implementation (oracle/_ref, built from /root/reference
        by oracle/build_ref.py) on the box's host cores, a bounded sample of the same workload.

Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DT = {"s": 0, "d": 1, "c": 2, "z": 3, "sb": 4}
FLOP_FACTOR = {"s": 2.0, "d": 2.0, "c": 8.0, "z": 8.0, "sb": 2.0}   # real flops per m*n*k
# Dense peaks in TFLOP/s per pipe.  bf16: MEASURED_PEAKS.json (driver-measured).  FP64 / FP32 are not in
# that file: measured on this pool's B200 with tools/peaks.cu (profiles/r01_peaks_microbench.json):
# DMMA.8x8x4 issue rate 36.8 TFLOP/s, FFMA 71.1 TFLOP/s (cuBLAS: dgemm 36.0, sgemm-pedantic 66.8).
PEAK_FALLBACK = {"d": 36.8, "z": 36.8, "s": 71.1, "c": 71.1, "sb": 1590.0}
# DRAM traffic of the dominant kernel per launch (dram__bytes_read.sum + dram__bytes_write.sum) from one
# `ncu --set full` capture of the same shape: profiles/r01_prof_d_16384_final2_summary.txt
NCU_TRAFFIC_BYTES = {("d", 16384, 16384, 16384): 54.017287e9 + 2.149455e9}


def measured_peak(dtype):
    if dtype == "sb":
        try:
            p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            return float(p["bf16_tflops"]), "MEASURED_PEAKS.json bf16_tflops (burst)"
        except Exception:
            return PEAK_FALLBACK["sb"], "fallback 1.59 PFLOP/s (B200_PROFILING.md)"
    try:
        p = json.load(open(os.path.join(ROOT, "profiles", "r01_peaks_microbench.json")))
        if dtype in ("d", "z"):
            return max(p[k] for k in p if k.startswith("dmma_tflops")), "tools/peaks.cu DMMA issue-rate microbenchmark on this pool (MEASURED_PEAKS.json has no FP64 entry)"
        return max(p[k] for k in p if k.startswith("ffma_tflops")), "tools/peaks.cu FFMA issue-rate microbenchmark on this pool (MEASURED_PEAKS.json has no FP32 entry)"
    except Exception:
        return PEAK_FALLBACK[dtype], "tools/peaks.cu value recorded in bench.py"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except (ValueError, IndexError):
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        busy = [x for x in sm if x > 300] or sm
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def workload_config(dtype, m, n, k, parallelism):
    """The `config` object both arms print (same workload naming for the driver's ratio)."""
    return {"workload": f"{dtype}gemm NN column-major {m}x{n}x{k} alpha=1 beta=0 (BASELINE configs[1] at N=1; configs[3] shape at N=8)",
            "parallelism": parallelism, "l2_policy": "inputs larger than L2 (A,B >= 2 GiB each vs 126 MB L2)",
            "inputs": "uniform(-0.5,0.5), fixed seed, resident in HBM"}


def weak_shape(world):
    return {1: (16384, 16384, 16384), 2: (16384, 32768, 16384), 4: (32768, 32768, 16384), 8: (32768, 32768, 32768)}.get(
        world, (16384, 16384 * world, 16384))


# ---------------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle import cpu
    kind = "reference" if cpu.have_reference() else "port"
    dtype = args.dtype
    n = args.ref_n
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(0)
    npdt = cpu.NP_IN[DT[dtype]]

    def operand():
        x = rng.random((n, n), dtype=np.float32) - 0.5
        if dtype == "sb":
            return cpu.Oracle().tobf16(x)
        if dtype in ("c", "z"):
            return (x + 1j * (rng.random((n, n), dtype=np.float32) - 0.5)).astype(npdt)
        return x.astype(npdt)
    a, b = operand(), operand()
    c = np.zeros((n, n), dtype=cpu.NP_OUT[DT[dtype]])
    if kind == "reference":
        ref = cpu.Reference()
        ref.set_threads(cores)
        call = lambda: ref.gemm(DT[dtype], 0, 0, n, n, n, 1.0, a, n, b, n, 0.0, c, n)
        desc = f"oracle/_ref/{ref.target} ({ref.config()}), OPENBLAS_NUM_THREADS={cores}"
    else:
        orc = cpu.Oracle()
        cores = 1
        call = lambda: orc.gemm(DT[dtype], 0, 0, n, n, n, 1.0, a, n, b, n, 0.0, c, n)
        desc = "oracle/gemm_oracle.c (scalar port, 1 thread)"
    for _ in range(args.warmup):
        call()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        call()
    dt = time.perf_counter() - t0
    flops = 2.0 * n * n * n * args.steps      # BASELINE metric is 2mnk/t for every precision
    val = flops / dt / 1e12
    sample = f"{dtype.upper()}GEMM {n}^3 NN column-major alpha=1 beta=0 on host cores ({desc}); bounded sample of the {weak_shape(args.gpus)} workload"
    wm, wn, wk = weak_shape(args.gpus)
    cfg = workload_config(dtype, wm, wn, wk, "host CPU cores (reference arm)")
    cfg["reference_sample"] = f"each step = one {dtype}gemm {n}x{n}x{n} NN alpha=1 beta=0 on the host cores (bounded sample of the workload above)"
    print(json.dumps({
        "impl": "reference", "metric": f"{dtype.upper()}GEMM TFLOP/s (2mnk/t)", "value": val, "unit": "TFLOP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"d": "f64", "s": "f32", "z": "c128", "c": "c64", "sb": "bf16->f32"}[dtype],
        "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": val, "unit": "TFLOP/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------- GPU arm
def cpu_baseline(dtype, seconds_budget=25.0):
    """Reference OpenBLAS on this box's host cores, bounded sample, rank 0 / N = 1 only."""
    import numpy as np
    from oracle import cpu
    cores = os.cpu_count() or 1
    n = 8192 if dtype in ("d", "s", "sb") else 4096
    rng = np.random.default_rng(1)
    npdt = cpu.NP_IN[DT[dtype]]
    x = rng.random((n, n), dtype=np.float32) - 0.5
    if dtype == "sb":
        a = b = cpu.Oracle().tobf16(x)
    elif dtype in ("c", "z"):
        a = b = (x + 1j * x[::-1]).astype(npdt)
    else:
        a = b = x.astype(npdt)
    c = np.zeros((n, n), dtype=cpu.NP_OUT[DT[dtype]])
    if cpu.have_reference():
        ref = cpu.Reference()
        ref.set_threads(cores)
        call = lambda: ref.gemm(DT[dtype], 0, 0, n, n, n, 1.0, a, n, b, n, 0.0, c, n)
        kind, what = "reference", f"oracle/_ref/{ref.target} {ref.config()}"
    else:
        n = 512
        orc = cpu.Oracle()
        a, b, c = a[:n, :n].copy(), b[:n, :n].copy(), c[:n, :n].copy()
        call = lambda: orc.gemm(DT[dtype], 0, 0, n, n, n, 1.0, a, n, b, n, 0.0, c, n)
        kind, what, cores = "port", "oracle/gemm_oracle.c scalar port", 1
    call()
    best, spent, reps = 1e30, 0.0, 0
    while reps < 3 and spent < seconds_budget:
        t0 = time.perf_counter(); call(); dt = time.perf_counter() - t0
        best = min(best, dt); spent += dt; reps += 1
    return {"value": 2.0 * n ** 3 / best / 1e12, "unit": "TFLOP/s", "cores": cores, "kind": kind,
            "sample": f"{dtype.upper()}GEMM {n}^3 NN alpha=1 beta=0, best of {reps} calls, OPENBLAS_NUM_THREADS={cores}, {what}"}


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import openblas_b200 as ob
    from openblas_b200 import summa

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; openblas_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dtype = args.dtype
    code = DT[dtype]
    tdt = {"s": torch.float32, "d": torch.float64, "c": torch.complex64, "z": torch.complex128, "sb": torch.bfloat16}[dtype]
    odt = torch.float32 if dtype == "sb" else tdt
    m, n, k = (args.m, args.n, args.k) if args.m else weak_shape(world)
    lib = ob.lib()
    assert lib.b200_init(local_rank) == 0, lib.b200_last_error()
    gen = torch.Generator(device=dev); gen.manual_seed(1234 + rank)

    def rand(cols, rows, t):
        if t.is_complex:
            r = torch.rand((cols, rows, 2), generator=gen, device=dev, dtype=torch.float64 if t == torch.complex128 else torch.float32) - 0.5
            return torch.view_as_complex(r)
        return (torch.rand((cols, rows), generator=gen, device=dev, dtype=torch.float32 if t == torch.bfloat16 else t) - 0.5).to(t)

    grid = summa.make_grid(world, rank) if world > 1 else None
    launches0 = ob.cblas.launch_count()
    stream = torch.cuda.current_stream(dev)
    if world == 1:
        a, b = rand(k, m, tdt), rand(n, k, tdt)
        c = torch.empty((n, m), dtype=odt, device=dev)
        step = lambda: ob.cblas.gemm_device(code, 0, 0, m, n, k, 1.0, a, m, b, k, 0.0, c, m, stream.cuda_stream)
        parallelism, launches_per_step = "1 GPU", 1
    else:
        nb = args.nb
        sm = summa.Summa(grid, m, n, k, nb, tdt, dev,
                         lambda mm, nn, kk, al, A, lda, B, ldb, be, C, ldc, st: ob.cblas.gemm_device(code, 0, 0, mm, nn, kk, al, A, lda, B, ldb, be, C, ldc, st))
        a, b = rand(sm.ka_loc, sm.m_loc, tdt), rand(sm.n_loc, sm.kb_loc, tdt)
        c = torch.empty((sm.n_loc, sm.m_loc), dtype=odt, device=dev)
        step = lambda: sm.run(1.0, a, b, 0.0, c)
        parallelism, launches_per_step = f"summa {grid.P}x{grid.Q} block-cyclic nb={nb}", len(sm.steps)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    e0.record(stream)
    for i in range(args.steps):
        kern_ev[i][0].record(stream)
        step()
        kern_ev[i][1].record(stream)
    e1.record(stream)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    clocks = sampler.stop() if rank == 0 else None
    launches = ob.cblas.launch_count() - launches0
    flops_step = 2.0 * m * n * k        # BASELINE metric: 2mnk/t (complex: 4x that many real flops)
    value = flops_step * args.steps / (total_ms * 1e-3) / 1e12

    # dominant kernel = the local GEMM; at N = 1 a step IS one launch, so the per-launch duration is
    # the event pair around it
    kern_ms = sorted(s.elapsed_time(e) for s, e in kern_ev)
    kern_ms_avg = sum(kern_ms) / len(kern_ms)
    peak, peak_src = measured_peak(dtype)
    real_flops_launch = FLOP_FACTOR[dtype] * (m * n * k if world == 1 else 0)
    # N > 1 end to end: every rank's shards start in pinned HOST memory; a step = H2D of the local A and B
    # pieces, the SUMMA sweep, D2H of the local C piece (all ranks take part, so it runs before the
    # rank-0-only reporting)
    e2e_multi = None
    if world > 1 and not args.no_e2e:
        ha, hb = a.cpu().pin_memory(), b.cpu().pin_memory()
        hc = torch.empty(c.shape, dtype=c.dtype).pin_memory()
        da, db, dc = torch.empty_like(a), torch.empty_like(b), torch.empty_like(c)

        def e2e_step():
            da.copy_(ha, non_blocking=True); db.copy_(hb, non_blocking=True)
            sm.run(1.0, da, db, 0.0, dc)
            hc.copy_(dc, non_blocking=True)
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        nst = max(1, min(args.steps, args.e2e_steps))
        for _ in range(nst):
            e2e_step()
        barrier()
        dt_e = torch.tensor([(time.perf_counter() - t0) / nst], device=dev, dtype=torch.float64)
        dist.all_reduce(dt_e, op=dist.ReduceOp.MAX)
        bytes_in = torch.tensor([float(ha.numel() * ha.element_size() + hb.numel() * hb.element_size())], device=dev, dtype=torch.float64)
        bytes_out = torch.tensor([float(hc.numel() * hc.element_size())], device=dev, dtype=torch.float64)
        dist.all_reduce(bytes_in); dist.all_reduce(bytes_out)
        e2e_multi = {"value": flops_step / float(dt_e.item()) / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": int(bytes_in.item()),
                     "d2h_bytes_per_step": int(bytes_out.item()), "ms_per_step": float(dt_e.item()) * 1e3, "steps": nst,
                     "api": "per rank: pinned host shards -> device, openblas_b200.summa.Summa.run (local product b200_gemm_async), C shard -> pinned host",
                     "result_checksum": float(hc[::97, ::89].double().sum())}
    out = None
    if rank == 0:
        e2e = None
        if world == 1 and not args.no_e2e:
            e2e = run_e2e(ob, torch, code, dtype, m, n, k, tdt, odt, args)
        elif world > 1:
            e2e = e2e_multi
        cpu_b = cpu_baseline(dtype) if (world == 1 and not args.no_cpu) else None
        achieved = (real_flops_launch / (kern_ms_avg * 1e-3) / 1e12) if world == 1 else (FLOP_FACTOR[dtype] * m * n * k / world / (total_ms / args.steps * 1e-3) / 1e12)
        out = {
            "metric": f"{dtype.upper()}GEMM TFLOP/s (2mnk/t)", "value": value, "unit": "TFLOP/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"d": "f64", "s": "f32", "z": "c128", "c": "c64", "sb": "bf16->f32"}[dtype], "data": "synthetic",
            "config": workload_config(dtype, m, n, k, parallelism),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": NCU_TRAFFIC_BYTES.get((dtype, m, n, k)) if world == 1 else None, "traffic_unit": "bytes/launch (ncu dram read+write)",
                         "algorithmic_bytes": (2 if dtype == "sb" else {"s": 4, "d": 8, "c": 8, "z": 16}[dtype]) * (m * k + k * n) + {"s": 4, "d": 8, "c": 8, "z": 16, "sb": 4}[dtype] * m * n,
                         "peak_source": peak_src,
                         "kernel": ob.cblas.last_kernel(), "kernel_ms_avg": kern_ms_avg if world == 1 else None,
                         "note": "achieved = real flops of one launch (2mnk, 8mnk complex) / CUDA-event duration of that launch; per GPU at N>1"},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if cpu_b:
            out["cpu_baseline"] = cpu_b
        if e2e:
            out["e2e"] = e2e
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if out:
        print(json.dumps(out))


def run_e2e(ob, torch, code, dtype, m, n, k, tdt, odt, args):
    """The call a user of the reference makes: cblas_?gemm with HOST pointers (pinned), so the
    timed region holds the H2D copies of A and B, the kernel and the D2H copy of C."""
    a = torch.empty((k, m), dtype=tdt).pin_memory()
    b = torch.empty((n, k), dtype=tdt).pin_memory()
    c = torch.empty((n, m), dtype=odt).pin_memory()
    for t in (a, b):
        if t.is_complex():
            torch.view_as_real(t).uniform_(-0.5, 0.5)
        elif t.dtype == torch.bfloat16:
            t.copy_(torch.rand(t.shape) - 0.5)
        else:
            t.uniform_(-0.5, 0.5)
    fn = ob.cblas.GEMM[code]
    steps = max(1, min(args.steps, args.e2e_steps))
    call = lambda: fn(ob.cblas.ColMajor, ob.cblas.NoTrans, ob.cblas.NoTrans, m, n, k, 1.0, a, m, b, k, 0.0, c, m)
    call()   # warm-up: grows the device workspace once
    t0 = time.perf_counter()
    for _ in range(steps):
        call()
    dt = (time.perf_counter() - t0) / steps
    checksum = float(torch.view_as_real(c).double().sum() if c.is_complex() else c[::257, ::263].double().sum())
    return {"value": 2.0 * m * n * k / dt / 1e12, "unit": "TFLOP/s",
            "h2d_bytes_per_step": a.numel() * a.element_size() + b.numel() * b.element_size(),
            "d2h_bytes_per_step": c.numel() * c.element_size(), "ms_per_step": dt * 1e3, "steps": steps,
            "api": "cblas_%sgemm(CblasColMajor, CblasNoTrans, CblasNoTrans, ...) on pinned host buffers; call is synchronous" % dtype,
            "result_checksum": checksum}


def run_sweep(args):
    """Developer view (not the driver contract): TFLOP/s of every precision over sizes and ops."""
    import torch
    import openblas_b200 as ob
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    stream = torch.cuda.current_stream(dev)
    rows = []
    sizes = [int(x) for x in args.sizes.split(",")]
    for dtype in args.sweep_dtypes.split(","):
        code = DT[dtype]
        tdt = {"s": torch.float32, "d": torch.float64, "c": torch.complex64, "z": torch.complex128, "sb": torch.bfloat16}[dtype]
        odt = torch.float32 if dtype == "sb" else tdt
        peak, _ = measured_peak(dtype)
        for nsz in sizes:
            if dtype in ("z", "c") and nsz > 8192:
                continue
            ops = [(0, 0)] if (nsz > 4096 and not args.all_ops) else [(0, 0), (1, 0), (0, 1), (1, 1)]
            if args.all_ops and dtype in ("z", "c"):       # BASELINE config 5: every N/T/C combination
                ops = [(x, y) for y in (0, 1, 3) for x in (0, 1, 3)]
            for ta, tb in ops:
                m = n = k = nsz
                mk = lambda: (torch.view_as_complex(torch.rand((nsz, nsz, 2), device=dev, dtype=torch.float64 if dtype == "z" else torch.float32) - 0.5)
                              if tdt.is_complex else (torch.rand((nsz, nsz), device=dev, dtype=torch.float32 if dtype == "sb" else tdt) - 0.5).to(tdt))
                a, b = mk(), mk()
                c = torch.zeros((nsz, nsz), dtype=odt, device=dev)
                f = lambda: ob.cblas.gemm_device(code, ta, tb, m, n, k, 1.0, a, nsz, b, nsz, 0.0, c, nsz, stream.cuda_stream)
                for _ in range(3):
                    f()
                torch.cuda.synchronize()
                reps = 3 if nsz >= 8192 else 10
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); [f() for _ in range(reps)]; e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                tf = FLOP_FACTOR[dtype] * nsz ** 3 / (ms * 1e-3) / 1e12
                rows.append({"dtype": dtype, "n": nsz, "op": "NTRC"[ta] + "NTRC"[tb], "ms": ms, "tflops_real": tf, "frac_of_peak": tf / peak,
                             "kernel": ob.cblas.last_kernel()})
                print(json.dumps(rows[-1]), flush=True)
                del a, b, c
    return rows


def run_sweep_level3(args):
    """Developer view: the symmetric level-3 family on device-resident operands, TFLOP/s with the
    conventional flop counts (SYMM 2*m*m*n, SYRK n*n*k, SYR2K 2*n*n*k, TRMM/TRSM m*m*n real; complex x4) -- what a GEMM
    of the same useful work would be credited with."""
    import ctypes as C
    import torch
    import openblas_b200 as ob
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    lib = ob.lib()
    rows = []
    i_ = lambda v: C.byref(C.c_int(int(v)))
    for dtype in [d for d in args.sweep_dtypes.split(",") if d in "sdcz"]:
        cplx = dtype in "cz"
        rdt = torch.float64 if dtype in "dz" else torch.float32
        ct = C.c_double if dtype in "dz" else C.c_float
        peak, _ = measured_peak(dtype)
        for nsz in [int(x) for x in args.sizes.split(",")]:
            n = k = nsz
            mk = lambda: (torch.rand((nsz, nsz, 2) if cplx else (nsz, nsz), device=dev, dtype=rdt) - 0.5)
            a, b, c = mk(), mk(), mk()
            a.mul_(1.0 / nsz)                                     # keeps the unit-diagonal TRSM sweeps bounded
            al2, be2 = (ct * 2)(0.7, 0.2), (ct * 2)(1.3, 0.1)
            alr, ber = ct(0.7), ct(1.3)
            pa, pb, pc = C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(c.data_ptr())
            jobs = [("symm", lambda: getattr(lib, dtype + "symm_")(C.c_char_p(b"L"), C.c_char_p(b"U"), i_(n), i_(n), al2, pa, i_(n), pb, i_(n), be2, pc, i_(n)), 2.0),
                    ("syrk", lambda: getattr(lib, dtype + "syrk_")(C.c_char_p(b"L"), C.c_char_p(b"N"), i_(n), i_(k), al2, pa, i_(n), be2, pc, i_(n)), 1.0),
                    ("syr2k", lambda: getattr(lib, dtype + "syr2k_")(C.c_char_p(b"U"), C.c_char_p(b"T"), i_(n), i_(k), al2, pa, i_(n), pb, i_(n), be2, pc, i_(n)), 2.0)]
            jobs += [("trmm", lambda: getattr(lib, dtype + "trmm_")(C.c_char_p(b"L"), C.c_char_p(b"L"), C.c_char_p(b"N"), C.c_char_p(b"N"), i_(n), i_(n), al2, pa, i_(n), pc, i_(n)), 1.0),
                     ("trsm", lambda: getattr(lib, dtype + "trsm_")(C.c_char_p(b"L"), C.c_char_p(b"L"), C.c_char_p(b"N"), C.c_char_p(b"U"), i_(n), i_(n), al2, pa, i_(n), pc, i_(n)), 1.0),
                     ("trsm_right_upper_T", lambda: getattr(lib, dtype + "trsm_")(C.c_char_p(b"R"), C.c_char_p(b"U"), C.c_char_p(b"T"), C.c_char_p(b"U"), i_(n), i_(n), al2, pa, i_(n), pc, i_(n)), 1.0)]
            if cplx:
                jobs += [("hemm", lambda: getattr(lib, dtype + "hemm_")(C.c_char_p(b"R"), C.c_char_p(b"L"), i_(n), i_(n), al2, pa, i_(n), pb, i_(n), be2, pc, i_(n)), 2.0),
                         ("herk", lambda: getattr(lib, dtype + "herk_")(C.c_char_p(b"U"), C.c_char_p(b"C"), i_(n), i_(k), C.byref(alr), pa, i_(n), C.byref(ber), pc, i_(n)), 1.0)]
            for name, f, factor in jobs:
                f(); torch.cuda.synchronize()
                reps = 3
                t0 = time.perf_counter()                      # the BLAS call is synchronous: wall time of the call itself
                for _ in range(reps):
                    f()
                ms = (time.perf_counter() - t0) / reps * 1e3
                tf = factor * (4.0 if cplx else 1.0) * nsz ** 3 / (ms * 1e-3) / 1e12
                rows.append({"routine": dtype + name, "n": nsz, "k": nsz, "ms": ms, "tflops_useful": tf, "frac_of_gemm_peak": tf / peak,
                             "launches_per_call": None})
                l0 = ob.cblas.launch_count(); f(); rows[-1]["launches_per_call"] = int(ob.cblas.launch_count() - l0)
                print(json.dumps(rows[-1]), flush=True)
            del a, b, c
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dtype", default="d", choices=list(DT))
    ap.add_argument("--m", type=int, default=0); ap.add_argument("--n", type=int, default=0); ap.add_argument("--k", type=int, default=0)
    ap.add_argument("--nb", type=int, default=2048, help="SUMMA distribution block / panel width")
    ap.add_argument("--ref-n", type=int, default=8192, help="size of the bounded CPU sample of the reference arm")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true"); ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--sweep", action="store_true"); ap.add_argument("--sizes", default="1024,2048,4096,8192,16384")
    ap.add_argument("--sweep-level3", action="store_true", help="developer view: SYMM/SYRK/SYR2K/HEMM/HERK on device operands")
    ap.add_argument("--sweep-dtypes", default="d,s,z,c,sb")
    ap.add_argument("--all-ops", action="store_true", help="sweep: all four N/T combinations at every size")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.sweep_level3:
        return run_sweep_level3(args)
    if args.sweep:
        return run_sweep(args)
    return run_gpu(args)


if __name__ == "__main__":
    main()
