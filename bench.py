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
# DRAM traffic of the dominant kernel per launch (dram__bytes_read.sum + dram__bytes_write.sum): read from the committed
# summary of one `ncu --set full` capture of the same kernel and shape (profiles/r02_dgemm_16384_ncu.json, written by
# tools/r02_call14.sh; a number measured under a profiler is evidence about the kernel, never a bench value)
def ncu_traffic(dtype, m, n, k):
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "r02_dgemm_16384_ncu.json")))
        if (rec["dtype"], rec["m"], rec["n"], rec["k"]) == (dtype, m, n, k):
            return float(rec["dram_bytes_read"]) + float(rec["dram_bytes_write"]), rec.get("source", "profiles/r02_dgemm_16384_ncu.json")
    except Exception:
        pass
    if (dtype, m, n, k) == ("d", 16384, 16384, 16384):
        return 54.017287e9 + 2.149455e9, "profiles/r01_prof_d_16384_final2_summary.txt (round-1 capture of the same kernel)"
    return None, None


_INRUN_PEAKS = None


def inrun_peaks(device=0):
    """FP64 / FP32 dense peaks measured in THIS lease: tools/peaks (built by __graft_entry__.build) runs the
    DMMA / FFMA / FFMA2 issue-rate probes and cuBLAS D/SGEMM (pedantic: no TF32) at 8192^3 on the same GPU,
    before the benchmark touches it.  None when the tool is missing or fails (the committed round-1
    measurement is used then, and the line says so)."""
    global _INRUN_PEAKS
    if _INRUN_PEAKS is not None:
        return _INRUN_PEAKS or None
    _INRUN_PEAKS = {}
    exe = os.path.join(ROOT, "tools", "peaks")
    if os.path.exists(exe):
        try:
            r = subprocess.run([exe, "quick", str(device)], capture_output=True, text=True, timeout=120)
            if r.returncode == 0:
                _INRUN_PEAKS = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception:
            _INRUN_PEAKS = {}
    return _INRUN_PEAKS or None


def measured_peak(dtype, device=0):
    """(peak TFLOP/s, where it comes from).  bf16: MEASURED_PEAKS.json (driver-written, burst).  FP64 / FP32:
    MEASURED_PEAKS.json has no entry and B200_PROFILING.md states no fallback, so the denominator is the
    issue rate of the pipe's own instruction measured in this run (DMMA.8x8x4 / FFMA), else the same probe's
    committed round-1 result."""
    if dtype == "sb":
        try:
            p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            return float(p["bf16_tflops"]), "MEASURED_PEAKS.json bf16_tflops (burst)"
        except Exception:
            return PEAK_FALLBACK["sb"], "fallback 1.59 PFLOP/s (B200_PROFILING.md)"
    key = "dmma_tflops" if dtype in ("d", "z") else "ffma_tflops"
    what = "DMMA.8x8x4" if dtype in ("d", "z") else "FFMA"
    p = inrun_peaks(device)
    if p:
        return max(v for k, v in p.items() if k.startswith(key)), f"{what} issue rate measured in this run by tools/peaks (MEASURED_PEAKS.json has no FP64/FP32 entry)"
    try:
        p = json.load(open(os.path.join(ROOT, "profiles", "r01_peaks_microbench.json")))
        return max(v for k, v in p.items() if k.startswith(key)), f"{what} issue rate, tools/peaks result committed in round 1 (profiles/r01_peaks_microbench.json); the in-run probe was unavailable"
    except Exception:
        return PEAK_FALLBACK[dtype], "tools/peaks value recorded in bench.py"


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


def grid_of(world):
    return {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (2, 4), 16: (4, 4)}.get(world, (1, world))


def workload_config(dtype, m, n, k, world, nb):
    """The `config` object BOTH arms print, key for key (the driver compares them): it names the workload,
    not the arm.  What ran it is in `arm`."""
    P, Q = grid_of(world)
    return {"workload": f"{dtype}gemm NN column-major {m}x{n}x{k} alpha=1 beta=0 (BASELINE configs[1] at N=1; configs[3] shape at N=8)",
            "parallelism": "one GEMM" if world == 1 else f"one GEMM, C distributed 2-D block-cyclically over a {P}x{Q} grid (nb={nb}) on the GPU arm",
            "l2_policy": "inputs larger than L2 (A,B >= 2 GiB each vs 126 MB L2)",
            "inputs": "uniform(-0.5,0.5), fixed seed"}


def weak_shape(world):
    return {1: (16384, 16384, 16384), 2: (16384, 32768, 16384), 4: (32768, 32768, 16384), 8: (32768, 32768, 32768)}.get(
        world, (16384, 16384 * world, 16384))


# ---------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's own CPU GEMM on the box's host cores, all threads, on the SAME workload as the GPU arm
    whenever (steps + warmup) calls of it fit in --ref-budget seconds (DGEMM 16384^3 is about 6 s per call
    on 16 Sapphire Rapids cores); otherwise every dimension is halved until they do, and the line says so
    (`same_workload` false, `reference_sample` in config)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle import cpu
    kind = "reference" if cpu.have_reference() else "port"
    dtype = args.dtype
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(0)
    npdt = cpu.NP_IN[DT[dtype]]
    wm, wn, wk = (args.m, args.n, args.k) if args.m else weak_shape(args.gpus)

    def operand(rows, cols):
        x = rng.random((cols, rows), dtype=np.float32) - 0.5
        if dtype == "sb":
            return cpu.Oracle().tobf16(x)
        if dtype in ("c", "z"):
            return (x + 1j * (rng.random((cols, rows), dtype=np.float32) - 0.5)).astype(npdt)
        return x.astype(npdt)
    if kind == "reference":
        ref = cpu.Reference()
        ref.set_threads(cores)
        gemm = lambda m, n, k, a, b, c: ref.gemm(DT[dtype], 0, 0, m, n, k, 1.0, a, m, b, k, 0.0, c, m)
        desc = f"oracle/_ref/{ref.target} ({ref.config()}), OPENBLAS_NUM_THREADS={cores}"
    else:
        orc = cpu.Oracle()
        cores = 1
        gemm = lambda m, n, k, a, b, c: orc.gemm(DT[dtype], 0, 0, m, n, k, 1.0, a, m, b, k, 0.0, c, m)
        desc = "oracle/gemm_oracle.c (scalar port, 1 thread)"
    # calibration call (untimed for the result): rate of a small GEMM -> how much of the workload fits the budget
    cn = 2048 if kind == "reference" else 256
    ca, cb_, cc = operand(cn, cn), operand(cn, cn), np.zeros((cn, cn), dtype=cpu.NP_OUT[DT[dtype]])
    gemm(cn, cn, cn, ca, cb_, cc)
    t0 = time.perf_counter(); gemm(cn, cn, cn, ca, cb_, cc); rate = 2.0 * cn ** 3 / (time.perf_counter() - t0)
    del ca, cb_, cc
    m, n, k = wm, wn, wk
    if args.ref_n:
        m = n = k = args.ref_n
    else:
        calls = args.steps + args.warmup
        while 2.0 * m * n * k / rate * calls > args.ref_budget and min(m, n, k) > 256:
            m, n, k = max(256, m // 2), max(256, n // 2), max(256, k // 2)
    same = (m, n, k) == (wm, wn, wk)
    a, b = operand(m, k), operand(k, n)
    c = np.zeros((n, m), dtype=cpu.NP_OUT[DT[dtype]])
    call = lambda: gemm(m, n, k, a, b, c)
    for _ in range(args.warmup):
        call()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        call()
    dt = time.perf_counter() - t0
    flops = 2.0 * m * n * k * args.steps      # BASELINE metric is 2mnk/t for every precision
    val = flops / dt / 1e12
    sample = (f"{dtype.upper()}GEMM {m}x{n}x{k} NN column-major alpha=1 beta=0 on host cores ({desc}); "
              + ("the full workload of the GPU arm" if same else f"bounded sample of the {wm}x{wn}x{wk} workload (rates, so comparable; {args.steps + args.warmup} full-size calls would exceed {args.ref_budget:.0f} s)"))
    cfg = workload_config(dtype, wm, wn, wk, args.gpus, args.nb)
    if not same:
        cfg["reference_sample"] = f"each step = one {dtype}gemm {m}x{n}x{k} NN alpha=1 beta=0 on the host cores (bounded sample of the workload above)"
    print(json.dumps({
        "impl": "reference", "metric": f"{dtype.upper()}GEMM TFLOP/s (2mnk/t)", "value": val, "unit": "TFLOP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"d": "f64", "s": "f32", "z": "c128", "c": "c64", "sb": "bf16->f32"}[dtype],
        "data": "synthetic",
        "config": cfg, "same_workload": same,
        "arm": f"host CPU cores: {desc}",
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
    if world == 1 and not args.no_peaks:
        inrun_peaks(local_rank)          # FP64 / FP32 denominators, measured on the idle GPU before the benchmark
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

    launches0 = ob.cblas.launch_count()
    stream = torch.cuda.current_stream(dev)
    cs = sm = None
    ri = cj = None
    if world == 1:
        a, b = rand(k, m, tdt), rand(n, k, tdt)
        c = torch.empty((n, m), dtype=odt, device=dev)
        step = lambda: ob.cblas.gemm_device(code, 0, 0, m, n, k, 1.0, a, m, b, k, 0.0, c, m, stream.cuda_stream)
        parallelism, launches_per_step = "1 GPU", 1
    else:
        nb = args.nb
        P, Q = summa.grid_shape(world)
        p_, q_ = rank // Q, rank % Q
        # hashed global matrices: every rank can regenerate any row of A / column of B, which is what the sampled
        # long-double check of its C block needs (no communication, nothing of it inside the timed region)
        idx = lambda nn, i, np_: torch.tensor(summa.local_index_map(nn, nb, i, np_), dtype=torch.int64, device=dev)
        ri, cj, ka_i, kb_i = idx(m, p_, P), idx(n, q_, Q), idx(k, q_, Q), idx(k, p_, P)

        def hashed(tag, gi, gj):
            if tdt.is_complex:
                return torch.complex(summa.hashed_entries(tag, gi, gj), summa.hashed_entries(tag + 10, gi, gj)).to(tdt)
            return summa.hashed_entries(tag, gi, gj).to(tdt)
        a, b = hashed(1, ri, ka_i), hashed(2, kb_i, cj)
        m_loc, n_loc, ka_loc, kb_loc = ri.numel(), cj.numel(), ka_i.numel(), kb_i.numel()
        c = torch.empty((n_loc, m_loc), dtype=odt, device=dev)
        torch.cuda.empty_cache()
        if args.summa == "c":
            cs = summa.CSumma(world, rank, dev)
            step = lambda: cs.gemm(code, m, n, k, nb, 1.0, a, max(1, m_loc), b, max(1, kb_loc), 0.0, c, max(1, m_loc), stream.cuda_stream)
            parallelism = f"{cs.describe()}; nb={nb}; b200_summa_gemm (csrc/summa.cu), one process per GPU"
        else:
            grid = summa.make_grid(world, rank)
            sm = summa.Summa(grid, m, n, k, nb, tdt, dev,
                             lambda mm, nn, kk, al, A, lda, B, ldb, be, C, ldc, st: ob.cblas.gemm_device(code, 0, 0, mm, nn, kk, al, A, lda, B, ldb, be, C, ldc, st))
            step = lambda: sm.run(1.0, a, b, 0.0, c)
            parallelism = f"summa.py {P}x{Q} block-cyclic nb={nb}: NCCL panel broadcasts (torch.distributed) + b200_gemm_async"
        launches_per_step = (k + nb - 1) // nb

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
    launches0 = ob.cblas.launch_count()              # kernels enqueued inside the timed region only
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
    main_kernel = ob.cblas.last_kernel()
    kern_ms = sorted(s.elapsed_time(e) for s, e in kern_ev)
    kern_ms_avg = sum(kern_ms) / len(kern_ms)
    peak, peak_src = measured_peak(dtype, local_rank)
    real_flops_launch = FLOP_FACTOR[dtype] * (m * n * k if world == 1 else 0)
    # N > 1 end to end: every rank's shards start in pinned HOST memory; a step = H2D of the local A and B
    # pieces, the SUMMA sweep, D2H of the local C piece (all ranks take part, so it runs before the
    # rank-0-only reporting)
    e2e_multi = None
    parity_multi = None
    if world > 1:
        # sampled long-double check of this rank's C block (outside every timed region), worst ratio over all ranks
        import numpy as np
        worst, samples = 0.0, 0
        if ri.numel() and cj.numel():
            gsel = torch.Generator(); gsel.manual_seed(100 + rank)
            allk = torch.arange(k, dtype=torch.int64, device=dev)
            eps = 2.0 ** -52 if dtype in ("d", "z") else 2.0 ** -23
            ldt = np.clongdouble if tdt.is_complex else np.longdouble
            for _ in range(16):
                il, jl = int(torch.randint(0, ri.numel(), (1,), generator=gsel)), int(torch.randint(0, cj.numel(), (1,), generator=gsel))
                ar = hashed(1, ri[il:il + 1], allk)[:, 0].cpu().numpy().astype(ldt)
                bc = hashed(2, allk, cj[jl:jl + 1])[0, :].cpu().numpy().astype(ldt)
                ref = np.dot(ar, bc)
                gauge = float(np.dot(np.abs(ar), np.abs(bc)))
                worst = max(worst, float(abs(ldt(c[jl, il].item()) - ref)) / (k * eps * gauge))
                samples += 1
        w = torch.tensor([worst], device=dev, dtype=torch.float64)
        ns = torch.tensor([float(samples)], device=dev, dtype=torch.float64)
        dist.all_reduce(w, op=dist.ReduceOp.MAX); dist.all_reduce(ns)
        parity_multi = {"worst_ratio": float(w.item()), "samples": int(ns.item()), "ranks": world,
                        "bound": "|C - C_ref| <= 2 * k * eps * (|A||B|) per entry; C_ref = long-double dot products of regenerated global rows / columns, 16 entries of every rank's block",
                        "ok": float(w.item()) <= 2.0}
    if world > 1 and not args.no_e2e:
        # end to end: every rank's shards start in pinned HOST memory and C ends there.  The library's driver takes the
        # host pointers itself: uploads go straight into the panel windows in k order, the products start as soon as
        # the first panels have landed everywhere, the last update of C comes back strip by strip.
        ha, hb = a.cpu().pin_memory(), b.cpu().pin_memory()
        hc = torch.empty(c.shape, dtype=c.dtype).pin_memory()
        if cs is not None:
            e2e_step = lambda: cs.gemm(code, m, n, k, args.nb, 1.0, ha, max(1, m_loc), hb, max(1, kb_loc), 0.0, hc, max(1, m_loc), stream.cuda_stream)
            api = "per rank: b200_summa_gemm on pinned HOST shards of A, B and C (synchronous)"
        else:
            da, db, dc = torch.empty_like(a), torch.empty_like(b), torch.empty_like(c)

            def e2e_step():
                da.copy_(ha, non_blocking=True); db.copy_(hb, non_blocking=True)
                sm.run(1.0, da, db, 0.0, dc)
                hc.copy_(dc, non_blocking=True)
            api = "per rank: pinned host shards -> device, openblas_b200.summa.Summa.run (local product b200_gemm_async), C shard -> pinned host"
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
        diff = float((hc.to(dev) - c).abs().max()) if c.numel() else 0.0       # the host-path result equals the device-path result
        e2e_multi = {"value": flops_step / float(dt_e.item()) / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": int(bytes_in.item()),
                     "d2h_bytes_per_step": int(bytes_out.item()), "ms_per_step": float(dt_e.item()) * 1e3, "steps": nst, "api": api,
                     "max_abs_diff_vs_device_path_rank0": diff, "result_checksum": float(hc[::97, ::89].double().sum())}
    out = None
    if rank == 0:
        e2e = None
        if world == 1 and not args.no_e2e:
            e2e = run_e2e(ob, torch, code, dtype, m, n, k, tdt, odt, args)
        elif world > 1:
            e2e = e2e_multi
        cpu_b = cpu_baseline(dtype) if (world == 1 and not args.no_cpu) else None
        parity = sampled_parity(torch, dtype, 0, 0, m, n, k, a, m, b, k, c, m) if world == 1 else parity_multi
        extra = run_extras(ob, torch, dev, stream, args) if (world == 1 and not args.no_extras) else None
        achieved = (real_flops_launch / (kern_ms_avg * 1e-3) / 1e12) if world == 1 else (FLOP_FACTOR[dtype] * m * n * k / world / (total_ms / args.steps * 1e-3) / 1e12)
        out = {
            "metric": f"{dtype.upper()}GEMM TFLOP/s (2mnk/t)", "value": value, "unit": "TFLOP/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"d": "f64", "s": "f32", "z": "c128", "c": "c64", "sb": "bf16->f32"}[dtype], "data": "synthetic",
            "config": workload_config(dtype, m, n, k, world, args.nb), "arm": parallelism,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(dtype, m, n, k)[0] if world == 1 else None, "traffic_unit": "bytes/launch (ncu dram read+write)",
                         "traffic_source": ncu_traffic(dtype, m, n, k)[1] if world == 1 else None,
                         "algorithmic_bytes": (2 if dtype == "sb" else {"s": 4, "d": 8, "c": 8, "z": 16}[dtype]) * (m * k + k * n) + {"s": 4, "d": 8, "c": 8, "z": 16, "sb": 4}[dtype] * m * n,
                         "peak_source": peak_src,
                         "kernel": main_kernel, "kernel_ms_avg": kern_ms_avg if world == 1 else None,
                         "note": "achieved = real flops of one launch (2mnk, 8mnk complex) / CUDA-event duration of that launch; per GPU at N>1"},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if parity:
            out["parity"] = parity
        if world == 1:
            out["peaks_in_run"] = inrun_peaks(local_rank)
        if extra:
            out["extra"] = extra
        if cpu_b:
            out["cpu_baseline"] = cpu_b
        if e2e:
            out["e2e"] = e2e
    if world > 1:
        dist.barrier()
        if cs is not None:
            cs.close()
        dist.destroy_process_group()
    if out:
        print(json.dumps(out))


def sampled_parity(torch, dtype, ta, tb, m, n, k, a, lda, b, ldb, c, ldc, samples=64, alpha=1.0, seed=7, beta=0.0, c0=None):
    """`samples` entries of C = alpha * op(A) op(B) + beta * C0 recomputed on the host in long double from the
    operands as they lie in HBM, outside every timed region.  Returns the worst ratio
    |C - C_ref| / (k * eps * (|alpha| sum|a||b| + |beta||c0|)): the north-star acceptance bound is ratio <= c with
    c = 2.  Column-major operands held as 2-D torch tensors (cols, ld); c0 = the C the call started from (beta != 0)."""
    import numpy as np
    g = torch.Generator(); g.manual_seed(seed)
    ii = torch.randint(0, m, (samples,), generator=g).tolist()
    jj = torch.randint(0, n, (samples,), generator=g).tolist()
    eps = 2.0 ** -52 if dtype in ("d", "z") else 2.0 ** -23
    ld_t = np.clongdouble if dtype in ("c", "z") else np.longdouble

    def host(t):
        t = t.cpu()
        if t.dtype == torch.bfloat16:
            t = t.float()
        return t.numpy().astype(ld_t)
    worst = 0.0
    for i, j in zip(ii, jj):
        if ta & 1:
            ar = host(a[i, :k])                 # op(A)(i, :) = column i of the stored k x m matrix
        else:
            ar = host(a[:k, i])                 # row i of the stored m x k matrix
        if tb & 1:
            bc = host(b[:k, j])
        else:
            bc = host(b[j, :k])
        if ta & 2:
            ar = np.conj(ar)
        if tb & 2:
            bc = np.conj(bc)
        ref = alpha * np.dot(ar, bc)
        gauge = float(abs(alpha) * np.dot(np.abs(ar), np.abs(bc)))
        if beta != 0.0:
            old = ld_t(c0[j, i].item())
            ref = ref + ld_t(beta) * old
            gauge += float(abs(beta) * abs(old))
        got = c[j, i].item()
        worst = max(worst, float(abs(ld_t(got) - ref)) / (k * eps * gauge))
    return {"worst_ratio": worst, "samples": samples, "bound": "|C - C_ref| <= 2 * k * eps * (|alpha||A||B| + |beta||C|) per entry, C_ref in long double on the host",
            "ok": worst <= 2.0}


def device_time_ms(torch, fn, reps, stream):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


def run_extras(ob, torch, dev, stream, args):
    """The other north-star targets, measured in the same driver run as the headline (device-resident operands,
    CUDA events on the launching stream, every result spot-checked against long-double dot products):
    SBGEMM 8192^3 on the default kernel (burst and >= 2 s sustained), SGEMM 16384^3, ZGEMM / CGEMM 8192^3 (NN and
    the other op combinations), the tall-skinny shapes of BASELINE config 5, DGEMM/SGEMM mid sizes of config 2."""
    T = {"s": torch.float32, "d": torch.float64, "c": torch.complex64, "z": torch.complex128, "sb": torch.bfloat16}
    ES = {"s": 4, "d": 8, "c": 8, "z": 16, "sb": 2}
    gen = torch.Generator(device=dev); gen.manual_seed(4321)
    hbm = None
    try:
        hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass

    def rand(cols, rows, dtype):
        t = T[dtype]
        if t.is_complex:                 # in place: the tall-skinny operands are 64 GiB each
            return torch.view_as_complex(torch.rand((cols, rows, 2), generator=gen, device=dev, dtype=torch.float64 if dtype == "z" else torch.float32).sub_(0.5))
        if dtype == "sb":
            return torch.rand((cols, rows), generator=gen, device=dev, dtype=torch.float32).sub_(0.5).to(t)
        return torch.rand((cols, rows), generator=gen, device=dev, dtype=t).sub_(0.5)

    def one(dtype, ops, m, n, k, reps, check=True, alpha=1.0, beta=0.0):
        """every op combination in `ops` on an m x n x k problem; returns {op: TFLOP/s (real flops)}, kernel, parity.
        beta != 0: the timed calls keep updating C in place; the checked call starts again from a saved C0 (when C is
        too big to keep a second copy -- the 64 GiB C of the tall-skinny shape -- the check is a beta = 0 call)."""
        code = DT[dtype]
        res, worst, kern = {}, 0.0, None
        odt = torch.float32 if dtype == "sb" else T[dtype]
        keep_c0 = beta != 0.0 and n * m * ES[dtype] <= (8 << 30)
        c = rand(n, m, dtype) if (beta != 0.0 and dtype != "sb") else torch.empty((n, m), dtype=odt, device=dev)
        c0 = c.clone() if keep_c0 else None
        for ta, tb in ops:
            ra, ca = (k, m) if ta & 1 else (m, k)
            rb, cb = (n, k) if tb & 1 else (k, n)
            a, b = rand(ca, ra, dtype), rand(cb, rb, dtype)
            f = lambda al=alpha, be=beta: ob.cblas.gemm_device(code, ta, tb, m, n, k, al, a, ra, b, rb, be, c, m, stream.cuda_stream)
            for _ in range(3):
                f()
            torch.cuda.synchronize(dev)
            ms = device_time_ms(torch, f, reps, stream)
            res["NTRC"[ta] + "NTRC"[tb]] = FLOP_FACTOR[dtype] * m * n * k / (ms * 1e-3) / 1e12
            kern = ob.cblas.last_kernel()
            if check:
                if keep_c0:
                    c.copy_(c0); f(); torch.cuda.synchronize(dev)
                    w = sampled_parity(torch, dtype, ta, tb, m, n, k, a, ra, b, rb, c, m, samples=16, alpha=alpha, beta=beta, c0=c0)
                else:
                    if beta != 0.0:
                        f(alpha, 0.0); torch.cuda.synchronize(dev)
                    w = sampled_parity(torch, dtype, ta, tb, m, n, k, a, ra, b, rb, c, m, samples=16, alpha=alpha)
                worst = max(worst, w["worst_ratio"])
            del a, b
        del c, c0
        return res, kern, worst

    out = {}
    NT4 = [(0, 0), (1, 0), (0, 1), (1, 1)]
    # (a) SBGEMM 8192^3, default kernel: burst = 20 back-to-back launches after warm-up; sustained = back to back for >= 2 s
    m = n = k = 8192
    a, b = rand(k, m, "sb"), rand(n, k, "sb")
    c = torch.empty((n, m), dtype=torch.float32, device=dev)
    f = lambda: ob.cblas.gemm_device(DT["sb"], 0, 0, m, n, k, 1.0, a, m, b, k, 0.0, c, m, stream.cuda_stream)
    for _ in range(5):
        f()
    torch.cuda.synchronize(dev)
    burst_ms = min(device_time_ms(torch, f, 20, stream) for _ in range(3))
    parity_nn = sampled_parity(torch, "sb", 0, 0, m, n, k, a, m, b, k, c, m, samples=16)["worst_ratio"]
    # the other op combinations BEFORE the sustained loop (which leaves the chip power-capped for a while)
    r_ops, _, w_ops = one("sb", [(1, 0), (0, 1), (1, 1)], 8192, 8192, 8192, 20)
    reps = int(2.2 / (burst_ms * 1e-3)) + 1
    sust_ms = device_time_ms(torch, f, reps, stream)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    pb, ps = float(peaks.get("bf16_tflops", PEAK_FALLBACK["sb"])), float(peaks.get("bf16_tflops_sustained", 0) or 0)
    tf = lambda ms_: 2.0 * m * n * k / (ms_ * 1e-3) / 1e12
    out["sbgemm_8192"] = {"kernel": ob.cblas.last_kernel(), "burst_tflops": tf(burst_ms), "burst_launches": 20, "sustained_tflops": tf(sust_ms),
                          "sustained_launches": reps, "sustained_seconds": sust_ms * reps * 1e-3, "peak_burst": pb, "peak_sustained": ps or None,
                          "frac_of_burst_peak": tf(burst_ms) / pb, "sustained_frac_of_sustained_peak": (tf(sust_ms) / ps) if ps else None,
                          "sustained_frac_of_burst_peak": tf(sust_ms) / pb, "other_ops_tflops": r_ops, "parity_worst_ratio": max(parity_nn, w_ops),
                          "order": "burst NN (best of 3 x 20 launches), the other ops (20 launches each), then the sustained NN loop, then 3 s idle"}
    del a, b, c
    time.sleep(3.0)      # let the power state settle before the FP32 / FP64 measurements

    # (b) FP32 / complex at the sizes the targets are quoted on
    # complex: BASELINE config 5 -- all nine N/T/C combinations, alpha = (0.7, -0.9), beta = (1.3, -1.1) (ctest/zin3:11-13)
    NTC9 = [(x, y) for y in (0, 1, 3) for x in (0, 1, 3)]
    CA, CB_ = 0.7 - 0.9j, 1.3 - 1.1j
    for key, dtype, size, ops, reps in (("sgemm_16384", "s", 16384, NT4, 3), ("zgemm_8192", "z", 8192, NTC9, 3), ("cgemm_8192", "c", 8192, NTC9, 3)):
        peak, src = measured_peak(dtype, dev.index or 0)
        cplx = dtype in ("z", "c")
        r, kern, w = one(dtype, ops, size, size, size, reps, alpha=CA if cplx else 1.0, beta=CB_ if cplx else 0.0)
        out[key] = {"tflops_real_flops": r, "tflops_2mnk": {o: v * 2.0 / FLOP_FACTOR[dtype] for o, v in r.items()}, "kernel": kern, "peak": peak,
                    "frac": r["NN"] / peak, "worst_op": min(r, key=r.get), "worst_op_frac": min(r.values()) / peak, "parity_worst_ratio": w,
                    "alpha_beta": "(0.7,-0.9), (1.3,-1.1)" if cplx else "1, 0"}
    out["peak_source_fp32"] = measured_peak("s", dev.index or 0)[1]

    # (c) BASELINE config 5, tall-skinny: flop bound and HBM bound side by side (algorithmic bytes = A + B + C once)
    ts = {}
    for dtype in ("z", "c"):
        peak, _ = measured_peak(dtype, dev.index or 0)
        for (mm, nn, kk) in ((65536, 256, 65536), (65536, 65536, 256)):
            r, kern, w = one(dtype, [(0, 0), (3, 0)] if nn == 256 else [(0, 0)], mm, nn, kk, 3, alpha=CA, beta=CB_)
            bytes_alg = ES[dtype] * (mm * kk + kk * nn + 2 * mm * nn)          # beta != 0: C is read and written
            flops = FLOP_FACTOR[dtype] * mm * nn * kk
            t_flop, t_hbm = flops / (peak * 1e12), (bytes_alg / (hbm * 1e9)) if hbm else None
            ts[f"{dtype}gemm_{mm}x{nn}x{kk}"] = {"tflops_real_flops": r, "tflops_2mnk": {o: v * 2.0 / FLOP_FACTOR[dtype] for o, v in r.items()}, "kernel": kern,
                                                 "frac_of_flop_peak": r["NN"] / peak, "algorithmic_bytes": bytes_alg,
                                                 "flop_bound_ms": t_flop * 1e3, "hbm_bound_ms": t_hbm * 1e3 if t_hbm else None,
                                                 "measured_ms": flops / (r["NN"] * 1e12) * 1e3, "parity_worst_ratio": w}
            torch.cuda.empty_cache()
    out["tall_skinny"] = ts

    # (d) BASELINE config 2 sweep, NN (the 16384^3 DGEMM point is the headline itself)
    sw = {}
    for dtype in ("d", "s"):
        peak, _ = measured_peak(dtype, dev.index or 0)
        for size in (1024, 2048, 4096, 8192, 12288):
            r, kern, w = one(dtype, [(0, 0)], size, size, size, 20 if size <= 4096 else 5, check=size <= 4096)
            sw[f"{dtype}gemm_{size}"] = {"tflops": r["NN"], "frac": r["NN"] / peak, "kernel": kern}
    out["square_sweep_nn"] = sw
    # SBGEMM over sizes (8192^3 is above): small grids leave SMs idle (no split-K), see DESIGN 6a
    pb = out["sbgemm_8192"]["peak_burst"]
    sbs = {}
    for size in (1024, 2048, 4096, 16384):
        r, kern, w = one("sb", NT4 if size == 4096 else [(0, 0)], size, size, size, 20 if size <= 4096 else 5, check=size <= 4096)
        sbs[f"sbgemm_{size}"] = {"tflops": r, "frac_of_burst_peak": r["NN"] / pb, "kernel": kern}
    out["sbgemm_sweep"] = sbs
    # (e) SURVEY 8 f3: the rest of level 3 at 8192 (timing only; parity is the -m gpu tests' job), and SBGEMMT
    torch.cuda.empty_cache()
    fam = {}
    for row in level3_rows(torch, ob, dev, ["d", "s", "z", "c"], [8192], []):
        fam[row["routine"]] = {"ms": row["ms"], "tflops_useful": row["tflops_useful"], "frac_of_gemm_peak": row["frac_of_gemm_peak"],
                               "launches_per_call": row["launches_per_call"]}
    fam["sbgemmt"] = sbgemmt_row(torch, ob, dev, 8192)
    fam["note"] = "n = k = 8192, device-resident operands, wall time of the synchronous Fortran-ABI calls (3 after a warm-up); flops as a GEMM of the same useful work"
    out["level3_8192"] = fam
    # (f) SURVEY 8 f1: ?GEMM3M on three real products against ?GEMM, 8192^3 NN, device operands (guarded: a failure here
    # must not cost the line its headline)
    try:
        out["gemm3m_8192"] = gemm3m_rows(torch, ob, dev, 8192)
    except Exception as exc:       # noqa: BLE001
        out["gemm3m_8192"] = {"error": repr(exc)}
    return out


def gemm3m_rows(torch, ob, dev, n):
    import ctypes as C
    lib = ob.lib()
    i_ = lambda v: C.byref(C.c_int(int(v)))
    res = {"note": "wall time of the synchronous Fortran-ABI calls (3 after two warm-ups); TFLOP/s in ZGEMM's 8mnk count for both"}
    for dtype, rdt, ct in (("z", torch.float64, C.c_double), ("c", torch.float32, C.c_float)):
        a, b, c = (torch.rand((n, n, 2), device=dev, dtype=rdt) - 0.5 for _ in range(3))
        al, be = (ct * 2)(0.7, -0.9), (ct * 2)(0.0, 0.0)
        row = {}
        for name in ("gemm_", "gemm3m_"):
            f = lambda: getattr(lib, dtype + name)(C.c_char_p(b"N"), C.c_char_p(b"N"), i_(n), i_(n), i_(n), al, C.c_void_p(a.data_ptr()), i_(n),
                                                   C.c_void_p(b.data_ptr()), i_(n), be, C.c_void_p(c.data_ptr()), i_(n))
            f(); f(); torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for _ in range(3):
                f()
            ms = (time.perf_counter() - t0) / 3 * 1e3
            row[name + "ms"], row[name + "tflops_8mnk"], row[name + "kernel"] = ms, 8.0 * n ** 3 / (ms * 1e-3) / 1e12, ob.cblas.last_kernel()
        row["speedup"] = row["gemm_ms"] / row["gemm3m_ms"]
        res[dtype] = row
        del a, b, c
        torch.cuda.empty_cache()
    return res


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
    import torch
    import openblas_b200 as ob
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    return level3_rows(torch, ob, dev, [d for d in args.sweep_dtypes.split(",") if d in "sdcz"], [int(x) for x in args.sizes.split(",")],
                       [r for r in args.sweep_routines.split(",") if r], echo=True)


def level3_rows(torch, ob, dev, dtypes, sizes, want, echo=False):
    """SURVEY 8 f3 routines through their Fortran entry points on device-resident n x n operands: wall time of the
    (synchronous) calls, 3 after one warm-up; one row per routine."""
    import ctypes as C
    lib = ob.lib()
    rows = []
    i_ = lambda v: C.byref(C.c_int(int(v)))
    for dtype in dtypes:
        cplx = dtype in "cz"
        rdt = torch.float64 if dtype in "dz" else torch.float32
        ct = C.c_double if dtype in "dz" else C.c_float
        peak, _ = measured_peak(dtype)
        for nsz in sizes:
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
                if want and not any(name.startswith(r) for r in want):
                    continue
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
                if echo:
                    print(json.dumps(rows[-1]), flush=True)
            del a, b, c
    return rows


def sbgemmt_row(torch, ob, dev, n):
    """SBGEMMT (sbgemmt_, lower triangle, NN) on device-resident bf16 operands: the triangle's n (n + 1) k flops per second."""
    import ctypes as C
    lib = ob.lib()
    i_ = lambda v: C.byref(C.c_int(int(v)))
    a = (torch.rand((n, n), device=dev) - 0.5).to(torch.bfloat16)
    b = (torch.rand((n, n), device=dev) - 0.5).to(torch.bfloat16)
    c = torch.zeros((n, n), device=dev, dtype=torch.float32)
    al, be = C.c_float(1.0), C.c_float(0.0)
    f = lambda: lib.sbgemmt_(C.c_char_p(b"L"), C.c_char_p(b"N"), C.c_char_p(b"N"), i_(n), i_(n), C.byref(al), C.c_void_p(a.data_ptr()), i_(n),
                             C.c_void_p(b.data_ptr()), i_(n), C.byref(be), C.c_void_p(c.data_ptr()), i_(n))
    f(); torch.cuda.synchronize()
    reps = 10
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    ms = (time.perf_counter() - t0) / reps * 1e3
    # element (i, j) of the column-major C is c[j, i]: the strict upper triangle (i < j) must still be zero
    untouched = bool((torch.tril(c, diagonal=-1) == 0).all().item())
    return {"routine": "sbgemmt", "n": n, "k": n, "ms": ms, "tflops_triangle": n * (n + 1.0) * n / (ms * 1e-3) / 1e12, "kernel": ob.cblas.last_kernel(),
            "other_triangle_untouched": untouched}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dtype", default="d", choices=list(DT))
    ap.add_argument("--m", type=int, default=0); ap.add_argument("--n", type=int, default=0); ap.add_argument("--k", type=int, default=0)
    ap.add_argument("--nb", type=int, default=4096, help="SUMMA distribution block / panel width")
    ap.add_argument("--summa", default="c", choices=["c", "py"], help="N > 1: the library's own driver (csrc/summa.cu) or its Python mirror over torch.distributed broadcasts")
    ap.add_argument("--ref-n", type=int, default=0, help="reference arm: force an n^3 sample (default: the GPU arm's workload, halved only if it cannot fit --ref-budget)")
    ap.add_argument("--ref-budget", type=float, default=240.0, help="reference arm: seconds the (steps + warmup) calls may take")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true"); ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra north-star measurements (SBGEMM, SGEMM, Z/C, tall-skinny, mid sizes)")
    ap.add_argument("--no-peaks", action="store_true", help="skip the in-run FP64/FP32 peak probe (tools/peaks quick)")
    ap.add_argument("--sweep", action="store_true"); ap.add_argument("--sizes", default="1024,2048,4096,8192,16384")
    ap.add_argument("--sweep-level3", action="store_true", help="developer view: SYMM/SYRK/SYR2K/HEMM/HERK on device operands")
    ap.add_argument("--sweep-dtypes", default="d,s,z,c,sb")
    ap.add_argument("--sweep-routines", default="", help="--sweep-level3: only routines whose name starts with one of these (comma separated)")
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
