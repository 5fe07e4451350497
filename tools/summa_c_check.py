#!/usr/bin/env python3
"""Multi-GPU correctness and timing check of the library's own SUMMA driver (csrc/summa.cu through
openblas_b200.summa.CSumma); run under torchrun (NCCL), one rank per GPU:

  torchrun --nproc-per-node N tools/summa_c_check.py [m n k nb] [--time M N K NB]

Every rank generates its block-cyclic pieces of hashed global matrices, runs the distributed product with device
operands, with HOST operands (pinned and pageable), with beta != 0, twice in a row (window reuse), and compares its
C block with one single-GPU DGEMM of the global rows / columns it owns (this library's own kernel, computed
redundantly) and with long-double dot products of sampled entries.  Exit code 1 on any mismatch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import openblas_b200 as ob
from openblas_b200 import summa

args = [a for a in sys.argv[1:] if not a.startswith("--")]
m, n, k, nb = (int(x) for x in (args[:4] if len(args) >= 4 else (3000, 2500, 2200, 256)))
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
cs = summa.CSumma(world, rank, dev)
if rank == 0:
    print(cs.describe(), flush=True)
D = 1
stream = torch.cuda.current_stream(dev)
bad = 0


def pieces(m, n, k, nb):
    ri = torch.tensor(summa.local_index_map(m, nb, cs.p, cs.P), dtype=torch.int64, device=dev)
    cj = torch.tensor(summa.local_index_map(n, nb, cs.q, cs.Q), dtype=torch.int64, device=dev)
    ka = torch.tensor(summa.local_index_map(k, nb, cs.q, cs.Q), dtype=torch.int64, device=dev)
    kb = torch.tensor(summa.local_index_map(k, nb, cs.p, cs.P), dtype=torch.int64, device=dev)
    return ri, cj, summa.hashed_entries(1, ri, ka), summa.hashed_entries(2, kb, cj), summa.hashed_entries(3, ri, cj)


def check(tag, m, n, k, nb, alpha, beta, c_got, ri, cj, c0):
    global bad
    if ri.numel() == 0 or cj.numel() == 0:
        print(f"[{rank}] {tag}: empty block OK", flush=True)
        return
    allk = torch.arange(k, dtype=torch.int64, device=dev)
    a_rows = summa.hashed_entries(1, ri, allk)          # (k, m_loc)
    b_cols = summa.hashed_entries(2, allk, cj)          # (n_loc, k)
    want = c0.clone()
    ob.cblas.gemm_device(D, 0, 0, ri.numel(), cj.numel(), k, alpha, a_rows, ri.numel(), b_cols, k, beta, want, ri.numel(), stream.cuda_stream)
    torch.cuda.synchronize(dev)
    got = c_got if c_got.is_cuda else c_got.to(dev)
    err = float((got - want).abs().max())
    # sampled entries in long double
    g = torch.Generator(); g.manual_seed(11 + rank)
    worst = 0.0
    for _ in range(8):
        il, jl = int(torch.randint(0, ri.numel(), (1,), generator=g)), int(torch.randint(0, cj.numel(), (1,), generator=g))
        ar = a_rows[:, il].cpu().numpy().astype(np.longdouble)
        bc = b_cols[jl, :].cpu().numpy().astype(np.longdouble)
        ref = alpha * np.dot(ar, bc) + beta * np.longdouble(float(c0[jl, il]))
        gauge = abs(alpha) * float(np.dot(np.abs(ar), np.abs(bc))) + abs(beta) * abs(float(c0[jl, il]))
        worst = max(worst, float(abs(np.longdouble(float(got[jl, il])) - ref)) / ((k + 2) * 2.0 ** -52 * gauge))
    ok = err <= 1e-11 * max(1.0, float(want.abs().max())) and worst <= 2.0
    bad += 0 if ok else 1
    print(f"[{rank}] {tag}: block {ri.numel()}x{cj.numel()} max|diff vs one-GPU dgemm|={err:.2e} sampled ratio={worst:.4f} {'OK' if ok else 'MISMATCH'}", flush=True)


ri, cj, a_loc, b_loc, c0 = pieces(m, n, k, nb)
m_loc, n_loc, ka_loc, kb_loc = cs.local_shapes(m, n, k, nb)
assert (m_loc, n_loc) == (ri.numel(), cj.numel()) and a_loc.shape == (ka_loc, m_loc) and b_loc.shape == (n_loc, kb_loc)
for rep in range(2):                                             # second pass reuses the window (flags of the previous call)
    c = c0.clone()
    cs.gemm(D, m, n, k, nb, 0.7, a_loc, max(1, m_loc), b_loc, max(1, kb_loc), 1.3, c, max(1, m_loc), stream.cuda_stream)
    torch.cuda.synchronize(dev)
    check(f"device operands pass {rep}", m, n, k, nb, 0.7, 1.3, c, ri, cj, c0)
for pin in (True, False):
    ha, hb, hc = a_loc.cpu(), b_loc.cpu(), c0.cpu().clone()
    if pin:
        ha, hb, hc = ha.pin_memory(), hb.pin_memory(), hc.pin_memory()
    cs.gemm(D, m, n, k, nb, 1.0, ha, max(1, m_loc), hb, max(1, kb_loc), 0.0 if pin else 0.5, hc, max(1, m_loc), stream.cuda_stream)
    check(f"host operands ({'pinned' if pin else 'pageable'})", m, n, k, nb, 1.0, 0.0 if pin else 0.5, hc, ri, cj, c0)
# k = 0 and a bigger problem after a small one (window growth)
c = c0.clone()
cs.gemm(D, m, n, 0, nb, 1.0, a_loc, max(1, m_loc), b_loc, max(1, kb_loc), 0.5, c, max(1, m_loc), stream.cuda_stream)
torch.cuda.synchronize(dev)
bad += 0 if torch.equal(c, c0 * 0.5) else 1
m2, n2, k2, nb2 = 2 * m + 37, 2 * n + 5, k + 300, 2 * nb
ri2, cj2, a2, b2, c02 = pieces(m2, n2, k2, nb2)
c = c02.clone()
cs.gemm(D, m2, n2, k2, nb2, -1.0, a2, max(1, ri2.numel()), b2, max(1, b2.shape[1]), 1.0, c, max(1, ri2.numel()), stream.cuda_stream)
torch.cuda.synchronize(dev)
check("after window growth", m2, n2, k2, nb2, -1.0, 1.0, c, ri2, cj2, c02)

if "--time" in sys.argv:
    i = sys.argv.index("--time")
    M, N, K, NB = (int(x) for x in sys.argv[i + 1:i + 5])
    del a_loc, b_loc, c0, c, a2, b2, c02
    torch.cuda.empty_cache()
    ml, nl, kal, kbl = cs.local_shapes(M, N, K, NB)
    g = torch.Generator(device=dev); g.manual_seed(rank)
    A = torch.rand((kal, ml), generator=g, device=dev, dtype=torch.float64) - 0.5
    B = torch.rand((nl, kbl), generator=g, device=dev, dtype=torch.float64) - 0.5
    Cm = torch.empty((nl, ml), device=dev, dtype=torch.float64)
    for _ in range(2):
        cs.gemm(D, M, N, K, NB, 1.0, A, ml, B, kbl, 0.0, Cm, ml, stream.cuda_stream)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record(stream)
    for _ in range(reps):
        cs.gemm(D, M, N, K, NB, 1.0, A, ml, B, kbl, 0.0, Cm, ml, stream.cuda_stream)
    e1.record(stream)
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"TIME {M}x{N}x{K} nb={NB} on {world} GPUs: {float(ms):.2f} ms/step, {2.0 * M * N * K / float(ms) / 1e9:.2f} TFLOP/s aggregate "
              f"({2.0 * M * N * K / world / float(ms) / 1e9:.2f} per GPU)", flush=True)

tot = torch.tensor([bad], device=dev)
if world > 1:
    dist.all_reduce(tot)
    dist.barrier()
cs.close()
if world > 1:
    dist.destroy_process_group()
if rank == 0:
    print("SUMMA C DRIVER:", "ALL OK" if int(tot) == 0 else f"{int(tot)} MISMATCHES", flush=True)
sys.exit(0 if int(tot) == 0 else 1)
