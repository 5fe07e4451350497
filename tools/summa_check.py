#!/usr/bin/env python3
"""Multi-GPU correctness check of openblas_b200/summa.py (run under torchrun, NCCL):
every rank builds the same global A, B, C0 from a seed, takes its block-cyclic pieces, runs the
SUMMA sweep with the library's DGEMM as the local product, and compares its piece of the result with
the same piece of a single-GPU b200 DGEMM of the whole problem (computed redundantly on every rank).
Prints one line per rank; exit code 1 on mismatch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import openblas_b200 as ob
from openblas_b200 import summa

m, n, k, nb = (int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (4096, 3072, 5120, 512)))
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
g = torch.Generator(device=dev); g.manual_seed(7)
A = torch.rand((k, m), generator=g, device=dev, dtype=torch.float64) - 0.5
B = torch.rand((n, k), generator=g, device=dev, dtype=torch.float64) - 0.5
C0 = torch.rand((n, m), generator=g, device=dev, dtype=torch.float64) - 0.5
grid = summa.make_grid(world, rank)
a_loc = summa.scatter_from_global(A, nb, grid, "p", "q")
b_loc = summa.scatter_from_global(B, nb, grid, "p", "q")
c_loc = summa.scatter_from_global(C0, nb, grid, "p", "q")
D = 1
gemm = lambda mm, nn, kk, al, a, lda, b, ldb, be, c, ldc, st: ob.cblas.gemm_device(D, 0, 0, mm, nn, kk, al, a, lda, b, ldb, be, c, ldc, st)
sm = summa.Summa(grid, m, n, k, nb, torch.float64, dev, gemm)
sm.run(0.7, a_loc, b_loc, 1.3, c_loc)
torch.cuda.synchronize()
want = C0.clone()
ob.cblas.gemm_device(D, 0, 0, m, n, k, 0.7, A, m, B, k, 1.3, want, m, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
want_loc = summa.scatter_from_global(want, nb, grid, "p", "q")
err = float((want_loc - c_loc).abs().max()) if c_loc.numel() else 0.0
scale = float(want_loc.abs().max()) if c_loc.numel() else 1.0
ok = err <= 1e-11 * max(1.0, scale) * (k / nb)
print(f"summa_check rank {rank}/{world} grid {grid.P}x{grid.Q} ({grid.p},{grid.q}) local {sm.m_loc}x{sm.n_loc} panels {len(sm.steps)} max|diff|={err:.3e} {'OK' if ok else 'MISMATCH'}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
