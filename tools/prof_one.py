#!/usr/bin/env python3
"""Profiling helper: run ONE precision/shape a few times on device-resident operands (for ncu).
usage: prof_one.py <dtype s|d|c|z|sb> <n> [ta tb] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import openblas_b200 as ob

dtype, n = sys.argv[1], int(sys.argv[2])
ta, tb = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (0, 0)
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
code = {"s": 0, "d": 1, "c": 2, "z": 3, "sb": 4}[dtype]
tdt = {"s": torch.float32, "d": torch.float64, "c": torch.complex64, "z": torch.complex128, "sb": torch.bfloat16}[dtype]
odt = torch.float32 if dtype == "sb" else tdt
dev = torch.device("cuda", 0)
if tdt.is_complex:
    mk = lambda: torch.view_as_complex(torch.rand((n, n, 2), device=dev, dtype=torch.float64 if dtype == "z" else torch.float32) - 0.5)
else:
    mk = lambda: (torch.rand((n, n), device=dev, dtype=torch.float32 if dtype == "sb" else tdt) - 0.5).to(tdt)
a, b = mk(), mk()
c = torch.zeros((n, n), dtype=odt, device=dev)
s = torch.cuda.current_stream(dev)
for _ in range(reps):
    ob.cblas.gemm_device(code, ta, tb, n, n, n, 1.0, a, n, b, n, 0.0, c, n, s.cuda_stream)
torch.cuda.synchronize()
print(dtype, n, ta, tb, ob.cblas.last_kernel(), float(c.float().abs().sum() if not c.is_complex() else c.abs().sum()))
