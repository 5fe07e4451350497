#!/usr/bin/env python3
"""Per-launch time of small device-resident GEMMs launched back to back on one stream (the regime of
gemm_batch and of the deep recursion levels of TRMM/TRSM).  usage: small_gemm_time.py [dtype] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import openblas_b200 as ob

dtype = sys.argv[1] if len(sys.argv) > 1 else "d"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 500
code = {"s": 0, "d": 1, "c": 2, "z": 3}[dtype]
tdt = {"s": torch.float32, "d": torch.float64, "c": torch.complex64, "z": torch.complex128}[dtype]
dev = torch.device("cuda", 0)
s = torch.cuda.current_stream(dev)
for (m, n, k) in [(32, 32, 32), (64, 64, 64), (128, 128, 128), (256, 256, 256), (64, 8192, 64), (128, 8192, 128), (256, 8192, 256), (512, 8192, 512)]:
    a = torch.rand((max(k, m), max(m, k)), device=dev, dtype=torch.float64).to(tdt)
    b = torch.rand((max(n, k), max(k, n)), device=dev, dtype=torch.float64).to(tdt)
    c = torch.zeros((n, m), device=dev, dtype=tdt)
    f = lambda: ob.cblas.gemm_device(code, 0, 0, m, n, k, 1.0, a, a.shape[1], b, b.shape[1], 0.0, c, m, s.cuda_stream)
    for _ in range(5):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print(f"{dtype} {m}x{n}x{k}: {us:8.1f} us/launch  {2.0 * m * n * k * (4 if dtype in 'cz' else 1) / us / 1e6:8.2f} TFLOP/s  {ob.cblas.last_kernel()}  "
          f"tile={os.environ.get('B200_DGEMM_TILE', 'auto')}", flush=True)
