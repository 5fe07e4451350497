#!/usr/bin/env python3
"""cblas_dgemm_batch timing: one group of `count` n x n x n matrices in host memory.
usage: batch_time.py [n] [count] [reps]      (B200_BATCH_PACKED=0 selects the matrix-by-matrix path)"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import openblas_b200 as ob

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
count = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
rng = np.random.default_rng(0)
a = rng.standard_normal((count, n, n)); b = rng.standard_normal((count, n, n)); c = np.zeros((count, n, n))
I = lambda v: (C.c_int * 1)(v)
adr = lambda x: C.cast(x, C.c_void_p)
P = lambda x: (C.c_void_p * count)(*[x[i].ctypes.data for i in range(count)])
A, B, Cc = P(a), P(b), P(c)
alpha, beta = (C.c_double * 1)(1.0), (C.c_double * 1)(0.0)
args = (ob.cblas.ColMajor, adr(I(ob.cblas.NoTrans)), adr(I(ob.cblas.NoTrans)), adr(I(n)), adr(I(n)), adr(I(n)), adr(alpha), adr(A),
        adr(I(n)), adr(B), adr(I(n)), adr(beta), adr(Cc), adr(I(n)), 1, adr(I(count)))
keep = args
ob.lib().cblas_dgemm_batch(*args)
ts = []
for _ in range(reps):
    t = time.perf_counter(); ob.lib().cblas_dgemm_batch(*args); ts.append(time.perf_counter() - t)
ref = np.einsum("bkj,bik->bij", a, b)      # column-major product seen through row-major numpy
err = float(np.abs(c - ref).max())
best = min(ts)
print(f"dgemm_batch n={n} count={count} packed={os.environ.get('B200_BATCH_PACKED', '1')}: {best * 1e3:.3f} ms/batch, "
      f"{2.0 * n ** 3 * count / best / 1e9:.1f} GFLOP/s, max err {err:.2e}, kernel {ob.cblas.last_kernel()}")
