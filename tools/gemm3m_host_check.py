#!/usr/bin/env python3
"""zgemm3m_ on HOST operands large enough for the panel pipeline (8192^3, pinned): every 4096 x 4096 block of C is
computed from three real products; compared entry by entry with zgemm_ of the same call (4-multiply kernel) under the
3M acceptance ratio, timed beside it."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import openblas_b200 as ob  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
lib = ob.lib()
i_ = lambda v: C.byref(C.c_int(int(v)))
a, b = (torch.view_as_complex((torch.rand((n, n, 2), dtype=torch.float64) - 0.5)).pin_memory() for _ in range(2))
c3, c4 = torch.zeros((n, n), dtype=torch.complex128).pin_memory(), torch.zeros((n, n), dtype=torch.complex128).pin_memory()
al, be = (C.c_double * 2)(0.7, -0.9), (C.c_double * 2)(0.0, 0.0)
out = {"n": n}
for name, c in (("zgemm_", c4), ("zgemm3m_", c3)):
    f = lambda: getattr(lib, name)(C.c_char_p(b"N"), C.c_char_p(b"C"), i_(n), i_(n), i_(n), al, C.c_void_p(a.data_ptr()), i_(n), C.c_void_p(b.data_ptr()), i_(n),
                                   be, C.c_void_p(c.data_ptr()), i_(n))
    f()
    t0 = time.perf_counter(); f(); out[name + "ms"] = (time.perf_counter() - t0) * 1e3
    out[name + "last_kernel"] = ob.cblas.last_kernel()
d = (c3 - c4)[::37, ::41]
err = (d.real.abs() + d.imag.abs()).max().item()
# gauge of an entry: |alpha| sum (|re| + |im|)(|re| + |im|) ~ 1.6 * n * 0.25 for uniform(-0.5, 0.5) parts; use the exact one on a few entries
A1 = (a.real.abs() + a.imag.abs()); B1 = (b.real.abs() + b.imag.abs())
rows = A1[:, ::4099].T.contiguous()                       # rows i of op(A) = A (column-major: a[l, i])
cols = B1[:, ::4099].T.contiguous()                       # op(B) = B^H: element (l, j) = conj(B(j, l)) = b[l, j]
gauge = 1.6 * (rows @ cols.T)
dd = (c3 - c4)[::4099, ::4099].T
ratio = ((dd.real.abs() + dd.imag.abs()) / (2.0 ** -52 * gauge)).max().item()
out.update({"max_abs1_diff_sampled": err, "ratio_3m_vs_4m_sampled": ratio, "ok": ratio < 16.0})
print(json.dumps(out))
