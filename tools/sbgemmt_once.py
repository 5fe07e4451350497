#!/usr/bin/env python3
"""Two SBGEMMT calls (lower triangle, NN, n = k = 8192 by default) on device-resident operands: the target of an
ncu capture (`ncu -k regex:sbgemm_tcgen05 --launch-skip 1 --launch-count 1 ...`)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import openblas_b200 as ob  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
lib = ob.lib()
a = (torch.rand((n, n), device="cuda") - 0.5).to(torch.bfloat16)
b = (torch.rand((n, n), device="cuda") - 0.5).to(torch.bfloat16)
c = torch.zeros((n, n), device="cuda", dtype=torch.float32)
i_ = lambda v: C.byref(C.c_int(int(v)))
al, be = C.c_float(1.0), C.c_float(0.0)
for _ in range(2):
    lib.sbgemmt_(C.c_char_p(b"L"), C.c_char_p(b"N"), C.c_char_p(b"N"), i_(n), i_(n), C.byref(al), C.c_void_p(a.data_ptr()), i_(n),
                 C.c_void_p(b.data_ptr()), i_(n), C.byref(be), C.c_void_p(c.data_ptr()), i_(n))
torch.cuda.synchronize()
print("kernel", ob.cblas.last_kernel())
