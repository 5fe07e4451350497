#!/usr/bin/env python3
"""One warm-up and one measured DTRSM (left, lower, no-trans, non-unit) on device operands, for a launch list under ncu."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import openblas_b200 as ob
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
lib = ob.lib()
a = (torch.rand((n, n), dtype=torch.float64, device="cuda") - 0.5) / n + torch.eye(n, dtype=torch.float64, device="cuda")
b = torch.rand((n, n), dtype=torch.float64, device="cuda") - 0.5
i_ = lambda v: C.byref(C.c_int(v)); al = C.c_double(1.0)
for _ in range(2):
    lib.dtrsm_(C.c_char_p(b"L"), C.c_char_p(b"L"), C.c_char_p(b"N"), C.c_char_p(b"N"), i_(n), i_(n), C.byref(al), C.c_void_p(a.data_ptr()), i_(n),
               C.c_void_p(b.data_ptr()), i_(n))
torch.cuda.synchronize()
print("done", ob.cblas.launch_count())
