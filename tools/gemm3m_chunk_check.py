#!/usr/bin/env python3
"""zgemm3m_ with the workspace bound lowered (B200_3M_WORKSPACE_BYTES) so that k = 4096 goes through in chunks, on device
operands, ops N/C and T/N, beta != 0: compared entry by entry with zgemm_ of the same call under the 3M acceptance ratio;
the launch count tells how many chunks ran."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
m, n, k = 1536, 2048, 4096
rs = 8
plane = lambda rows, cols: -(-(-(-rows * rs // 128) * 128) * cols // 256) * 256
need = lambda kc, ta, tb: 3 * ((plane(kc, m) if ta else plane(m, kc)) + (plane(n, kc) if tb else plane(kc, n)) + plane(m, n))
import openblas_b200 as ob  # noqa: E402

lib = ob.lib()
i_ = lambda v: C.byref(C.c_int(int(v)))
out = []
for (cha, chb, ta, tb) in ((b"N", b"C", 0, 1), (b"T", b"N", 1, 0)):
    os.environ["B200_3M_WORKSPACE_BYTES"] = str(need(1024, ta, tb))          # chunks of 1024 fit, 2048 do not: 4 chunks
    ra, ca = (k, m) if ta else (m, k)
    rb, cb = (n, k) if tb else (k, n)
    a = torch.view_as_complex(torch.rand((ca, ra + 2, 2), device="cuda", dtype=torch.float64) - 0.5)
    b = torch.view_as_complex(torch.rand((cb, rb + 4, 2), device="cuda", dtype=torch.float64) - 0.5)
    c0 = torch.view_as_complex(torch.rand((n, m + 6, 2), device="cuda", dtype=torch.float64) - 0.5)
    c3, c4 = c0.clone(), c0.clone()
    al, be = (C.c_double * 2)(0.7, -0.9), (C.c_double * 2)(1.3, -1.1)
    res = {"ops": (cha + chb).decode()}
    for name, c in (("zgemm_", c4), ("zgemm3m_", c3)):
        before = ob.cblas.launch_count()
        getattr(lib, name)(C.c_char_p(cha), C.c_char_p(chb), i_(m), i_(n), i_(k), al, C.c_void_p(a.data_ptr()), i_(ra + 2), C.c_void_p(b.data_ptr()), i_(rb + 4),
                           be, C.c_void_p(c.data_ptr()), i_(m + 6))
        torch.cuda.synchronize()
        res[name + "launches"] = ob.cblas.launch_count() - before
        res[name + "last_kernel"] = ob.cblas.last_kernel()
    d = (c3 - c4)[:, :m]
    err = (d.real.abs() + d.imag.abs()).max().item()
    # |alpha| sum (|re| + |im|)(|re| + |im|) for uniform(-0.5, 0.5) parts is about 1.14 * 0.25 * k; the exact gauge is not needed for a yes / no at this margin
    gauge = 1.14 * 0.25 * k
    res.update({"max_abs1_diff": err, "ratio": err / (2.0 ** -52 * gauge), "padding_rows_untouched": bool(torch.equal(torch.view_as_real(c3[:, m:]), torch.view_as_real(c0[:, m:])))})
    res["ok"] = res["ratio"] < 16.0 and res["zgemm3m_launches"] == 24 and res["padding_rows_untouched"]
    out.append(res)
print(json.dumps(out))
