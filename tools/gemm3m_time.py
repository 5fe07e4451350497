#!/usr/bin/env python3
"""?GEMM3M (three real GEMMs) against ?GEMM (four-multiply complex kernel) on device-resident operands, NN, wall time of
the synchronous Fortran-ABI calls (3 after a warm-up).  One JSON line per point; TFLOP/s are counted as 8 mnk for both,
i.e. what a ZGEMM of the same problem would be credited with."""
import ctypes as C
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import openblas_b200 as ob  # noqa: E402


def main():
    lib = ob.lib()
    i_ = lambda v: C.byref(C.c_int(int(v)))
    for dtype, rdt, ct in (("z", torch.float64, C.c_double), ("c", torch.float32, C.c_float)):
        for n in (2048, 4096, 8192):
            a, b, c = (torch.rand((n, n, 2), device="cuda", dtype=rdt) - 0.5 for _ in range(3))
            al, be = (ct * 2)(0.7, -0.9), (ct * 2)(0.0, 0.0)
            row = {"dtype": dtype, "n": n}
            for name in ("gemm_", "gemm3m_"):
                f = lambda: getattr(lib, dtype + name)(C.c_char_p(b"N"), C.c_char_p(b"N"), i_(n), i_(n), i_(n), al, C.c_void_p(a.data_ptr()), i_(n),
                                                       C.c_void_p(b.data_ptr()), i_(n), be, C.c_void_p(c.data_ptr()), i_(n))
                f(); f(); torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(3):
                    f()
                ms = (time.perf_counter() - t0) / 3 * 1e3
                row[name + "ms"] = ms
                row[name + "tflops_8mnk"] = 8.0 * n ** 3 / (ms * 1e-3) / 1e12
                row[name + "kernel"] = ob.cblas.last_kernel()
            row["speedup"] = row["gemm_ms"] / row["gemm3m_ms"]
            print(json.dumps(row), flush=True)
            del a, b, c


if __name__ == "__main__":
    main()
