#!/usr/bin/env python3
"""Writes tests/golden/f_rows_golden.npz: inputs and the REFERENCE's outputs (oracle/_ref/generic, i.e. the
unmodified reference compiled by oracle/build_ref.py; OPENBLAS_NUM_THREADS=1) for the SURVEY 8 f3/f4 routines
added in round 2 -- ?gemmt (interface/gemmt.c), sbgemv (interface/sbgemv.c), sbdot (interface/bf16dot.c).
Run in the authoring container (needs /root/reference built into oracle/_ref); the .npz travels."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cpu  # noqa: E402


def operand(rng, dtype, cols, ld):
    t = cpu.NP_OUT[dtype]
    x = rng.random((cols, ld)) - 0.5
    if dtype in (cpu.CX, cpu.Z):
        x = x + 1j * (rng.random((cols, ld)) - 0.5)
    return x.astype(t)


def main():
    ref = cpu.Reference("generic")
    ref.set_threads(1)
    orc = cpu.Oracle()
    rng = np.random.default_rng(20261017)
    out, idx = {}, 0
    # ---- gemmt: every uplo x transa x transb, Fortran ABI; a few CBLAS row-major ones
    for dtype in (cpu.S, cpu.D, cpu.CX, cpu.Z):
        cplx = dtype in (cpu.CX, cpu.Z)
        nt = 4 if cplx else 2
        for uplo in (0, 1):
            for ta in range(nt):
                for tb in range(nt):
                    for (m, k, cblas, rowmajor) in ((13, 9, False, False), (7, 21, True, True)) if (ta + tb + uplo) % 3 == 0 else ((13, 9, False, False),):
                        # stored shapes as the column-major problem sees them; row-major callers hold the transposes,
                        # which is the same memory with the roles below (cpu.call_gemmt passes it through)
                        ra, ca = (k, m) if ta & 1 else (m, k)
                        rb, cb = (m, k) if tb & 1 else (k, m)
                        if rowmajor:
                            ra, ca, rb, cb = ca, ra, cb, rb
                        lda, ldb, ldc = ra + 2, rb + 1, m + 3
                        a, b, c0 = operand(rng, dtype, ca, lda), operand(rng, dtype, cb, ldb), operand(rng, dtype, m, ldc)
                        alpha, beta = ((0.7 - 0.9j, 1.3 - 1.1j) if cplx else (0.7, 1.3))
                        c = c0.copy()
                        cpu.call_gemmt(ref.lib, dtype, uplo, ta, tb, m, k, alpha, a.copy(), lda, b.copy(), ldb, beta, c, ldc, cblas=cblas, rowmajor=rowmajor)   # the reference conjugates its "b" operand IN PLACE for transb = R / C (gemmt.c:466-476) and leaves it so
                        key = f"gemmt{idx}"; idx += 1
                        out[key + "_meta"] = np.array([dtype, uplo, ta, tb, m, k, lda, ldb, ldc, int(cblas), int(rowmajor)])
                        out[key + "_a"], out[key + "_b"], out[key + "_c0"], out[key + "_c"] = a, b, c0, c
    out["gemmt_count"] = np.array([idx])
    # ---- sbgemv / sbdot
    idx = 0
    for trans in (0, 1):
        for (m, n) in ((17, 9), (5, 40), (33, 31)):
            for incx, incy in ((1, 1), (2, 3), (-1, 2), (3, -2)):
                for alpha, beta in ((0.7, 1.3), (1.0, 0.0), (0.0, 0.5)):
                    lda = m + 3
                    lenx, leny = (m, n) if trans else (n, m)
                    a = orc.tobf16(rng.random((n, lda), dtype=np.float32) - 0.5)
                    x = orc.tobf16(rng.random(1 + (lenx - 1) * abs(incx), dtype=np.float32) - 0.5)
                    y0 = (rng.random(1 + (leny - 1) * abs(incy)) - 0.5).astype(np.float32)
                    y = y0.copy()
                    cpu.call_sbgemv(ref.lib, trans, m, n, alpha, a, lda, x, incx, beta, y, incy)
                    key = f"sbgemv{idx}"; idx += 1
                    out[key + "_meta"] = np.array([trans, m, n, lda, incx, incy])
                    out[key + "_ab"] = np.array([alpha, beta], dtype=np.float32)
                    out[key + "_a"], out[key + "_x"], out[key + "_y0"], out[key + "_y"] = a, x, y0, y
    out["sbgemv_count"] = np.array([idx])
    idx = 0
    for n in (0, 1, 7, 100, 1000):
        for incx, incy in ((1, 1), (2, 3), (-1, 2), (-2, -1)):
            x = orc.tobf16(rng.random(1 + max(n - 1, 0) * abs(incx), dtype=np.float32) - 0.5)
            y = orc.tobf16(rng.random(1 + max(n - 1, 0) * abs(incy), dtype=np.float32) - 0.5)
            d = cpu.call_sbdot(ref.lib, n, x, incx, y, incy)
            key = f"sbdot{idx}"; idx += 1
            out[key + "_meta"] = np.array([n, incx, incy])
            out[key + "_x"], out[key + "_y"], out[key + "_d"] = x, y, np.array([d], dtype=np.float32)
    out["sbdot_count"] = np.array([idx])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "f_rows_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
