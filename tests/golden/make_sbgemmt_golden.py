#!/usr/bin/env python3
"""Writes tests/golden/sbgemmt_golden.npz: inputs and the REFERENCE's outputs for SBGEMMT (interface/sbgemmt.c),
from oracle/_ref/generic (the unmodified reference compiled by oracle/build_ref.py), one thread.  Every
uplo x transa x transb through sbgemmt_, a third of them also through cblas_sbgemmt in both orders (the
row-major branch keeps uplo as given, sbgemmt.c:239-240), k == 0 (C untouched whatever beta is) and alpha == 0.
Run in the authoring container (needs /root/reference built into oracle/_ref); the .npz travels."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cpu  # noqa: E402


def cases():
    """(uplo, ta, tb, m, k, alpha, beta, cblas, rowmajor)"""
    n = 0
    for uplo in (0, 1):
        for ta in range(4):
            for tb in range(4):
                yield uplo, ta, tb, 11, 23, 0.7, 1.3, False, False
                if n % 3 == 0:
                    yield uplo, ta, tb, 9, 14, 1.0, 0.0, True, False
                    yield uplo, ta, tb, 9, 14, -0.6, 0.5, True, True
                n += 1
    yield 1, 0, 1, 6, 0, 0.7, 0.5, False, False      # k == 0
    yield 0, 1, 0, 6, 0, 0.7, 0.0, True, True
    yield 1, 0, 0, 8, 5, 0.0, 0.5, False, False      # alpha == 0
    yield 0, 1, 1, 8, 5, 0.0, 0.0, True, False
    yield 1, 0, 0, 70, 200, 0.7, 1.3, False, False    # 70 * 200 > 9216: the reference's threaded SBGEMV would split this one


def main():
    ref = cpu.Reference("generic")
    ref.set_threads(1)
    orc = cpu.Oracle()
    rng = np.random.default_rng(20261018)
    out, count = {}, 0
    for uplo, ta, tb, m, k, alpha, beta, cblas, rowmajor in cases():
        # shapes as the caller stores them: column-major op(A) is m x k, op(B) k x m; a row-major caller holds
        # row-major arrays, i.e. the same bytes read as the transposed column-major ones
        ra, ca = (k, m) if ta & 1 else (m, k)
        rb, cb = (m, k) if tb & 1 else (k, m)
        if rowmajor:
            ra, ca, rb, cb = ca, ra, cb, rb
        lda, ldb, ldc = max(ra, 1) + 1, max(rb, 1) + 2, m + 3
        a = orc.tobf16(rng.random((max(ca, 1), lda), dtype=np.float32) - 0.5)
        b = orc.tobf16(rng.random((max(cb, 1), ldb), dtype=np.float32) - 0.5)
        c0 = (rng.random((m, ldc)) - 0.5).astype(np.float32)
        c = c0.copy()
        cpu.call_sbgemmt(ref.lib, uplo, ta, tb, m, k, alpha, a, lda, b, ldb, beta, c, ldc, cblas=cblas, rowmajor=rowmajor)
        key = f"case{count}"
        count += 1
        out[key + "_meta"] = np.array([uplo, ta, tb, m, k, lda, ldb, ldc, int(cblas), int(rowmajor)], dtype=np.int64)
        out[key + "_scal"] = np.array([alpha, beta], dtype=np.float32)
        out[key + "_a"], out[key + "_b"], out[key + "_c0"], out[key + "_c"] = a, b, c0, c
    out["count"] = np.array([count])
    path = os.path.join(ROOT, "tests", "golden", "sbgemmt_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, count, "cases", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
