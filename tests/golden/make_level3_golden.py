#!/usr/bin/env python3
"""Writes tests/golden/level3_golden.npz, trxm_golden.npz and errexit_level3_reference.txt: fixtures
produced by the REFERENCE itself for the rest of level 3 (SYMM/HEMM, SYRK/HERK, SYR2K/HER2K,
TRMM/TRSM), run in the authoring container where /root/reference is mounted and oracle/_ref has
been built from it by oracle/build_ref.py (GENERIC target, one thread: deterministic).

  level3_golden.npz                 seeded inputs and the reference's outputs for every precision,
                                    routine, side / uplo / trans combination, ragged sizes, padded
                                    leading dimensions, the ctest alpha/beta values and the
                                    alpha == 0 / k == 0 / beta == 0 / beta == 1 corners
  trxm_golden.npz                   the same for TRMM / TRSM: every side / uplo / trans / diag combination
  errexit_level3_reference.txt      what the reference's entry points hand to xerbla_ for the
                                    table of illegal calls in tests/c/errexit_level3.c
  errexit_fuzz_reference.txt        the same for the 4000 fixed-seed random calls of tests/c/errexit_fuzz.c
                                    (GEMM included)
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cpu  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def operand(rng, dtype, cols, ld):
    x = rng.random((cols, ld)) - 0.5
    if dtype in (cpu.CX, cpu.Z):
        x = x + 1j * (rng.random((cols, ld)) - 0.5)
    return x.astype(cpu.NP_IN[dtype])


def cases():
    """(kind, dtype, herm, two_or_side, uplo, trans, m, n, k, alpha, beta)"""
    sizes = [(1, 1, 1), (2, 3, 5), (7, 5, 3), (9, 19, 2), (19, 9, 7), (13, 11, 24), (17, 18, 21)]
    i = 0
    for dtype in (cpu.S, cpu.D, cpu.CX, cpu.Z):
        cplx = dtype in (cpu.CX, cpu.Z)
        scal = [(0.7 - 0.9j, 1.3 - 1.1j), (0.0, 1.3 - 1.1j), (1.0, 0.0), (0.7 - 0.9j, 1.0)] if cplx else \
               [(0.7, 1.3), (0.0, 1.3), (1.0, 0.0), (0.7, 1.0)]
        for herm in ((0, 1) if cplx else (0,)):
            for side in (0, 1):
                for uplo in (0, 1):
                    for rep in range(2):
                        m, n, _ = sizes[i % len(sizes)]
                        alpha, beta = scal[i % len(scal)] if rep else scal[0]
                        i += 1
                        yield "symm", dtype, herm, side, uplo, 0, m, n, 0, alpha, beta
            for two in (0, 1):
                for uplo in (0, 1):
                    for trans in (0, 1):
                        for rep in range(2):
                            _, n, k = sizes[i % len(sizes)]
                            alpha, beta = scal[i % len(scal)] if rep else scal[0]
                            if herm:
                                beta = complex(beta).real
                                if not two:
                                    alpha = complex(alpha).real
                            if rep and i % 5 == 0:
                                k = 0
                            i += 1
                            yield "rankk", dtype, herm, two, uplo, trans, n, n, k, alpha, beta


def main():
    ref = cpu.Reference("generic")
    ref.set_threads(1)
    rng = np.random.default_rng(20261018)
    out, meta = {}, []
    for idx, (kind, dtype, herm, x, uplo, trans, m, n, k, alpha, beta) in enumerate(cases()):
        if kind == "symm":
            ka = n if x else m
            lda, ldb, ldc = ka + 2, m + 1, m + 3
            a, b, c0 = operand(rng, dtype, ka, lda), operand(rng, dtype, n, ldb), operand(rng, dtype, n, ldc)
            c0[:, m:] = -1e10
            c = c0.copy()
            cpu.call_symm(ref.lib, dtype, herm, x, uplo, m, n, alpha, a, lda, b, ldb, beta, c, ldc)
        else:
            rows, cols = (k, n) if trans else (n, k)
            lda, ldb, ldc = max(rows, 1) + 1, max(rows, 1) + 2, n + 3
            a, b, c0 = operand(rng, dtype, max(cols, 1), lda), operand(rng, dtype, max(cols, 1), ldb), operand(rng, dtype, n, ldc)
            c0[:, n:] = -1e10
            c = c0.copy()
            cpu.call_rankk(ref.lib, dtype, herm, x, uplo, trans, n, k, alpha, a, lda, b, ldb, beta, c, ldc)
        if kind == "rankk" and not x:
            b = b[:0]                      # SYRK / HERK have no B
        out[f"a{idx}"], out[f"b{idx}"], out[f"c0_{idx}"], out[f"c{idx}"] = a, b, c0, c
        meta.append([0 if kind == "symm" else 1, dtype, herm, x, uplo, trans, m, n, k, lda, ldb, ldc, complex(alpha).real,
                     complex(alpha).imag, complex(beta).real, complex(beta).imag])
    out["meta"] = np.array(meta, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "level3_golden.npz"), **out)
    print("level3_golden.npz:", len(meta), "cases")

    # ---- TRMM / TRSM: every side / uplo / trans / diag combination; A's diagonal gets +1 like ctest's
    # DMAKE (c_dblat3.f:2084-2087), its unreferenced triangle and a unit diagonal are set to NaN
    rng = np.random.default_rng(20261019)
    out, meta = {}, []
    sizes = [(1, 1), (2, 3), (7, 5), (9, 19), (19, 9), (13, 11), (17, 18)]
    i = 0
    for dtype in (cpu.S, cpu.D, cpu.CX, cpu.Z):
        cplx = dtype in (cpu.CX, cpu.Z)
        for solve in (0, 1):
            for side in (0, 1):
                for uplo in (0, 1):
                    for trans in range(4 if cplx else 2):
                        for unit in (0, 1):
                            m, n = sizes[i % len(sizes)]
                            alpha = [(0.7 - 0.9j) if cplx else 0.7, 1.0, 0.0][i % 7 // 3 if i % 7 >= 5 else 0]
                            i += 1
                            ka = n if side else m
                            lda, ldb = ka + 2, m + 3
                            a, b0 = operand(rng, dtype, ka, lda), operand(rng, dtype, n, ldb)
                            a[np.arange(ka), np.arange(ka)] += 1.0
                            jj, ii = np.meshgrid(np.arange(ka), np.arange(lda), indexing="ij")
                            a[(ii < jj) if uplo else ((ii > jj) & (ii < ka))] = np.nan
                            if unit:
                                a[np.arange(ka), np.arange(ka)] = np.nan
                            b0[:, m:] = -1e10
                            b = b0.copy()
                            cpu.call_trxm(ref.lib, dtype, solve, side, uplo, trans, unit, m, n, alpha, a, lda, b, ldb)
                            idx = len(meta)
                            out[f"a{idx}"], out[f"b0_{idx}"], out[f"b{idx}"] = a, b0, b
                            meta.append([dtype, solve, side, uplo, trans, unit, m, n, lda, ldb, complex(alpha).real, complex(alpha).imag])
    out["meta"] = np.array(meta, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "trxm_golden.npz"), **out)
    print("trxm_golden.npz:", len(meta), "cases")

    refdir = os.path.dirname(cpu.ref_path("generic"))
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "errexit_level3")
        subprocess.check_call(["gcc", "-O1", "-Wall", f"-I{ROOT}/include", os.path.join(ROOT, "tests", "c", "errexit_level3.c"),
                               "-o", exe, f"-L{refdir}", "-lopenblas_ref", f"-Wl,-rpath,{refdir}"])
        text = subprocess.check_output([exe], text=True)
    open(os.path.join(OUT, "errexit_level3_reference.txt"), "w").write(text)
    print("errexit_level3_reference.txt:", len(text.splitlines()), "probes")
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "errexit_fuzz")
        subprocess.check_call(["gcc", "-O1", "-Wall", f"-I{ROOT}/include", os.path.join(ROOT, "tests", "c", "errexit_fuzz.c"),
                               "-o", exe, f"-L{refdir}", "-lopenblas_ref", f"-Wl,-rpath,{refdir}"])
        text = subprocess.check_output([exe], text=True)
    open(os.path.join(OUT, "errexit_fuzz_reference.txt"), "w").write(text)
    print("errexit_fuzz_reference.txt:", len(text.splitlines()), "probes")


if __name__ == "__main__":
    main()
