"""CPU tests that PIN the round-2 restatements (oracle/level3_oracle.c: oracle_gemmt, oracle_sbgemv,
oracle_sbdot and their argument checks) to the reference:

  1. against tests/golden/f_rows_golden.npz -- outputs of the unmodified reference (generic target) written
     by tests/golden/make_f_rows_golden.py: ?gemmt within the summation-order bound (the reference sums each
     column with GEMV), SBGEMV and SBDOT bit for bit (same order, products of bf16 values are exact in fp32)
  2. live against oracle/_ref/generic where it is present: bigger shapes, NaN in everything that must not be
     read or written, both ABIs and row-major, argument checks value for value
  3. the shapes and scalars of the reference's own acceptance test utest/test_extensions/test_dgemmt.c:215-2040
     (gemmt against the triangle of gemm), restated against the GEMM oracle
"""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import cpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EPS = {cpu.S: 2.0 ** -23, cpu.CX: 2.0 ** -23, cpu.D: 2.0 ** -52, cpu.Z: 2.0 ** -52}


def operand(rng, dtype, cols, ld):
    x = rng.random((cols, ld)) - 0.5
    if dtype in (cpu.CX, cpu.Z):
        x = x + 1j * (rng.random((cols, ld)) - 0.5)
    return x.astype(cpu.NP_OUT[dtype])


def tri_mask(m, ldc, uplo):
    """(m, ldc) boolean array, True inside the uplo triangle of the column-major m x m matrix"""
    jj, ii = np.meshgrid(np.arange(m), np.arange(ldc), indexing="ij")
    return ((ii >= jj) if uplo else (ii <= jj)) & (ii < m)


def colmajor_problem(meta):
    """the column-major problem a (possibly row-major) gemmt call denotes: interface/gemmt.c:305-330"""
    dtype, uplo, ta, tb, m, k, lda, ldb, ldc, cblas, rowmajor = (int(v) for v in meta)
    return dtype, uplo, ta, tb, m, k, lda, ldb, ldc, cblas, rowmajor


def gemmt_expected(oracle, dtype, uplo, ta, tb, m, k, alpha, a, lda, b, ldb, beta, c0, ldc, rowmajor):
    want = c0.copy()
    if rowmajor:      # C^T = op(B)^T op(A)^T: operands swap, uplo flips
        gauge = oracle.gemmt(dtype, 1 - uplo, tb, ta, m, k, alpha, b, ldb, a, lda, beta, want, ldc)
        return want, gauge, tri_mask(m, ldc, 1 - uplo)
    gauge = oracle.gemmt(dtype, uplo, ta, tb, m, k, alpha, a, lda, b, ldb, beta, want, ldc)
    return want, gauge, tri_mask(m, ldc, uplo)


def check_gemmt(dtype, m, k, ldc, got, want, gauge, mask, c0, what):
    d = np.abs(got.astype(np.complex128) - want.astype(np.complex128))[:, :m]
    g = gauge[:m, :m]
    inside = mask[:, :m]
    ratio = (d[inside] / ((k + 2) * EPS[dtype] * np.maximum(g[inside], 1e-300))).max() if inside.any() else 0.0
    assert ratio <= 2.0, (what, ratio)
    assert np.array_equal(got.view(np.uint8)[~np.repeat(mask, got.itemsize, axis=1)], c0.view(np.uint8)[~np.repeat(mask, got.itemsize, axis=1)]), \
        (what, "bytes outside the triangle changed")


def test_gemmt_oracle_matches_reference_golden(oracle):
    g = np.load(os.path.join(ROOT, "tests", "golden", "f_rows_golden.npz"))
    for i in range(int(g["gemmt_count"][0])):
        key = f"gemmt{i}"
        dtype, uplo, ta, tb, m, k, lda, ldb, ldc, cblas, rowmajor = colmajor_problem(g[key + "_meta"])
        cplx = dtype in (cpu.CX, cpu.Z)
        alpha, beta = ((0.7 - 0.9j, 1.3 - 1.1j) if cplx else (0.7, 1.3))
        want, gauge, mask = gemmt_expected(oracle, dtype, uplo, ta, tb, m, k, alpha, g[key + "_a"], lda, g[key + "_b"], ldb, beta,
                                           g[key + "_c0"], ldc, rowmajor)
        check_gemmt(dtype, m, k, ldc, g[key + "_c"], want, gauge, mask, g[key + "_c0"], key)


def test_sbgemv_and_sbdot_oracle_match_reference_golden_bitwise(oracle):
    g = np.load(os.path.join(ROOT, "tests", "golden", "f_rows_golden.npz"))
    for i in range(int(g["sbgemv_count"][0])):
        key = f"sbgemv{i}"
        trans, m, n, lda, incx, incy = (int(v) for v in g[key + "_meta"])
        alpha, beta = (float(v) for v in g[key + "_ab"])
        y = g[key + "_y0"].copy()
        oracle.sbgemv(trans, m, n, alpha, g[key + "_a"], lda, g[key + "_x"], incx, beta, y, incy)
        assert np.array_equal(y.view(np.uint32), g[key + "_y"].view(np.uint32)), key
    for i in range(int(g["sbdot_count"][0])):
        key = f"sbdot{i}"
        n, incx, incy = (int(v) for v in g[key + "_meta"])
        d, _ = oracle.sbdot(n, g[key + "_x"], incx, g[key + "_y"], incy)
        assert np.float32(d) == g[key + "_d"][0], key


needs_ref = pytest.mark.skipif(not cpu.have_reference("generic"), reason="oracle/_ref/generic not built")


@needs_ref
def test_gemmt_oracle_vs_live_reference(oracle):
    ref = cpu.Reference("generic")
    ref.set_threads(1)
    rng = np.random.default_rng(5)
    for dtype in (cpu.S, cpu.D, cpu.CX, cpu.Z):
        cplx = dtype in (cpu.CX, cpu.Z)
        for uplo in (0, 1):
            for ta in range(4 if cplx else 2):
                for tb in range(4 if cplx else 2):
                    for cblas, rowmajor in ((False, False), (True, False), (True, True)):
                        m, k = 37, 150
                        ra, ca = (k, m) if ta & 1 else (m, k)
                        rb, cb = (m, k) if tb & 1 else (k, m)
                        if rowmajor:
                            ra, ca, rb, cb = ca, ra, cb, rb
                        lda, ldb, ldc = ra + 2, rb + 1, m + 3
                        a, b, c0 = operand(rng, dtype, ca, lda), operand(rng, dtype, cb, ldb), operand(rng, dtype, m, ldc)
                        mask = tri_mask(m, ldc, (1 - uplo) if rowmajor else uplo)
                        c0[~mask] = np.nan                       # outside the triangle (and the padding rows): never read, never written
                        for alpha, beta in (((0.7 - 0.9j, 1.3 - 1.1j) if cplx else (0.7, 1.3)), (1.0, 0.0), (0.0, 0.5)):
                            start = c0.copy()
                            if beta == 0.0:
                                start[mask] = np.nan             # beta == 0 never reads C
                            got = start.copy()
                            cpu.call_gemmt(ref.lib, dtype, uplo, ta, tb, m, k, alpha, a.copy(), lda, b.copy(), ldb, beta, got, ldc, cblas=cblas, rowmajor=rowmajor)
                            want, gauge, mk = gemmt_expected(oracle, dtype, uplo, ta, tb, m, k, alpha, a, lda, b, ldb, beta, start, ldc, rowmajor)
                            check_gemmt(dtype, m, k, ldc, got, want, gauge, mk, start, (dtype, uplo, ta, tb, cblas, rowmajor, alpha))


XERBLA_T = C.CFUNCTYPE(None, C.c_char_p, C.POINTER(C.c_int), C.c_int)


@needs_ref
def test_argument_checks_match_the_live_reference(oracle):
    """every illegal-argument probe: the info value the reference hands its xerbla_ equals oracle_check_*"""
    ref = cpu.Reference("generic")
    seen = []
    # the reference calls its own xerbla_ (DEEPBIND), which prints; capture through the info the oracle predicts
    # instead: run the reference in a child so its message can be parsed from stderr
    import subprocess, sys, textwrap
    probes = []
    for rowmajor in (0, 1):
        for (uplo, ta, tb, m, k, lda, ldb, ldc) in ((-1, 0, 0, 2, 2, 2, 2, 2), (0, -1, 0, 2, 2, 2, 2, 2), (0, 0, -1, 2, 2, 2, 2, 2), (0, 0, 0, -1, 2, 2, 2, 2),
                                                    (0, 0, 0, 2, -1, 2, 2, 2), (0, 0, 0, 3, 2, 2, 3, 3), (0, 1, 0, 3, 4, 3, 4, 3), (0, 0, 0, 3, 4, 3, 3, 3),
                                                    (0, 0, 1, 3, 2, 3, 2, 3), (0, 0, 0, 3, 2, 3, 2, 2), (0, 0, 0, 3, 2, 2, 1, 2), (1, 1, 1, 3, 5, 4, 2, 3),
                                                    (0, 1, 1, 3, 5, 2, 4, 3)):
            probes.append((rowmajor, uplo, ta, tb, m, k, lda, ldb, ldc))
    code = textwrap.dedent(f"""
        import sys, ctypes as C
        sys.path.insert(0, {ROOT!r})
        from oracle import cpu
        ref = cpu.Reference("generic")
        buf = (C.c_double * 64)()
        libc = C.CDLL(None)
        U = {{-1: 0, 0: 121, 1: 122}}; T = {{-1: 0, 0: 111, 1: 112}}
        for p in {probes!r}:
            rowmajor, uplo, ta, tb, m, k, lda, ldb, ldc = p
            sys.stdout.write("probe %r\\n" % (p,)); sys.stdout.flush()
            ref.lib.cblas_dgemmt(101 if rowmajor else 102, U[uplo], T[ta], T[tb], C.c_int(m), C.c_int(k), C.c_double(1.0), buf, C.c_int(lda), buf,
                                 C.c_int(ldb), C.c_double(0.0), buf, C.c_int(ldc))
            libc.fflush(None)
    """)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    got, cur = {}, None
    for ln in r.stdout.splitlines():
        if ln.startswith("probe "):
            cur = eval(ln[6:]); got[cur] = -1
        elif "parameter number" in ln and cur is not None:
            got[cur] = int(ln.split("parameter number")[1].split()[0])
    for p in probes:
        rowmajor, uplo, ta, tb, m, k, lda, ldb, ldc = p
        if rowmajor:      # the swapped problem: a = B, b = A
            want = oracle.check_gemmt(1, (1 - uplo) if uplo >= 0 else -1, tb, ta, m, k, ldb, lda, ldc, -1)
        else:
            want = oracle.check_gemmt(0, uplo, ta, tb, m, k, lda, ldb, ldc, -1)
        assert got[p] == want, (p, got[p], want)


@needs_ref
def test_sbgemv_sbdot_oracle_vs_live_reference_bitwise(oracle):
    ref = cpu.Reference("generic")
    ref.set_threads(1)
    rng = np.random.default_rng(6)
    for trans in (0, 1):
        for cblas, rowmajor in ((False, False), (True, False), (True, True)):
            for incx, incy in ((1, 1), (-3, 2)):
                m, n = 83, 131
                lda = m + 5
                a = oracle.tobf16(rng.random((n, lda), dtype=np.float32) - 0.5)
                # row-major m x n with leading dimension lda is the column-major n x m matrix: cblas swaps m, n and flips trans
                cm, cn, ctr = (n, m, 1 - trans) if rowmajor else (m, n, trans)
                if rowmajor:
                    a = oracle.tobf16(rng.random((m, n + 5), dtype=np.float32) - 0.5); lda = n + 5
                lenx, leny = (cm, cn) if ctr else (cn, cm)
                x = oracle.tobf16(rng.random(1 + (lenx - 1) * abs(incx), dtype=np.float32) - 0.5)
                y0 = (rng.random(1 + (leny - 1) * abs(incy)) - 0.5).astype(np.float32)
                for alpha, beta in ((0.7, 1.3), (1.0, 0.0), (0.0, 0.5), (0.0, 1.0)):
                    got, want = y0.copy(), y0.copy()
                    cpu.call_sbgemv(ref.lib, trans, m, n, alpha, a, lda, x, incx, beta, got, incy, cblas=cblas, rowmajor=rowmajor)
                    oracle.sbgemv(ctr, cm, cn, alpha, a, lda, x, incx, beta, want, incy)
                    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (trans, cblas, rowmajor, incx, incy, alpha, beta)
    for n in (0, 5, 4097):
        for incx, incy in ((1, 1), (-2, 3)):
            x = oracle.tobf16(rng.random(1 + max(n - 1, 0) * abs(incx), dtype=np.float32) - 0.5)
            y = oracle.tobf16(rng.random(1 + max(n - 1, 0) * abs(incy), dtype=np.float32) - 0.5)
            for cblas in (False, True):
                assert np.float32(cpu.call_sbdot(ref.lib, n, x, incx, y, incy, cblas=cblas)) == np.float32(oracle.sbdot(n, x, incx, y, incy)[0])


def test_gemmt_is_the_triangle_of_gemm_like_the_reference_utest(oracle):
    """utest/test_extensions/test_dgemmt.c: ?gemmt_trusted = ?gemm, then only the triangle is compared.  Its
    shapes (m = k = 50 ... with alpha / beta in {0, 1, 2}) against the bit-exact GEMM oracle."""
    rng = np.random.default_rng(8)
    for dtype in (cpu.D, cpu.Z):
        cplx = dtype == cpu.Z
        for uplo in (0, 1):
            for ta, tb in ((0, 0), (1, 0), (0, 1), (1, 1)) + (((3, 2), (2, 3)) if cplx else ()):
                for alpha, beta in ((1.0, 1.0), (2.0, 0.0), (0.0, 2.0), (1.0, 0.0)):
                    m, k = 50, 50
                    a, b, c0 = operand(rng, dtype, m, m), operand(rng, dtype, m, m), operand(rng, dtype, m, m)
                    t, full = c0.copy(), c0.copy()
                    gauge = oracle.gemmt(dtype, uplo, ta, tb, m, k, alpha, a, m, b, m, beta, t, m)
                    oracle.gemm(dtype, ta, tb, m, m, k, alpha, a, m, b, m, beta, full, m)
                    mask = tri_mask(m, m, uplo)
                    d = np.abs(t.astype(np.complex128) - full.astype(np.complex128))
                    assert (d[mask] <= 2 * (k + 2) * EPS[dtype] * np.maximum(gauge[mask], 1e-300)).all()
                    assert np.array_equal(t[~mask], c0[~mask])
