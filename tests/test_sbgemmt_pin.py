"""CPU tests that PIN oracle_sbgemmt / oracle_check_sbgemmt (oracle/level3_oracle.c) to the reference's SBGEMMT
(interface/sbgemmt.c over kernel/x86_64/sbgemv_n.c / sbgemv_t.c):

  1. against tests/golden/sbgemmt_golden.npz -- outputs of the unmodified reference (generic target, one thread)
     written by tests/golden/make_sbgemmt_golden.py -- BIT FOR BIT (same summation order, bf16 products exact in fp32)
  2. live against oracle/_ref/generic where present: random shapes, both ABIs and orders, NaN in everything that
     must not be read or written, k == 0, alpha == 0; with the reference's threads on, within the k * eps bound
  3. the library's host code (interface_level3.c, runtime_level3.inl) over the CUDA stand-in (tests/hostsim): both
     the triangle-masked launch and the block-column scheme, host pointers, against the oracle
"""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import cpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EPS32 = 2.0 ** -23
needs_ref = pytest.mark.skipif(not cpu.have_reference("generic"), reason="oracle/_ref/generic not built")


def tri_mask(m, ldc, uplo):
    jj, ii = np.meshgrid(np.arange(m), np.arange(ldc), indexing="ij")
    return ((ii >= jj) if uplo else (ii <= jj)) & (ii < m)


def expected(oracle, uplo, ta, tb, m, k, alpha, a, lda, b, ldb, beta, c0, ldc, rowmajor):
    """what a (possibly row-major) SBGEMMT call must leave in C: the row-major branch swaps the operands and their
    ops but keeps uplo (sbgemmt.c:232-240), so it is the column-major problem below on the same bytes"""
    want = c0.copy()
    if rowmajor:
        gauge = oracle.sbgemmt(uplo, tb, ta, m, k, alpha, b, ldb, a, lda, beta, want, ldc)
    else:
        gauge = oracle.sbgemmt(uplo, ta, tb, m, k, alpha, a, lda, b, ldb, beta, want, ldc)
    return want, gauge, tri_mask(m, ldc, uplo)


def same_bits(x, y):
    return np.array_equal(x.view(np.uint32), y.view(np.uint32))


def test_oracle_matches_the_reference_golden_bitwise(oracle):
    g = np.load(os.path.join(ROOT, "tests", "golden", "sbgemmt_golden.npz"))
    n = int(g["count"][0])
    assert n == 59
    for i in range(n):
        key = f"case{i}"
        uplo, ta, tb, m, k, lda, ldb, ldc, cblas, rowmajor = (int(v) for v in g[key + "_meta"])
        alpha, beta = (float(v) for v in g[key + "_scal"])
        want, _, mask = expected(oracle, uplo, ta, tb, m, k, alpha, g[key + "_a"], lda, g[key + "_b"], ldb, beta, g[key + "_c0"], ldc, rowmajor)
        assert same_bits(want, g[key + "_c"]), (i, uplo, ta, tb, m, k, cblas, rowmajor)
        assert same_bits(g[key + "_c"][~mask], g[key + "_c0"][~mask])
        if k == 0:
            assert same_bits(g[key + "_c"], g[key + "_c0"])           # the reference leaves C alone, beta or not


def problem(rng, oracle, ta, tb, m, k, rowmajor, pad=(2, 1, 3)):
    ra, ca = (k, m) if ta & 1 else (m, k)
    rb, cb = (m, k) if tb & 1 else (k, m)
    if rowmajor:
        ra, ca, rb, cb = ca, ra, cb, rb
    lda, ldb, ldc = max(ra, 1) + pad[0], max(rb, 1) + pad[1], max(m, 1) + pad[2]
    a = oracle.tobf16(rng.random((max(ca, 1), lda), dtype=np.float32) - 0.5)
    b = oracle.tobf16(rng.random((max(cb, 1), ldb), dtype=np.float32) - 0.5)
    c0 = (rng.random((max(m, 1), ldc)) - 0.5).astype(np.float32)
    return a, lda, b, ldb, c0, ldc


@needs_ref
def test_oracle_vs_live_reference_bitwise(oracle):
    ref = cpu.Reference("generic")
    ref.set_threads(1)
    rng = np.random.default_rng(77)
    for uplo in (0, 1):
        for ta in range(4):
            for tb in range(4):
                for cblas, rowmajor in ((False, False), (True, False), (True, True)):
                    for m, k in ((29, 61), (1, 7), (16, 1), (5, 0)):
                        a, lda, b, ldb, c0, ldc = problem(rng, oracle, ta, tb, m, k, rowmajor)
                        mask = tri_mask(m, ldc, uplo)
                        c0[~mask] = np.nan                          # the other triangle and the padding rows: never read, never written
                        for alpha, beta in ((0.7, 1.3), (1.0, 0.0), (0.0, 0.5), (0.0, 1.0)):
                            start = c0.copy()
                            if beta == 0.0 and k > 0:
                                start[mask] = np.nan                # beta == 0 never reads C
                            got = start.copy()
                            cpu.call_sbgemmt(ref.lib, uplo, ta, tb, m, k, alpha, a, lda, b, ldb, beta, got, ldc, cblas=cblas, rowmajor=rowmajor)
                            want, _, _ = expected(oracle, uplo, ta, tb, m, k, alpha, a, lda, b, ldb, beta, start, ldc, rowmajor)
                            assert same_bits(got, want), (uplo, ta, tb, cblas, rowmajor, m, k, alpha, beta)


@needs_ref
def test_oracle_vs_threaded_reference_within_the_bound(oracle):
    """above 9216 elements per column the reference splits each SBGEMV over threads (sbgemmt.c:372-375,
    driver/level2/sbgemv_thread.c); the transposed kernel then sums partial dot products: bound, not bits"""
    ref = cpu.Reference("generic")
    ref.set_threads(4)
    rng = np.random.default_rng(78)
    try:
        for ta, tb in ((0, 0), (1, 0), (0, 1), (1, 1)):
            m, k = 150, 320
            a, lda, b, ldb, c0, ldc = problem(rng, oracle, ta, tb, m, k, False)
            got = c0.copy()
            cpu.call_sbgemmt(ref.lib, 1, ta, tb, m, k, 0.7, a, lda, b, ldb, 1.3, got, ldc)
            want, gauge, mask = expected(oracle, 1, ta, tb, m, k, 0.7, a, lda, b, ldb, 1.3, c0, ldc, False)
            d = np.abs(got.astype(np.float64) - want.astype(np.float64))[:, :m]
            ratio = (d[mask[:, :m]] / ((k + 2) * EPS32 * np.maximum(gauge[:m, :m][mask[:, :m]], 1e-300))).max()
            assert ratio <= 1.0, (ta, tb, ratio)
            assert same_bits(got[~mask], c0[~mask])
    finally:
        ref.set_threads(1)


def test_argument_table_restates_sbgemmt_c(oracle):
    """interface/sbgemmt.c:121-137 and :286-301 written out by hand, position by position"""
    ok = -1
    assert oracle.check_sbgemmt(0, 0, 0, 0, 3, 2, 3, 2, 3, ok) == ok
    assert oracle.check_sbgemmt(0, -1, -1, -1, -1, -1, 0, 0, 0, ok) == 1
    assert oracle.check_sbgemmt(0, 0, -1, -1, -1, -1, 0, 0, 0, ok) == 2
    assert oracle.check_sbgemmt(0, 0, 0, -1, -1, -1, 0, 0, 0, ok) == 3
    assert oracle.check_sbgemmt(0, 0, 0, 0, -1, -1, 0, 0, 0, ok) == 4
    assert oracle.check_sbgemmt(0, 0, 0, 0, 3, -1, 0, 0, 0, ok) == 5
    assert oracle.check_sbgemmt(0, 0, 0, 0, 3, 2, 2, 1, 2, ok) == 8       # lda < m
    assert oracle.check_sbgemmt(0, 0, 1, 0, 3, 2, 1, 1, 2, ok) == 8       # transposed A: lda < k
    assert oracle.check_sbgemmt(0, 0, 0, 0, 3, 2, 3, 1, 2, ok) == 10      # ldb < k
    assert oracle.check_sbgemmt(0, 0, 0, 1, 3, 2, 3, 2, 2, ok) == 10      # transposed B: ldb < m
    assert oracle.check_sbgemmt(0, 0, 0, 0, 3, 2, 3, 2, 2, ok) == 13
    assert oracle.check_sbgemmt(0, 0, 0, 0, 0, 0, 1, 1, 0, ok) == 13      # ldc < max(1, m)
    # row-major, as the swapped problem sees them (a = caller's B): the positions swap too
    assert oracle.check_sbgemmt(1, 0, -1, 0, 3, 2, 3, 3, 3, ok) == 3
    assert oracle.check_sbgemmt(1, 0, 0, -1, 3, 2, 3, 3, 3, ok) == 2
    assert oracle.check_sbgemmt(1, 0, 0, 0, 3, 2, 2, 3, 3, ok) == 10
    assert oracle.check_sbgemmt(1, 0, 0, 0, 3, 2, 3, 1, 3, ok) == 8


# ---------------------------------------------------------------- the library's host code over the CUDA stand-in
@pytest.fixture(scope="module")
def hostsim():
    import subprocess
    out = subprocess.check_output([os.path.join(ROOT, "tests", "hostsim", "build.sh")], text=True).strip().splitlines()[-1]
    return C.CDLL(out)


@pytest.mark.parametrize("scheme", ["1", "0"])
def test_library_host_code_over_the_stand_in(hostsim, oracle, scheme, monkeypatch):
    monkeypatch.setenv("B200_RANKK_TRI", scheme)          # 1: one triangle-masked launch, 0: block columns + merge
    rng = np.random.default_rng(79)
    for uplo in (0, 1):
        for ta in range(4):
            for tb in range(4):
                for cblas, rowmajor in ((False, False), (True, True)):
                    for m, k in ((150, 40), (33, 7), (4, 3)):
                        a, lda, b, ldb, c0, ldc = problem(rng, oracle, ta, tb, m, k, rowmajor)
                        mask = tri_mask(m, ldc, uplo)
                        for alpha, beta in ((0.7, 1.3), (1.0, 0.0), (0.0, 0.5)):
                            got = c0.copy()
                            cpu.call_sbgemmt(hostsim, uplo, ta, tb, m, k, alpha, a, lda, b, ldb, beta, got, ldc, cblas=cblas, rowmajor=rowmajor)
                            want, gauge, _ = expected(oracle, uplo, ta, tb, m, k, alpha, a, lda, b, ldb, beta, c0, ldc, rowmajor)
                            d = np.abs(got.astype(np.float64) - want.astype(np.float64))[:, :m]
                            inside = mask[:, :m]
                            ratio = (d[inside] / ((k + 2) * EPS32 * np.maximum(gauge[:m, :m][inside], 1e-300))).max()
                            assert ratio <= 1.0, (scheme, uplo, ta, tb, cblas, rowmajor, m, k, alpha, ratio)
                            assert same_bits(got[~mask], c0[~mask]), (scheme, uplo, ta, tb, cblas, rowmajor, m, k, "outside the triangle")
    # k == 0: untouched whatever beta is
    a, lda, b, ldb, c0, ldc = problem(rng, oracle, 0, 0, 20, 0, False)
    got = c0.copy()
    cpu.call_sbgemmt(hostsim, 1, 0, 0, 20, 0, 0.7, a, lda, b, ldb, 0.0, got, ldc)
    assert same_bits(got, c0)
