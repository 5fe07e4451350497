"""CPU tests that PIN the oracle (oracle/gemm_oracle.c) to the reference:

  1. bit-for-bit against golden vectors the reference's GENERIC-target build produced
     (tests/golden/gemm_golden.npz, written by tests/golden/make_golden.py)
  2. bit-for-bit against oracle/_ref/generic and within the summation-order bound against the
     best SIMD build, live, when oracle/_ref is present (authoring container and GPU box)
  3. the DMMCH-style checker is self-checked with exact integer data, as the reference does
     (ctest/c_dblat3.f:211-272)
  4. bf16 conversions against vectors from the reference's sbstobf16_/sbf16tos_
  5. argument validation against the expectations of the reference's error-exit test
     (ctest/c_d3chke.c:45-272)
"""
import os

import numpy as np
import pytest

from oracle import cpu
from helpers import ALL_DTYPES, C_BOUND, NAMES, THRESH, alpha_beta, ntrans, problem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_oracle_matches_reference_golden_bitwise(oracle, golden):
    meta = golden["meta"]
    assert len(meta) == 88
    for idx, row in enumerate(meta):
        dtype, ta, tb, m, n, k, lda, ldb, ldc = (int(v) for v in row[:9])
        alpha = complex(row[9], row[10]) if dtype in (cpu.CX, cpu.Z) else row[9]
        beta = complex(row[11], row[12]) if dtype in (cpu.CX, cpu.Z) else row[11]
        c = golden[f"c0_{idx}"].copy()
        oracle.gemm(dtype, ta, tb, m, n, k, alpha, golden[f"a{idx}"], lda, golden[f"b{idx}"], ldb, beta, c, ldc)
        want = golden[f"c{idx}"]
        assert np.array_equal(c.view(np.uint8), want.view(np.uint8)), (idx, NAMES[dtype], ta, tb, m, n, k)
        # padding rows of C are never written
        assert np.all(c[:, m:].real == -1e10)


@pytest.mark.skipif(not cpu.have_reference("generic"), reason="oracle/_ref not built")
@pytest.mark.parametrize("dtype", ALL_DTYPES)
def test_oracle_vs_live_reference(oracle, dtype):
    gen = cpu.Reference("generic")
    gen.set_threads(1)
    best = cpu.Reference()
    rng = np.random.default_rng(100 + dtype)
    alphas, betas = alpha_beta(dtype)
    worst = 0.0
    for (m, n, k) in [(3, 2, 1), (33, 17, 250), (64, 48, 519), (130, 70, 300)]:
        for ta in range(ntrans(dtype)):
            for tb in range(ntrans(dtype)):
                alpha, beta = alphas[(m + ta) % 3], betas[(n + tb) % 3]
                a, lda, b, ldb, c0, ldc = problem(rng, oracle, dtype, ta, tb, m, n, k)
                c1, c2, c3 = c0.copy(), c0.copy(), c0.copy()
                oracle.gemm(dtype, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c1, ldc)
                gen.gemm(dtype, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c2, ldc)
                best.gemm(dtype, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c3, ldc)
                assert np.array_equal(c1.view(np.uint8), c2.view(np.uint8)), (NAMES[dtype], m, n, k, ta, tb)
                r = oracle.ratio(dtype, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c0, ldc, c1, ldc, c3, ldc)
                worst = max(worst, r)
                # and both pass the reference's own acceptance test
                assert oracle.mmch(dtype, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c0, ldc, c3, ldc) < THRESH
    assert worst <= C_BOUND, worst


@pytest.mark.parametrize("dtype", ALL_DTYPES)
def test_small_matrix_restatement_agrees(oracle, dtype):
    """interface/gemm.c:551-571 small-matrix kernels vs the blocked driver: same result to
    rounding order."""
    rng = np.random.default_rng(7)
    for ta in range(ntrans(dtype)):
        for tb in range(ntrans(dtype)):
            m, n, k = 9, 7, 35
            a, lda, b, ldb, c0, ldc = problem(rng, oracle, dtype, ta, tb, m, n, k)
            alpha, beta = alpha_beta(dtype)[0][2], alpha_beta(dtype)[1][2]
            c1, c2 = c0.copy(), c0.copy()
            oracle.gemm(dtype, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c1, ldc)
            oracle.gemm(dtype, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c2, ldc, small=True)
            assert oracle.ratio(dtype, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c0, ldc, c1, ldc, c2, ldc) <= C_BOUND
            assert np.all(c2[:, m:].real == -1e10)


def test_mmch_selfcheck_exact_integers(oracle):
    """c_dblat3.f:211-272: with small integer data the product is exact, so the checker must
    report 0 for the right answer and a large ratio for a wrong one."""
    n = 8
    for dtype in (cpu.S, cpu.D):
        t = cpu.NP_IN[dtype]
        ab = np.array([[max(i - j + 1, 0) for i in range(n)] for j in range(n)], dtype=t)  # (cols, ld)
        bb = np.array([[j * 3 + 1 if i == 0 else 3 for i in range(n)] for j in range(n)], dtype=t)
        want = (ab.T.astype(np.float64) @ bb.T.astype(np.float64)).T.astype(t)
        c0 = np.zeros((n, n), dtype=t)
        assert oracle.mmch(dtype, 0, 0, n, n, n, 1.0, ab, n, bb, n, 0.0, c0, n, np.ascontiguousarray(want), n) == 0.0
        wrong = np.ascontiguousarray(want + 1)
        assert oracle.mmch(dtype, 0, 0, n, n, n, 1.0, ab, n, bb, n, 0.0, c0, n, wrong, n) > 1e3
        got = c0.copy()
        oracle.gemm(dtype, 0, 0, n, n, n, 1.0, ab, n, bb, n, 0.0, got, n)
        assert np.array_equal(got, want)


def test_bf16_conversion_golden(oracle):
    g = np.load(os.path.join(ROOT, "tests", "golden", "bf16_golden.npz"))
    got = oracle.tobf16(g["x"])
    assert np.array_equal(got, g["bf16"])
    back = oracle.bf16to(g["bf16"])
    assert np.array_equal(back.view(np.uint32), g["back"].view(np.uint32))


def test_beta_zero_never_reads_c(oracle):
    """kernel/generic/gemm_beta.c:52-71: NaN/Inf in C must not survive beta == 0."""
    rng = np.random.default_rng(3)
    for dtype in ALL_DTYPES:
        a, lda, b, ldb, c0, ldc = problem(rng, oracle, dtype, 0, 0, 5, 4, 3)
        c0[:, :5] = np.nan
        c = c0.copy()
        oracle.gemm(dtype, 0, 0, 5, 4, 3, 1.0, a, lda, b, ldb, 0.0, c, ldc)
        assert np.all(np.isfinite(c[:4, :5]))
        # alpha == 0: A and B are not read (level3.c:252-259) -> NaN there cannot leak
        a2 = a.copy()
        if dtype != cpu.SB:
            a2[:] = np.nan
        else:
            a2[:] = 0x7fc0
        c = operand_finite(c0)
        oracle.gemm(dtype, 0, 0, 5, 4, 3, 0.0, a2, lda, b, ldb, 1.3, c, ldc)
        assert np.all(np.isfinite(c[:4, :5]))


def operand_finite(c0):
    c = c0.copy()
    c[np.isnan(c)] = 0.25
    return c


# (order-independent part of) ctest/c_d3chke.c:45-272: each tuple is a bad call and the info
# xerbla_ must receive; trans codes after decoding, column-major
ERROR_EXITS = [
    # transa transb  m  n  k lda ldb ldc  info
    (-1, 0, 0, 0, 0, 1, 1, 1, 1),
    (-1, 1, 0, 0, 0, 1, 1, 1, 1),
    (0, -1, 0, 0, 0, 1, 1, 1, 2),
    (1, -1, 0, 0, 0, 1, 1, 1, 2),
    (0, 0, -1, 0, 0, 1, 1, 1, 3),
    (0, 1, -1, 0, 0, 1, 1, 1, 3),
    (1, 0, -1, 0, 0, 1, 1, 1, 3),
    (1, 1, -1, 0, 0, 1, 1, 1, 3),
    (0, 0, 0, -1, 0, 1, 1, 1, 4),
    (1, 1, 0, -1, 0, 1, 1, 1, 4),
    (0, 0, 0, 0, -1, 1, 1, 1, 5),
    (1, 0, 0, 0, -1, 1, 1, 1, 5),
    (0, 0, 2, 0, 0, 1, 1, 2, 8),
    (0, 1, 2, 0, 0, 1, 1, 2, 8),
    (1, 0, 0, 0, 2, 1, 2, 1, 8),
    (1, 1, 0, 0, 2, 1, 1, 1, 8),
    (0, 0, 0, 0, 2, 1, 1, 1, 10),
    (1, 0, 0, 0, 2, 2, 1, 1, 10),
    (0, 1, 0, 2, 0, 1, 1, 1, 10),
    (1, 1, 0, 2, 0, 1, 1, 1, 10),
    (0, 0, 2, 0, 0, 2, 1, 1, 13),
    (0, 1, 2, 0, 0, 2, 1, 1, 13),
    (1, 0, 2, 0, 0, 1, 1, 1, 13),
    (1, 1, 2, 0, 0, 1, 1, 1, 13),
]


def test_argument_validation_table(oracle):
    for ta, tb, m, n, k, lda, ldb, ldc, want in ERROR_EXITS:
        assert oracle.check_args(ta, tb, m, n, k, lda, ldb, ldc, -1) == want
    assert oracle.check_args(0, 0, 3, 4, 5, 3, 5, 3, -1) == -1
    # OpenBLAS accepts ldc == m == 0 (weaker than Netlib's max(1,m)): interface/gemm.c:276
    assert oracle.check_args(0, 0, 0, 4, 5, 1, 5, 0, 0) == 0


# ------------------------------------------------------------------------------------------
# symmetric level-3 family: oracle/level3_oracle.c pinned BY BOUND
# ------------------------------------------------------------------------------------------
def test_level3_oracle_matches_reference_golden_within_bound(oracle):
    """tests/golden/level3_golden.npz holds outputs of the reference's GENERIC build for every
    routine / precision / side / uplo / trans combination.  The oracle sums in its own order, so
    it must agree within c*K*eps*gauge; whatever the routine may not touch must be bit-identical."""
    import level3_helpers as L
    g = np.load(os.path.join(ROOT, "tests", "golden", "level3_golden.npz"))
    meta = g["meta"]
    assert len(meta) == 144
    worst = 0.0
    for idx, row in enumerate(meta):
        case = L.meta_case(row)
        a, b, c0, ref_c = g[f"a{idx}"], g[f"b{idx}"], g[f"c0_{idx}"], g[f"c{idx}"]
        _, want, gauge, K, touched = L.run_case(lambda *args: None, oracle, case, a, b if b.size else a, c0)
        worst = max(worst, L.check_case(case, ref_c, want, gauge, K, touched, c0))
    assert worst < 1.0


@pytest.mark.skipif(not cpu.have_reference("generic"), reason="oracle/_ref not built")
def test_level3_oracle_vs_live_reference(oracle):
    """Same comparison live, including the best SIMD build of the reference, bigger k, NaN in the
    triangle of A that must not be read and in the part of C that must not be touched."""
    import level3_helpers as L
    rng = np.random.default_rng(77)
    for ref in (cpu.Reference("generic"), cpu.Reference()):
        ref.set_threads(1)
        call = L.bind(ref.lib)
        for dtype in (cpu.S, cpu.D, cpu.CX, cpu.Z):
            cplx = dtype in (cpu.CX, cpu.Z)
            for herm in ((0, 1) if cplx else (0,)):
                for x in (0, 1):
                    for uplo in (0, 1):
                        m, n = 45, 38
                        ka = n if x else m
                        a, b, c0 = L.operand(rng, dtype, ka, ka + 1), L.operand(rng, dtype, n, m + 2), L.operand(rng, dtype, n, m + 3)
                        jj, ii = np.meshgrid(np.arange(ka), np.arange(ka + 1), indexing="ij")
                        a[(ii < jj) if uplo else ((ii > jj) & (ii < ka))] = np.nan      # the other triangle of A
                        alpha, beta = ((0.7 - 0.9j, 1.3 - 1.1j) if cplx else (0.7, 1.3))
                        case = (0, dtype, herm, x, uplo, 0, m, n, 0, ka + 1, m + 2, m + 3, alpha, beta)
                        got, want, gauge, K, touched = L.run_case(call, oracle, case, a, b, c0)
                        L.check_case(case, got, want, gauge, K, touched, c0)
                        for trans in (0, 1):
                            nn, k = 41, 130
                            rows, cols = (k, nn) if trans else (nn, k)
                            a, b, c0 = L.operand(rng, dtype, cols, rows + 1), L.operand(rng, dtype, cols, rows + 2), L.operand(rng, dtype, nn, nn + 3)
                            jj, ii = np.meshgrid(np.arange(nn), np.arange(nn + 3), indexing="ij")
                            c0[((ii < jj) if uplo else (ii > jj)) & (ii < nn)] = np.nan     # the other triangle of C
                            al = 0.7 if (herm and not x) or not cplx else 0.7 - 0.9j
                            be = 1.3 if herm or not cplx else 1.3 - 1.1j
                            case = (1, dtype, herm, x, uplo, trans, nn, nn, k, rows + 1, rows + 2, nn + 3, al, be)
                            got, want, gauge, K, touched = L.run_case(call, oracle, case, a, b, c0)
                            L.check_case(case, got, want, gauge, K, touched, c0)


def test_level3_argument_checks_match_reference_error_exits(oracle):
    """oracle_check_symm / oracle_check_rankk against the info values the REFERENCE reported for
    the column-major probes of tests/c/errexit_level3.c (tests/golden/errexit_level3_reference.txt)."""
    want = {}
    for ln in open(os.path.join(ROOT, "tests", "golden", "errexit_level3_reference.txt")):
        if ln.startswith("col ") and "info=" in ln:
            want[ln[4:34].strip()] = int(ln.rsplit("info=", 1)[1])
    assert oracle.check_symm(-1, 0, 0, 0, 1, 1, 1, -1) == want["dsymm side"] == 1
    assert oracle.check_symm(0, -1, 0, 0, 1, 1, 1, -1) == want["dsymm uplo"] == 2
    assert oracle.check_symm(0, 0, -1, 0, 1, 1, 1, -1) == want["dsymm m<0"]
    assert oracle.check_symm(1, 1, 0, -1, 1, 1, 1, -1) == want["dsymm n<0"]
    assert oracle.check_symm(0, 0, 2, 0, 1, 2, 2, -1) == want["dsymm lda left"]
    assert oracle.check_symm(1, 0, 0, 2, 1, 2, 2, -1) == want["dsymm lda right"]
    assert oracle.check_symm(0, 1, 2, 0, 2, 1, 2, -1) == want["dsymm ldb left m=2"]
    assert oracle.check_symm(0, 0, 2, 0, 2, 2, 1, -1) == want["dsymm ldc m=2"]
    assert oracle.check_symm(-1, 0, 2, 3, 2, 2, 1, -1) == want["dsymm side + others"]
    assert oracle.check_rankk(0, -1, 0, 0, 0, 1, 1, 1, -1) == want["dsyrk uplo"]
    assert oracle.check_rankk(0, 0, -1, 0, 0, 1, 1, 1, -1) == want["dsyrk trans"]
    assert oracle.check_rankk(0, 0, 0, -1, 0, 1, 1, 1, -1) == want["dsyrk n<0"]
    assert oracle.check_rankk(0, 1, 1, 0, -1, 1, 1, 1, -1) == want["dsyrk k<0"]
    assert oracle.check_rankk(0, 0, 0, 2, 0, 1, 1, 2, -1) == want["dsyrk lda N n=2"]
    assert oracle.check_rankk(0, 0, 1, 0, 2, 1, 1, 1, -1) == want["dsyrk lda T k=2"]
    assert oracle.check_rankk(0, 1, 0, 2, 0, 2, 2, 1, -1) == want["dsyrk ldc"] == 10
    assert oracle.check_rankk(1, 0, 1, 0, 2, 2, 1, 1, -1) == want["dsyr2k ldb T"] == 9
    assert oracle.check_rankk(1, 1, 0, 2, 0, 2, 2, 1, -1) == want["dsyr2k ldc"] == 12
    assert oracle.check_rankk(1, 0, 0, 2, 0, 1, 2, 2, -1) == want["dsyr2k lda"] == 7


def test_trxm_oracle_and_checker_accept_reference_golden(oracle):
    """tests/golden/trxm_golden.npz: the reference's TRMM results must sit within the bound of the
    oracle's, and its TRSM solutions must pass the oracle's restatement of ctest's residual check
    (the check is what pins TRSM: forward errors depend on conditioning, residuals do not)."""
    import level3_helpers as L
    g = np.load(os.path.join(ROOT, "tests", "golden", "trxm_golden.npz"))
    assert len(g["meta"]) == 192
    worst_mm = worst_sm = 0.0
    for idx, row in enumerate(g["meta"]):
        case = L.trxm_meta_case(row)
        r = L.check_trxm(oracle, None, case, g[f"a{idx}"], g[f"b0_{idx}"], ref_b=g[f"b{idx}"])
        if case[1]:
            worst_sm = max(worst_sm, r)
            # and the oracle's own substitution passes the same check
            mine = g[f"b0_{idx}"].copy()
            oracle.trxm(case[0], 1, case[2], case[3], case[4], case[5], case[6], case[7], case[10], g[f"a{idx}"], case[8], mine, case[9])
            L.check_trxm(oracle, None, case, g[f"a{idx}"], g[f"b0_{idx}"], ref_b=mine)
        else:
            worst_mm = max(worst_mm, r)
    assert worst_mm < 1.0 and worst_sm < L.TRSM_THRESH


def test_trxm_argument_checks_match_reference_error_exits(oracle):
    want = {}
    for ln in open(os.path.join(ROOT, "tests", "golden", "errexit_level3_reference.txt")):
        if ln.startswith("col ") and "info=" in ln:
            want[ln[4:34].strip()] = int(ln.rsplit("info=", 1)[1])
    assert oracle.check_trxm(-1, 0, 0, 0, 0, 0, 1, 1, -1) == want["dtrmm side"] == 1
    assert oracle.check_trxm(0, -1, 0, 0, 0, 0, 1, 1, -1) == want["dtrmm uplo"] == 2
    assert oracle.check_trxm(0, 0, -1, 0, 0, 0, 1, 1, -1) == want["dtrmm trans"] == 3
    assert oracle.check_trxm(0, 0, 0, -1, 0, 0, 1, 1, -1) == want["dtrmm diag"] == 4
    assert oracle.check_trxm(0, 0, 0, 0, -1, 0, 1, 1, -1) == want["dtrmm m<0"] == 5
    assert oracle.check_trxm(1, 1, 1, 1, 0, -1, 1, 1, -1) == want["dtrmm n<0"] == 6
    assert oracle.check_trxm(0, 0, 0, 0, 2, 0, 1, 2, -1) == want["dtrmm lda left m=2"] == 9
    assert oracle.check_trxm(1, 0, 0, 0, 0, 2, 1, 2, -1) == want["dtrmm lda right n=2"] == 9
    assert oracle.check_trxm(0, 1, 3, 0, 2, 0, 2, 1, -1) == want["dtrmm ldb m=2"] == 11
    assert oracle.check_trxm(-1, 0, 0, 0, 2, 3, 1, 1, -1) == want["dtrmm side + others"] == 1


@pytest.mark.skipif(not cpu.have_reference("generic"), reason="oracle/_ref not built")
def test_gemm_batch_reference_equals_small_matrix_oracle_bitwise(oracle):
    """the reference's cblas_dgemm_batch (interface/gemm_batch.c) sends matrices that pass
    GEMM_SMALL_MATRIX_PERMIT to the small-matrix kernels; per matrix its GENERIC build must equal the
    oracle's restatement of those kernels bit for bit, and sit within the bound of the blocked oracle."""
    import ctypes as C
    ref = cpu.Reference("generic")
    ref.set_threads(1)
    rng = np.random.default_rng(4)
    groups = [(0, 1, 12, 9, 20, 3), (1, 0, 40, 33, 8, 2), (0, 0, 64, 48, 130, 2)]
    alphas, betas = [0.7, 1.0, -0.4], [1.3, 0.0, 1.0]
    cb = {0: 111, 1: 112}
    probs = []
    for (ta, tb, m, n, k, cnt) in groups:
        for _ in range(cnt):
            a, lda, b, ldb, c0, ldc = problem(rng, oracle, cpu.D, ta, tb, m, n, k, pad=(1, 1, 1))
            probs.append((a, lda, b, ldb, c0, c0.copy(), ldc))
    ints = lambda v: (C.c_int * len(v))(*v)
    first = [sum(g[5] for g in groups[:i]) for i in range(len(groups))]
    ptrs = lambda j: (C.c_void_p * len(probs))(*[p[j].ctypes.data for p in probs])
    ref.lib.cblas_dgemm_batch(102, ints([cb[g[0]] for g in groups]), ints([cb[g[1]] for g in groups]), ints([g[2] for g in groups]),
                              ints([g[3] for g in groups]), ints([g[4] for g in groups]), (C.c_double * 3)(*alphas), ptrs(0),
                              ints([probs[f][1] for f in first]), ptrs(2), ints([probs[f][3] for f in first]),
                              (C.c_double * 3)(*betas), ptrs(5), ints([probs[f][6] for f in first]), len(groups),
                              ints([g[5] for g in groups]))
    i = 0
    for gi, (ta, tb, m, n, k, cnt) in enumerate(groups):
        for _ in range(cnt):
            a, lda, b, ldb, c0, got, ldc = probs[i]
            i += 1
            small = c0.copy()
            oracle.gemm(cpu.D, ta, tb, m, n, k, alphas[gi], a, lda, b, ldb, betas[gi], small, ldc, small=True)
            assert np.array_equal(small.view(np.uint8), got.view(np.uint8)), (gi, m, n, k)
            blocked = c0.copy()
            oracle.gemm(cpu.D, ta, tb, m, n, k, alphas[gi], a, lda, b, ldb, betas[gi], blocked, ldc)
            assert oracle.ratio(cpu.D, ta, tb, m, n, k, alphas[gi], a, lda, b, ldb, betas[gi], c0, ldc, got, ldc, blocked, ldc) <= C_BOUND
