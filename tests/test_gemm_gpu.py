"""GPU parity tests: the CUDA path, called through the C ABI of libopenblas_b200.so, against the
CPU oracle (oracle/gemm_oracle.c, bit-identical to the reference's GENERIC build) on the same
seeded inputs.  Acceptance is the reference's own: ctest's err/(eps*gauge) < 16
(ctest/c_dblat3.f:2198-2318, threshold ctest/din3:7) and the north-star componentwise bound
|C - C_ref| <= c*k*eps*(|alpha||A||B| + |beta||C|) with c = helpers.C_BOUND, eps = 2^-52 for
d/z and 2^-23 for s/c/sb."""
import os

import numpy as np
import pytest

from oracle import cpu
from helpers import (ALL_DTYPES, C_BOUND, NAMES, SBGEMM_ABS_TOL, THRESH, alpha_beta, ntrans, operand, problem,
                     stored_dims)

pytestmark = pytest.mark.gpu
CB = {0: 111, 1: 112, 2: 114, 3: 113}


def run_cblas(ob, dtype, order, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc):
    ob.cblas.GEMM[dtype](order, CB[ta], CB[tb], m, n, k, alpha, a, lda, b, ldb, beta, c, ldc)


def check(oracle, dtype, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c0, ldc, got, tag=""):
    want = c0.copy()
    oracle.gemm(dtype, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, want, ldc)
    err = oracle.mmch(dtype, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c0, ldc, got, ldc)
    ratio = oracle.ratio(dtype, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c0, ldc, got, ldc, want, ldc)
    ctx = (NAMES[dtype], ta, tb, m, n, k, alpha, beta, tag)
    assert err < THRESH, ("ctest ratio", err) + ctx
    assert ratio <= C_BOUND, ("componentwise", ratio) + ctx
    # only the m x n window may change: padding rows keep their rogue value bit for bit
    assert np.array_equal(got[:, m:].view(np.uint8), c0[:, m:].view(np.uint8)), ("padding",) + ctx
    return ratio


@pytest.mark.parametrize("dtype", ALL_DTYPES)
def test_golden_vectors(ob, oracle, golden, dtype):
    """The reference-generated fixtures, through the Fortran-ABI symbol."""
    meta = golden["meta"]
    seen = 0
    for idx, row in enumerate(meta):
        if int(row[0]) != dtype:
            continue
        _, ta, tb, m, n, k, lda, ldb, ldc = (int(v) for v in row[:9])
        cplx = dtype in (cpu.CX, cpu.Z)
        alpha = complex(row[9], row[10]) if cplx else row[9]
        beta = complex(row[11], row[12]) if cplx else row[11]
        a, b, c0 = golden[f"a{idx}"], golden[f"b{idx}"], golden[f"c0_{idx}"]
        got = c0.copy()
        ob.cblas.fortran_gemm(dtype, "NTRC"[ta], "NTRC"[tb], m, n, k, alpha, a, lda, b, ldb, beta, got, ldc)
        ratio = oracle.ratio(dtype, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c0, ldc, got, ldc,
                             golden[f"c{idx}"], ldc)
        assert ratio <= C_BOUND, (idx, ratio)
        assert np.array_equal(got[:, m:].view(np.uint8), c0[:, m:].view(np.uint8))
        seen += 1
    assert seen >= 8


@pytest.mark.parametrize("kernel", ["auto", "generic"])
@pytest.mark.parametrize("dtype", ALL_DTYPES)
def test_all_ops_ragged_shapes_colmajor(ob, oracle, dtype, kernel):
    """Every op combination x ragged shapes x the ctest alpha/beta grid, padded lds."""
    ob.cblas.set_kernel(ob.cblas.K_GENERIC if kernel == "generic" else ob.cblas.K_AUTO)
    try:
        rng = np.random.default_rng(1000 + dtype)
        alphas, betas = alpha_beta(dtype)
        shapes = [(1, 1, 1), (3, 5, 2), (35, 9, 7), (65, 63, 129), (128, 128, 64), (200, 77, 301), (257, 130, 96)]
        worst = 0.0
        for si, (m, n, k) in enumerate(shapes):
            for ta in range(ntrans(dtype)):
                for tb in range(ntrans(dtype)):
                    alpha = alphas[(si + ta + tb) % 3]
                    beta = betas[(si + 2 * ta + tb) % 3]
                    a, lda, b, ldb, c0, ldc = problem(rng, oracle, dtype, ta, tb, m, n, k)
                    got = c0.copy()
                    run_cblas(ob, dtype, ob.cblas.ColMajor, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, got, ldc)
                    worst = max(worst, check(oracle, dtype, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c0, ldc, got,
                                             ob.cblas.last_kernel()))
        print(NAMES[dtype], kernel, "worst componentwise ratio", worst)
    finally:
        ob.cblas.set_kernel(ob.cblas.K_AUTO)


@pytest.mark.parametrize("dtype", ALL_DTYPES)
def test_row_major_is_the_transposed_problem(ob, oracle, dtype):
    """interface/gemm.c:423-471.  Row-major C(m x n) with ldc >= n is the column-major n x m
    problem B'*A'; check against the oracle on that swapped problem."""
    rng = np.random.default_rng(2000 + dtype)
    for (m, n, k) in [(5, 7, 3), (33, 18, 40), (130, 65, 77)]:
        for ta in range(ntrans(dtype)):
            for tb in range(ntrans(dtype)):
                alpha, beta = alpha_beta(dtype)[0][2], alpha_beta(dtype)[1][2]
                # build the swapped column-major problem, then call row-major with roles restored
                a2, lda2, b2, ldb2, c0, ldc = problem(rng, oracle, dtype, tb, ta, n, m, k)
                got = c0.copy()
                run_cblas(ob, dtype, ob.cblas.RowMajor, ta, tb, m, n, k, alpha, b2, ldb2, a2, lda2, beta, got, ldc)
                check(oracle, dtype, tb, ta, n, m, k, alpha, a2, lda2, b2, ldb2, beta, c0, ldc, got, "rowmajor")


@pytest.mark.parametrize("dtype", ALL_DTYPES)
def test_beta_zero_and_alpha_zero_semantics(ob, oracle, dtype):
    """beta == 0 never reads C (NaN/Inf there vanish, kernel/generic/gemm_beta.c:52-71);
    alpha == 0 or k == 0 never reads A/B and only scales C (level3.c:229-259);
    alpha == 0 and beta == 1 leaves C bit-identical."""
    rng = np.random.default_rng(3000 + dtype)
    for (m, n, k) in [(9, 5, 7), (150, 140, 130)]:
        a, lda, b, ldb, c0, ldc = problem(rng, oracle, dtype, 0, 0, m, n, k)
        cn = c0.copy()
        cn[:, :m] = np.nan
        got = cn.copy()
        run_cblas(ob, dtype, ob.cblas.ColMajor, 0, 0, m, n, k, 1.0, a, lda, b, ldb, 0.0, got, ldc)
        assert np.all(np.isfinite(got[:n, :m])), "beta == 0 read C"
        check(oracle, dtype, 0, 0, m, n, k, 1.0, a, lda, b, ldb, 0.0, c0, ldc, got, "beta0")
        # alpha == 0 with NaN operands
        an = a.copy()
        an[:] = 0x7fc0 if dtype == cpu.SB else np.nan
        got = c0.copy()
        beta = alpha_beta(dtype)[1][2]
        run_cblas(ob, dtype, ob.cblas.ColMajor, 0, 0, m, n, k, 0.0, an, lda, b, ldb, beta, got, ldc)
        assert np.all(np.isfinite(got[:n, :m])), "alpha == 0 read A"
        check(oracle, dtype, 0, 0, m, n, k, 0.0, a, lda, b, ldb, beta, c0, ldc, got, "alpha0")
        # k == 0
        got = c0.copy()
        run_cblas(ob, dtype, ob.cblas.ColMajor, 0, 0, m, n, 0, 1.0, an, lda, b, ldb, beta, got, ldc)
        check(oracle, dtype, 0, 0, m, n, 0, 1.0, a, lda, b, ldb, beta, c0, ldc, got, "k0")
        # alpha == 0, beta == 1: untouched
        got = c0.copy()
        run_cblas(ob, dtype, ob.cblas.ColMajor, 0, 0, m, n, k, 0.0, a, lda, b, ldb, 1.0, got, ldc)
        assert np.array_equal(got.view(np.uint8), c0.view(np.uint8))


def test_inputs_are_not_modified(ob, oracle):
    """ctest LDE (c_dblat3.f:2320): A and B bitwise unchanged after the call."""
    rng = np.random.default_rng(5)
    for dtype in ALL_DTYPES:
        a, lda, b, ldb, c0, ldc = problem(rng, oracle, dtype, 1, 0, 40, 30, 20)
        a0, b0 = a.copy(), b.copy()
        run_cblas(ob, dtype, ob.cblas.ColMajor, 1, 0, 40, 30, 20, 1.0, a, lda, b, ldb, 0.0, c0, ldc)
        assert np.array_equal(a.view(np.uint8), a0.view(np.uint8))
        assert np.array_equal(b.view(np.uint8), b0.view(np.uint8))


@pytest.mark.parametrize("dtype", ALL_DTYPES)
def test_device_pointers_and_determinism(ob, oracle, dtype):
    """Device-resident operands are used in place; repeated calls are bit-identical
    (cpp_thread_test/dgemm_thread_safety.cpp:70-90 demands run-to-run equality)."""
    import torch
    rng = np.random.default_rng(4000 + dtype)
    m, n, k = 384, 320, 448
    for ta in range(ntrans(dtype)):
        tb = (ta + 1) % ntrans(dtype)
        a, lda, b, ldb, c0, ldc = problem(rng, oracle, dtype, ta, tb, m, n, k, pad=(0, 8, 16))
        alpha, beta = alpha_beta(dtype)[0][2], alpha_beta(dtype)[1][2]
        view = {np.uint16: np.int16}.get(a.dtype.type, a.dtype.type)
        da = torch.from_numpy(a.view(view)).cuda()
        db = torch.from_numpy(b.view(view)).cuda()
        outs = []
        for _ in range(3):
            dc = torch.from_numpy(c0.copy()).cuda()
            ob.cblas.gemm_any(dtype, ta, tb, m, n, k, alpha, da, lda, db, ldb, beta, dc, ldc)
            outs.append(dc.cpu().numpy())
        check(oracle, dtype, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c0, ldc, outs[0], "device")
        assert np.array_equal(outs[0].view(np.uint8), outs[1].view(np.uint8))
        assert np.array_equal(outs[0].view(np.uint8), outs[2].view(np.uint8))


def test_conj_notrans_matches_preconjugated(ob, oracle):
    """utest/test_extensions/test_zgemm.c: op 'R' must equal 'N' on a conjugated copy (1e-13)."""
    rng = np.random.default_rng(6)
    for dtype, tol in ((cpu.Z, 1e-13), (cpu.CX, 1e-4)):
        m = n = k = 100
        a, lda, b, ldb, c0, ldc = problem(rng, oracle, dtype, 0, 0, m, n, k, pad=(0, 0, 0))
        alpha, beta = -2 + 1j, 1 - 1j
        g1, g2 = c0.copy(), c0.copy()
        run_cblas(ob, dtype, ob.cblas.ColMajor, 2, 0, m, n, k, alpha, a, lda, b, ldb, beta, g1, ldc)
        run_cblas(ob, dtype, ob.cblas.ColMajor, 0, 0, m, n, k, alpha, np.conj(a), lda, b, ldb, beta, g2, ldc)
        assert np.max(np.abs(g1 - g2)) <= tol * max(1.0, np.max(np.abs(g2)))


def test_sbgemm_vs_sgemm_like_reference_test(ob, oracle):
    """test/compare_sgemm_sbgemm.c:99-203: sbgemm on bf16-rounded inputs within 1.0 of sgemm on
    the fp32 inputs and of a naive fp32 loop, inputs in (0.5, 1.5), NN/TN/NT/TT."""
    rng = np.random.default_rng(8)
    for x in (1, 7, 32, 100, 256):
        af = (rng.random((x, x), dtype=np.float32) + 0.5)
        bf = (rng.random((x, x), dtype=np.float32) + 0.5)
        ah, bh = oracle.tobf16(af), oracle.tobf16(bf)
        for ta in (0, 1):
            for tb in (0, 1):
                c = np.zeros((x, x), dtype=np.float32)
                cc = np.zeros((x, x), dtype=np.float32)
                ob.cblas.fortran_gemm(cpu.S, "NT"[ta], "NT"[tb], x, x, x, 1.0, af, x, bf, x, 0.0, c, x)
                ob.cblas.fortran_gemm(cpu.SB, "NT"[ta], "NT"[tb], x, x, x, 1.0, ah, x, bh, x, 0.0, cc, x)
                opa = oracle.bf16to(ah).T if not ta else oracle.bf16to(ah)
                opb = oracle.bf16to(bh).T if not tb else oracle.bf16to(bh)
                dd = (opa.astype(np.float32) @ opb.astype(np.float32)).T
                assert np.max(np.abs(cc - c)) <= SBGEMM_ABS_TOL
                assert np.max(np.abs(cc - dd)) <= SBGEMM_ABS_TOL


@pytest.mark.parametrize("dtype", [cpu.CX, cpu.Z])
def test_gemm3m_over_the_ctest_grid(ob, oracle, dtype):
    """The sweep of the reference's GEMM3M acceptance driver (ctest/c_zblat3_3m.f ZCHK1 with ctest/zin3_3m: every
    m, n, k in {0, 1, 2, 3, 5, 9, 35}, the nine N/T/C combinations, alpha and beta from the file's grids, both
    layouts), through cblas_?gemm3m, judged like the driver does (ZMMCH: err / (eps * gauge) < 16) plus the
    componentwise bound and untouched padding.  Row-major calls pass the transposed problem's storage."""
    rng = np.random.default_rng(33 + dtype)
    alphas, betas = alpha_beta(dtype)
    fn = ob.cblas.cgemm3m if dtype == cpu.CX else ob.cblas.zgemm3m
    sizes = (0, 1, 2, 3, 5, 9, 35)
    calls = 0
    for order in (ob.cblas.ColMajor, ob.cblas.RowMajor):
        for m in sizes:
            for n in sizes:
                for k in (sizes if order == ob.cblas.ColMajor else (0, 3, 35)):
                    for ta in (0, 1, 3):
                        for tb in (0, 1, 3):
                            ai, bi = rng.integers(0, 3), rng.integers(0, 3)
                            alpha, beta = alphas[ai], betas[bi]
                            if order == ob.cblas.ColMajor:
                                a, lda, b, ldb, c0, ldc = problem(rng, oracle, dtype, ta, tb, m, n, k)
                                got = c0.copy()
                                fn(order, CB[ta], CB[tb], m, n, k, alpha, a, lda, b, ldb, beta, got, ldc)
                                if m and n:
                                    check(oracle, dtype, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c0, ldc, got, "3m-col")
                                else:
                                    assert np.array_equal(got.view(np.uint8), c0.view(np.uint8))
                            else:
                                # row-major C (m x n) = column-major C^T (n x m) = op(B)^T op(A)^T: the column-major problem (tb, ta, n, m, k)
                                b_, ldb, a_, lda, c0, ldc = problem(rng, oracle, dtype, tb, ta, n, m, k)
                                got = c0.copy()
                                fn(order, CB[ta], CB[tb], m, n, k, alpha, a_, lda, b_, ldb, beta, got, ldc)
                                if m and n:
                                    check(oracle, dtype, tb, ta, n, m, k, alpha, b_, ldb, a_, lda, beta, c0, ldc, got, "3m-row")
                                else:
                                    assert np.array_equal(got.view(np.uint8), c0.view(np.uint8))
                            calls += 1
    assert calls == 7 * 7 * 7 * 9 + 7 * 7 * 3 * 9


def test_gemm3m_and_batch(ob, oracle):
    """f1/f2 of SURVEY 8(f): gemm3m has the GEMM contract; gemm_batch runs every matrix of every
    group (interface/gemm_batch.c:322-366)."""
    import ctypes as C
    rng = np.random.default_rng(9)
    m, n, k = 33, 21, 17
    a, lda, b, ldb, c0, ldc = problem(rng, oracle, cpu.Z, 3, 1, m, n, k)
    got = c0.copy()
    ob.cblas.zgemm3m(ob.cblas.ColMajor, CB[3], CB[1], m, n, k, 0.7 - 0.9j, a, lda, b, ldb, 1.3 - 1.1j, got, ldc)
    check(oracle, cpu.Z, 3, 1, m, n, k, 0.7 - 0.9j, a, lda, b, ldb, 1.3 - 1.1j, c0, ldc, got, "3m")

    # (transa, transb, m, n, k, count); the third group reaches the DMMA kernel, the fourth has beta == 0
    groups = [(0, 1, 12, 9, 20, 3), (1, 0, 40, 33, 8, 2), (0, 0, 128, 96, 160, 12), (1, 1, 7, 5, 3, 4)]
    alphas, betas = [0.7, 1.0, -0.4, 2.0], [1.3, 0.0, 1.0, 0.0]
    probs, ptr_a, ptr_b, ptr_c = [], [], [], []
    first_of = []
    for (ta, tb, gm, gn, gk, cnt) in groups:
        for _ in range(cnt):
            pa, plda, pb, pldb, pc0, pldc = problem(rng, oracle, cpu.D, ta, tb, gm, gn, gk, pad=(1, 1, 1))
            got_c = pc0.copy()
            if betas[len(first_of)] == 0.0:
                got_c[:, :gm] = np.nan          # beta == 0 must not read C
            probs.append((ta, tb, gm, gn, gk, pa, plda, pb, pldb, pc0, got_c, pldc))
        first_of.append(len(probs))
    I = lambda vals: (C.c_int * len(vals))(*vals)
    ta_arr, tb_arr = I([CB[g[0]] for g in groups]), I([CB[g[1]] for g in groups])
    m_arr, n_arr, k_arr = I([g[2] for g in groups]), I([g[3] for g in groups]), I([g[4] for g in groups])
    first = [sum(g[5] for g in groups[:i]) for i in range(len(groups))]
    lda_arr = I([probs[f][6] for f in first]); ldb_arr = I([probs[f][8] for f in first]); ldc_arr = I([probs[f][11] for f in first])
    alpha = (C.c_double * len(groups))(*alphas); beta = (C.c_double * len(groups))(*betas)
    A = (C.c_void_p * len(probs))(*[p[5].ctypes.data for p in probs])
    B = (C.c_void_p * len(probs))(*[p[7].ctypes.data for p in probs])
    Cc = (C.c_void_p * len(probs))(*[p[10].ctypes.data for p in probs])
    gs = I([g[5] for g in groups])
    adr = lambda x: C.cast(x, C.c_void_p)
    ob.lib().cblas_dgemm_batch(ob.cblas.ColMajor, adr(ta_arr), adr(tb_arr), adr(m_arr), adr(n_arr), adr(k_arr),
                               adr(alpha), adr(A), adr(lda_arr), adr(B), adr(ldb_arr), adr(beta), adr(Cc),
                               adr(ldc_arr), len(groups), adr(gs))
    i = 0
    for gi, (ta, tb, gm, gn, gk, cnt) in enumerate(groups):
        for _ in range(cnt):
            p = probs[i]; i += 1
            check(oracle, cpu.D, ta, tb, gm, gn, gk, alpha[gi], p[5], p[6], p[7], p[8], beta[gi], p[9], p[11], p[10], "batch")


def test_bf16_helpers_match_reference_vectors(ob):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "bf16_golden.npz"))
    x = g["x"]
    h = np.zeros(x.size, dtype=np.uint16)
    ob.lib().cblas_sbstobf16(x.size, x.ctypes.data, 1, h.ctypes.data, 1)
    assert np.array_equal(h, g["bf16"])
    back = np.zeros(x.size, dtype=np.float32)
    ob.lib().cblas_sbf16tos(x.size, h.ctypes.data, 1, back.ctypes.data, 1)
    assert np.array_equal(back.view(np.uint32), g["back"].view(np.uint32))
    # strided + negative increment (interface/tobf16.c: pointer moved to the far end)
    h2 = np.full(2 * 50, 0xdead, dtype=np.uint16)
    ob.lib().cblas_sbstobf16(50, x.ctypes.data, 1, h2.ctypes.data, -2)
    assert np.array_equal(h2[::2][::-1], g["bf16"][:50]) and np.all(h2[1::2] == 0xdead)
    # one element per call, the way test/compare_sgemm_sbgemm.c converts its matrices (the in-place pinned path), all four directions
    import ctypes as C
    L = ob.lib()
    for i in range(0, x.size, max(1, x.size // 64)):
        one, one_back, d_in, d_out, d_back = np.zeros(1, np.uint16), np.zeros(1, np.float32), x[i:i + 1].astype(np.float64), np.zeros(1, np.uint16), np.zeros(1, np.float64)
        L.cblas_sbstobf16(1, C.c_void_p(x[i:i + 1].ctypes.data), 1, C.c_void_p(one.ctypes.data), 1)
        L.cblas_sbf16tos(1, C.c_void_p(one.ctypes.data), 1, C.c_void_p(one_back.ctypes.data), 1)
        L.cblas_sbdtobf16(1, C.c_void_p(d_in.ctypes.data), 1, C.c_void_p(d_out.ctypes.data), 1)
        L.cblas_dbf16tod(1, C.c_void_p(d_out.ctypes.data), 1, C.c_void_p(d_back.ctypes.data), 1)
        assert one[0] == g["bf16"][i] and one_back.view(np.uint32)[0] == g["back"].view(np.uint32)[i]
        assert d_out[0] == g["bf16"][i] and np.float32(d_back[0]).view(np.uint32) == g["back"].view(np.uint32)[i]


def test_concurrent_callers_get_identical_results(ob, oracle):
    """cpp_thread_test/dgemm_thread_safety.cpp: concurrent cblas_dgemm on identical inputs, every
    thread's C equal to thread 0's."""
    import threading
    rng = np.random.default_rng(10)
    m = n = k = 256
    a, lda, b, ldb, c0, ldc = problem(rng, oracle, cpu.D, 0, 0, m, n, k, pad=(0, 0, 0))
    outs = [c0.copy() for _ in range(8)]

    def work(i):
        for _ in range(3):
            o = c0.copy()
            run_cblas(ob, cpu.D, ob.cblas.ColMajor, 0, 0, m, n, k, 1.0, a, lda, b, ldb, 0.1, o, ldc)
            outs[i] = o
    ts = [threading.Thread(target=work, args=(i,)) for i in range(8)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for o in outs[1:]:
        assert np.array_equal(o.view(np.uint8), outs[0].view(np.uint8))
    check(oracle, cpu.D, 0, 0, m, n, k, 1.0, a, lda, b, ldb, 0.1, c0, ldc, outs[0], "threads")


def test_summa_schedule_on_one_gpu(ob, oracle):
    """The SUMMA panel loop (openblas_b200/summa.py: double-buffered panel copies on a side stream,
    event hand-offs, per-panel DGEMM with beta applied only at the first panel) on a 1x1 grid must
    equal one DGEMM of the whole problem."""
    import torch
    from openblas_b200 import summa
    dev = torch.device("cuda", 0)
    m, n, k, nb = 700, 520, 900, 128
    g = torch.Generator(device=dev); g.manual_seed(3)
    A = torch.rand((k, m), generator=g, device=dev, dtype=torch.float64) - 0.5
    B = torch.rand((n, k), generator=g, device=dev, dtype=torch.float64) - 0.5
    C0 = torch.rand((n, m), generator=g, device=dev, dtype=torch.float64) - 0.5
    grid = summa.Grid(1, 1, 0, 0, 0, None, None, [0], [0])
    gemm = lambda mm, nn, kk, al, a, lda, b, ldb, be, c, ldc, st: ob.cblas.gemm_device(cpu.D, 0, 0, mm, nn, kk, al, a, lda, b, ldb, be, c, ldc, st)
    sm = summa.Summa(grid, m, n, k, nb, torch.float64, dev, gemm)
    got = C0.clone()
    for _ in range(2):          # second sweep reuses the buffers / events
        got.copy_(C0)
        sm.run(0.7, A, B, 1.3, got)
    torch.cuda.synchronize()
    a, b, c0 = A.cpu().numpy(), B.cpu().numpy(), C0.cpu().numpy()
    check(oracle, cpu.D, 0, 0, m, n, k, 0.7, a, m, b, k, 1.3, c0, m, got.cpu().numpy(), "summa-1gpu")
    assert sm.launches == 2 * ((k + nb - 1) // nb)


@pytest.mark.parametrize("dtype", [cpu.D, cpu.S, cpu.Z, cpu.CX, cpu.SB])
def test_tile_aligned_shapes_take_the_roofline_kernels(ob, oracle, dtype):
    """Shapes that are multiples of every kernel's tile with 16-byte friendly leading dimensions:
    the paths the benchmarks run (DGEMM: the bulk-copy/mbarrier variant), all op combinations,
    beta != 0, several C tiles per CTA and several ring wraps."""
    import torch
    rng = np.random.default_rng(5000 + dtype)
    shapes = [(256, 512, 256), (384, 768, 416)]     # kept small: the CPU oracle is O(mnk) scalar code
    cplx = dtype in (cpu.CX, cpu.Z)
    if cplx:                                         # complex tiles are half as wide: the same tile counts at half the extents
        shapes = [(128, 256, 256), (256, 384, 416)]
    for si, (m, n, k) in enumerate(shapes):
        for ta in range(ntrans(dtype)):
            for tb in range(ntrans(dtype)):
                if cplx and si == 1 and (tb - ta) % 4 != 1:
                    continue                         # the larger complex shape: a Latin quarter of the 16 op pairs (every op of A and of B once)
                a, lda, b, ldb, c0, ldc = problem(rng, oracle, dtype, ta, tb, m, n, k, pad=(8, 16, 24))
                alpha, beta = alpha_beta(dtype)[0][2], alpha_beta(dtype)[1][2]
                view = {np.uint16: np.int16}.get(a.dtype.type, a.dtype.type)
                da, db = torch.from_numpy(a.view(view)).cuda(), torch.from_numpy(b.view(view)).cuda()
                dc = torch.from_numpy(c0.copy()).cuda()
                ob.cblas.gemm_any(dtype, ta, tb, m, n, k, alpha, da, lda, db, ldb, beta, dc, ldc)
                kern = ob.cblas.last_kernel()
                assert "generic" not in kern, kern
                check(oracle, dtype, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c0, ldc, dc.cpu().numpy(), kern)
    if dtype == cpu.D:
        assert "bulk" in kern or "pw" in kern, kern


@pytest.mark.parametrize("tile", ["64", "128"])
@pytest.mark.parametrize("dtype", [cpu.D, cpu.S])
def test_both_tile_sizes(ob, oracle, dtype, tile, monkeypatch):
    """The DGEMM and SGEMM launchers pick the 64x64 or the 128x128 kernel from the grid's wave
    count; force each one over interior, ragged and k-tail shapes of every op combination."""
    import torch
    monkeypatch.setenv("B200_DGEMM_TILE" if dtype == cpu.D else "B200_SGEMM_TILE", tile)
    rng = np.random.default_rng(77 + int(tile) + dtype)
    for (m, n, k) in [(192, 320, 160), (203, 141, 75), (64, 64, 64), (130, 66, 33)]:
        for ta in range(2):
            for tb in range(2):
                ra, rb = (k if ta else m), (n if tb else k)       # even leading dimensions: 16-byte aligned columns (DGEMM)
                a, lda, b, ldb, c0, ldc = problem(rng, oracle, dtype, ta, tb, m, n, k, pad=(2 + ra % 2, 4 + rb % 2, 5))
                da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
                dc = torch.from_numpy(c0.copy()).cuda()
                ob.cblas.gemm_any(dtype, ta, tb, m, n, k, -0.6, da, lda, db, ldb, 0.8, dc, ldc)
                kern = ob.cblas.last_kernel()
                assert ("64x64" in kern) == (tile == "64"), kern
                check(oracle, dtype, ta, tb, m, n, k, -0.6, a, lda, b, ldb, 0.8, c0, ldc, dc.cpu().numpy(), kern)


@pytest.mark.parametrize("ops", [(0, 0), (1, 1)])
def test_pipelined_host_path_matches_device_path(ob, oracle, ops):
    """Big host-pointer calls go through the panel pipeline of runtime.cu (H2D of A row panels / B
    column panels, block GEMMs and D2H of C blocks overlapped).  The result must be bit-identical to
    the same GEMM on device-resident copies (same kernel, same k order per element), padding rows
    of the host C untouched, and sampled entries must match a float64 dot product."""
    import torch
    ta, tb = ops
    rng = np.random.default_rng(77)
    m, n, k = 4100, 4300, 1000
    ra, ca = (k, m) if ta else (m, k)
    rb, cb = (n, k) if tb else (k, n)
    lda, ldb, ldc = ra + 4, rb + 2, m + 6
    a = rng.random((ca, lda)) - 0.5
    b = rng.random((cb, ldb)) - 0.5
    c0 = rng.random((n, ldc)) - 0.5
    c0[:, m:] = -1e10
    alpha, beta = 0.7, 1.3
    got = c0.copy()
    ob.cblas.dgemm(ob.cblas.ColMajor, CB[ta], CB[tb], m, n, k, alpha, a, lda, b, ldb, beta, got, ldc)
    da, db, dc = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), torch.from_numpy(c0.copy()).cuda()
    ob.cblas.gemm_any(cpu.D, ta, tb, m, n, k, alpha, da, lda, db, ldb, beta, dc, ldc)
    dev = dc.cpu().numpy()
    # interior 2048-blocks run the same kernel family on the same k order; edge blocks may take another
    # kernel (e.g. the 4-row sliver goes to the generic one), so compare to rounding, not bit for bit
    assert np.max(np.abs(got[:, :m] - dev[:, :m])) <= 1e-13 * k
    assert np.all(got[:, m:] == -1e10)
    for (i, j) in [(0, 0), (m - 1, n - 1), (2047, 2048), (2048, 2047), (4099, 17), (1234, 4299)]:
        row = a[i, :k] if ta else a[:k, i]
        col = b[:k, j] if tb else b[j, :k]
        want = alpha * float(np.dot(row, col)) + beta * c0[j, i]
        assert abs(got[j, i] - want) <= 1e-12 * k


def test_midsize_pageable_host_operands(ob, oracle):
    """Between the packed small path (<= 4 MiB) and the panel pipeline: pageable numpy operands go
    through the threaded pinned-slot staging of runtime.cu (h2d_any / d2h_any), padded lds."""
    rng = np.random.default_rng(88)
    m, n, k = 1500, 900, 700
    for ta, tb in ((0, 0), (1, 0)):
        a, lda, b, ldb, c0, ldc = problem(rng, oracle, cpu.D, ta, tb, m, n, k, pad=(3, 5, 7))
        got = c0.copy()
        run_cblas(ob, cpu.D, ob.cblas.ColMajor, ta, tb, m, n, k, 0.7, a, lda, b, ldb, 1.3, got, ldc)
        check(oracle, cpu.D, ta, tb, m, n, k, 0.7, a, lda, b, ldb, 1.3, c0, ldc, got, "pageable-mid")


@pytest.mark.parametrize("dtype", [cpu.Z, cpu.CX])
def test_gemm3m_large_products_use_three_real_gemms(ob, oracle, dtype):
    """?gemm3m with every extent >= 512: split3 of both operands, three REAL GEMMs on the roofline kernels, combine3
    (runtime.cu: gemm3m_on_device).  Op pairs that cover N / T / R / C on each side, beta != 0 and beta == 0 over a NaN
    C, device and pageable host operands.  Judged like the reference's 3M acceptance driver (ctest/c_zblat3c_3m.c:
    err / (eps * gauge) < 16 with gauge = |alpha| sum (|re| + |im|)(|re| + |im|) + |beta| (|re| + |im|) of C), against a
    complex128 product; 2 + 3 + 1 launches; the plain ?gemm entry point keeps the 4-multiply kernel."""
    import ctypes as C
    import torch
    lib = ob.lib()
    rng = np.random.default_rng(3300 + dtype)
    eps = 2.0 ** -52 if dtype == cpu.Z else 2.0 ** -23
    ct = np.complex128 if dtype == cpu.Z else np.complex64
    name = "cblas_" + cpu.DTYPE_NAMES[dtype] + "gemm3m"
    m, n, k = 640, 768, 512
    abs1 = lambda z: np.abs(z.real) + np.abs(z.imag)
    for ta, tb, where in ((0, 0, "device"), (1, 2, "device"), (3, 1, "host"), (2, 3, "device")):
        ra, ca = (k, m) if ta & 1 else (m, k)
        rb, cb = (n, k) if tb & 1 else (k, n)
        lda, ldb, ldc = ra + 8, rb + 4, m + 6
        a = ((rng.random((ca, lda)) - 0.5) + 1j * (rng.random((ca, lda)) - 0.5)).astype(ct)
        b = ((rng.random((cb, ldb)) - 0.5) + 1j * (rng.random((cb, ldb)) - 0.5)).astype(ct)
        c0 = ((rng.random((n, ldc)) - 0.5) + 1j * (rng.random((n, ldc)) - 0.5)).astype(ct)
        # numpy holds the column-major matrices transposed: element (i, j) is x[j, i]
        opx = lambda x, rows, t: {0: x[:, :rows].T, 1: x[:, :rows], 2: x[:, :rows].T.conj(), 3: x[:, :rows].conj()}[t]
        X, Y = opx(a.astype(np.complex128), ra, ta), opx(b.astype(np.complex128), rb, tb)          # m x k, k x n
        for alpha, beta in ((0.7 - 0.9j, 1.3 - 1.1j), (1.0 + 0.0j, 0.0j)):
            ref = alpha * (X @ Y) + beta * c0[:, :m].T.astype(np.complex128)
            gauge = abs1(np.complex128(alpha)) * (abs1(X) @ abs1(Y)) + abs1(np.complex128(beta)) * abs1(c0[:, :m].T.astype(np.complex128))
            start = c0.copy()
            if beta == 0:
                start[:, :m] = np.nan
            al, be = np.array([alpha], dtype=ct), np.array([beta], dtype=ct)
            before = ob.cblas.launch_count()
            if where == "device":
                da, db, dc = (torch.from_numpy(np.ascontiguousarray(x).view(np.float64 if dtype == cpu.Z else np.float32)).cuda() for x in (a, b, start))
                getattr(lib, name)(102, cpu.CBLAS_TRANS[ta], cpu.CBLAS_TRANS[tb], m, n, k, C.c_void_p(al.ctypes.data), C.c_void_p(da.data_ptr()), lda,
                                   C.c_void_p(db.data_ptr()), ldb, C.c_void_p(be.ctypes.data), C.c_void_p(dc.data_ptr()), ldc)
                got = dc.cpu().numpy().view(ct).reshape(start.shape)
            else:
                got = start.copy()
                getattr(lib, name)(102, cpu.CBLAS_TRANS[ta], cpu.CBLAS_TRANS[tb], m, n, k, C.c_void_p(al.ctypes.data), C.c_void_p(a.ctypes.data), lda,
                                   C.c_void_p(b.ctypes.data), ldb, C.c_void_p(be.ctypes.data), C.c_void_p(got.ctypes.data), ldc)
            assert ob.cblas.last_kernel() == "combine3", ob.cblas.last_kernel()
            assert ob.cblas.launch_count() - before == 6
            ratio = (abs1(got[:, :m].T.astype(np.complex128) - ref) / (eps * gauge)).max()
            assert ratio < 16.0, (ta, tb, where, alpha, ratio)
            assert np.array_equal(got[:, m:].view(np.uint8), start[:, m:].view(np.uint8)), "padding rows changed"
    # the plain entry point is untouched by all this
    a2, b2, c2 = ((rng.random((k, m)) - 0.5) + 0j).astype(ct), ((rng.random((n, k)) - 0.5) + 0j).astype(ct), np.zeros((n, m), dtype=ct)
    run_cblas(ob, dtype, ob.cblas.ColMajor, 0, 0, m, n, k, 1.0, a2, m, b2, k, 0.0, c2, m)
    assert ob.cblas.last_kernel() != "combine3"
