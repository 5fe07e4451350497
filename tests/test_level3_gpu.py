"""GPU parity tests of the symmetric level-3 family through the C ABI:
SYMM/HEMM, SYRK/HERK, SYR2K/HER2K for s, d, c, z against the CPU oracle
(oracle/level3_oracle.c) and the reference-generated golden outputs, within
|x - y| <= 2 * K * eps * gauge; everything the routine may not touch must keep its bits."""
import os

import numpy as np
import pytest

from oracle import cpu
import level3_helpers as L

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_level3_golden_vectors(ob, oracle):
    """All 144 reference-generated cases: our result against the REFERENCE's stored result."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "level3_golden.npz"))
    call = L.bind(ob.lib())
    worst = 0.0
    for idx, row in enumerate(g["meta"]):
        case = L.meta_case(row)
        a, b, c0, ref_c = g[f"a{idx}"], g[f"b{idx}"], g[f"c0_{idx}"], g[f"c{idx}"]
        got, _, gauge, K, touched = L.run_case(call, oracle, case, a, b if b.size else a, c0)
        worst = max(worst, L.check_case(case, got, ref_c, gauge, K, touched, c0))
    assert worst < 1.0


@pytest.mark.parametrize("rankk_tri", ["1", "0"])
@pytest.mark.parametrize("dtype", [cpu.S, cpu.D, cpu.CX, cpu.Z])
def test_level3_all_flag_combinations(ob, oracle, dtype, rankk_tri, monkeypatch):
    """Both rank-k schemes (B200_RANKK_TRI=1: the triangle masked inside the GEMM kernel, one launch per
    product; =0: block columns with merged diagonal blocks).  Sizes that cross the block-column width of the rank-k path (128) and reach the fast GEMM
    kernels; NaN in the triangle of A that must not be read and in the triangle of C that must
    not be touched; beta == 0 over a NaN-filled triangle (C must not be read)."""
    monkeypatch.setenv("B200_RANKK_TRI", rankk_tri)
    call = L.bind(ob.lib())
    rng = np.random.default_rng(300 + dtype)
    cplx = dtype in (cpu.CX, cpu.Z)
    for herm in ((0, 1) if cplx else (0,)):
        for x in (0, 1):
            for uplo in (0, 1):
                m, n = 150, 70
                ka = n if x else m
                a, b, c0 = L.operand(rng, dtype, ka, ka + 1), L.operand(rng, dtype, n, m + 2), L.operand(rng, dtype, n, m + 3)
                jj, ii = np.meshgrid(np.arange(ka), np.arange(ka + 1), indexing="ij")
                a[(ii < jj) if uplo else ((ii > jj) & (ii < ka))] = np.nan
                alpha, beta = ((0.7 - 0.9j, 1.3 - 1.1j) if cplx else (0.7, 1.3))
                case = (0, dtype, herm, x, uplo, 0, m, n, 0, ka + 1, m + 2, m + 3, alpha, beta)
                got, want, gauge, K, touched = L.run_case(call, oracle, case, a, b, c0)
                L.check_case(case, got, want, gauge, K, touched, c0)
                for trans in (0, 1):
                    for (nn, k, beta_zero) in [(300, 90, False), (140, 33, True)]:
                        rows, cols = (k, nn) if trans else (nn, k)
                        pa, pb = (2, 4) if nn == 300 else (1, 2)      # even leading dimensions at 300: DGEMM's masked kernel needs 16-byte columns
                        a, b = L.operand(rng, dtype, cols, rows + pa), L.operand(rng, dtype, cols, rows + pb)
                        c0 = L.operand(rng, dtype, nn, nn + 3)
                        jj, ii = np.meshgrid(np.arange(nn), np.arange(nn + 3), indexing="ij")
                        if beta_zero:
                            c0[:, :nn] = np.nan                                   # nothing of C may be read
                        else:
                            c0[((ii < jj) if uplo else (ii > jj)) & (ii < nn)] = np.nan
                        al = 0.7 if (herm and not x) or not cplx else 0.7 - 0.9j
                        be = 0.0 if beta_zero else (1.3 if herm or not cplx else 1.3 - 1.1j)
                        case = (1, dtype, herm, x, uplo, trans, nn, nn, k, rows + pa, rows + pb, nn + 3, al, be)
                        got, want, gauge, K, touched = L.run_case(call, oracle, case, a, b, c0)
                        L.check_case(case, got, want, gauge, K, touched, c0)
                        assert not np.isnan(got[touched]).any()
                        if nn == 300:
                            kern = ob.cblas.last_kernel()
                            assert (kern in ("tri_merge",)) == (rankk_tri == "0"), kern


def test_level3_scaling_only_paths(ob, oracle):
    """alpha == 0 and k == 0: only the triangle is scaled; beta == 1 leaves C bit-identical;
    HERK zeroes the diagonal's imaginary part when it scales (zherkf.f:207-230)."""
    call = L.bind(ob.lib())
    rng = np.random.default_rng(5)
    n = 37
    for dtype, herm in [(cpu.D, 0), (cpu.Z, 1), (cpu.CX, 0)]:
        for (alpha, k, beta) in [(0.0, 9, 1.3), (0.7, 0, 1.3), (0.0, 9, 1.0), (0.7, 0, 0.0)]:
            for uplo in (0, 1):
                a, c0 = L.operand(rng, dtype, max(k, 1), n + 1), L.operand(rng, dtype, n, n + 2)
                case = (1, dtype, herm, 0, uplo, 0, n, n, k, n + 1, n + 1, n + 2, alpha, beta)
                got, want, gauge, K, touched = L.run_case(call, oracle, case, a, a, c0)
                if beta == 1.0:
                    assert np.array_equal(got.view(np.uint8), c0.view(np.uint8))
                else:
                    L.check_case(case, got, want, gauge, K, touched, c0)
                    if herm:
                        assert np.all(np.diagonal(got[:n, :n]).imag == 0.0)
    # SYMM with alpha == 0: C = beta * C over the whole m x n window, beta == 0 writes exact zeros over NaN
    a, b, c0 = L.operand(rng, cpu.D, 20, 21), L.operand(rng, cpu.D, 30, 21), L.operand(rng, cpu.D, 30, 22)
    c0[:, :20] = np.nan
    case = (0, cpu.D, 0, 0, 0, 0, 20, 30, 0, 21, 21, 22, 0.0, 0.0)
    got, want, gauge, K, touched = L.run_case(call, oracle, case, a, b, c0)
    assert np.all(got[:, :20] == 0.0) and np.array_equal(got[:, 20:].view(np.uint8), c0[:, 20:].view(np.uint8))


def test_level3_cblas_row_major_and_device_pointers(ob, oracle):
    """Row-major CBLAS entry points (flipped side / uplo / trans, HER2K's conjugated alpha) against
    the column-major oracle on the transposed problem, and device-resident operands."""
    import ctypes as C
    import torch
    lib = ob.lib()
    rng = np.random.default_rng(11)
    # row-major ZHER2K, trans = NoTrans: C (n x n) += alpha A B^H + conj(alpha) B A^H with A, B n x k row-major.
    n, k = 45, 52
    A = (rng.random((n, k)) - 0.5 + 1j * (rng.random((n, k)) - 0.5))
    B = (rng.random((n, k)) - 0.5 + 1j * (rng.random((n, k)) - 0.5))
    C0 = (rng.random((n, n)) - 0.5 + 1j * (rng.random((n, n)) - 0.5))
    alpha, beta = 0.7 - 0.9j, 1.3
    got = C0.copy()
    al = np.array([alpha.real, alpha.imag])
    lib.cblas_zher2k(101, 121, 111, n, k, al.ctypes.data_as(C.c_void_p), A.ctypes.data_as(C.c_void_p), k,
                     B.ctypes.data_as(C.c_void_p), k, C.c_double(beta), got.ctypes.data_as(C.c_void_p), n)
    full = alpha * A @ B.conj().T + np.conj(alpha) * B @ A.conj().T + beta * C0
    up = np.triu(np.ones((n, n), dtype=bool))
    assert np.allclose(got[up], np.where(np.eye(n, dtype=bool), full.real + 0j, full)[up], rtol=0, atol=1e-12)
    assert np.array_equal(got[~up], C0[~up])                      # CblasUpper in row-major: the lower part is untouched
    # row-major DSYMM, Side = Right, Uplo = Lower: C (m x n) = alpha B S + beta C
    m, n = 33, 41
    S = rng.random((n, n)) - 0.5
    S = np.tril(S) + np.tril(S, -1).T
    Bm, Cm = rng.random((m, n)) - 0.5, rng.random((m, n)) - 0.5
    Sl = np.tril(S) + np.triu(np.full((n, n), np.nan), 1)         # only the lower triangle is valid
    got = Cm.copy()
    lib.cblas_dsymm(101, 142, 122, m, n, C.c_double(0.7), Sl.ctypes.data_as(C.c_void_p), n, Bm.ctypes.data_as(C.c_void_p), n,
                    C.c_double(1.3), got.ctypes.data_as(C.c_void_p), n)
    assert np.allclose(got, 0.7 * Bm @ S + 1.3 * Cm, rtol=0, atol=1e-12)
    # device pointers: DSYRK on torch tensors, n large enough for several block columns and the DMMA kernel
    n, k = 1100, 200
    a = torch.rand((k, n), dtype=torch.float64, device="cuda") - 0.5          # column-major n x k, lda = n
    c = torch.rand((n, n), dtype=torch.float64, device="cuda") - 0.5
    c_before = c.clone()
    one = lambda v: C.byref(C.c_int(v))
    al, be = C.c_double(0.7), C.c_double(1.3)
    lib.dsyrk_(C.c_char_p(b"L"), C.c_char_p(b"N"), one(n), one(k), C.byref(al), C.c_void_p(a.data_ptr()), one(n), C.byref(be),
               C.c_void_p(c.data_ptr()), one(n))
    torch.cuda.synchronize()
    An = a.cpu().numpy().T                                         # n x k
    want = 0.7 * An @ An.T + 1.3 * c_before.cpu().numpy().T
    got = c.cpu().numpy().T
    low = np.tril(np.ones((n, n), dtype=bool))
    assert np.allclose(got[low], want[low], rtol=0, atol=1e-11)
    assert np.array_equal(got[~low], c_before.cpu().numpy().T[~low])
    assert "dgemm_dmma" in ob.cblas.last_kernel() or "tri_merge" in ob.cblas.last_kernel()


def test_trxm_golden_vectors(ob, oracle):
    """All 192 reference-generated TRMM / TRSM cases (every side / uplo / trans / diag combination)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "trxm_golden.npz"))
    worst = 0.0
    for idx, row in enumerate(g["meta"]):
        case = L.trxm_meta_case(row)
        worst = max(worst, L.check_trxm(oracle, ob.lib(), case, g[f"a{idx}"], g[f"b0_{idx}"]))
    assert worst < L.TRSM_THRESH


@pytest.mark.parametrize("dtype", [cpu.S, cpu.D, cpu.CX, cpu.Z])
def test_trxm_recursive_sizes(ob, oracle, dtype):
    """Triangles of 150 and 203 rows: three / four levels of the recursive split, ragged 64-blocks at
    the end, the coupling GEMMs on the fast kernels; every side / uplo / trans / diag combination."""
    rng = np.random.default_rng(900 + dtype)
    cplx = dtype in (cpu.CX, cpu.Z)
    alpha = (0.7 - 0.9j) if cplx else 0.7
    for solve in (0, 1):
        for side in (0, 1):
            for uplo in (0, 1):
                for trans in range(4 if cplx else 2):
                    for unit in (0, 1):
                        m, n = (150, 90) if (trans + unit) % 2 == 0 else (70, 203)
                        ka = n if side else m
                        a = L.tri_operand(rng, dtype, ka, ka + 1, uplo, unit)
                        if solve:      # keep the triangular system well conditioned at this size: scale the off-diagonal part
                            off = ~np.eye(ka, ka + 1, dtype=bool)
                            a[off] *= 4.0 / ka
                        b0 = L.operand(rng, dtype, n, m + 2)
                        case = (dtype, solve, side, uplo, trans, unit, m, n, ka + 1, m + 2, alpha)
                        L.check_trxm(oracle, ob.lib(), case, a, b0)


def test_trsm_device_pointers_large(ob):
    """DTRSM on device-resident operands, 2048 x 1024, left lower: residual against a float64 product."""
    import ctypes as C
    import torch
    lib = ob.lib()
    m, n = 2048, 1024
    a = torch.rand((m, m), dtype=torch.float64, device="cuda") - 0.5          # column-major view: a[j, i] = A(i, j)
    a = a * (4.0 / m) + torch.eye(m, dtype=torch.float64, device="cuda") * 1.5
    b = torch.rand((n, m), dtype=torch.float64, device="cuda") - 0.5
    b0 = b.clone()
    one = lambda v: C.byref(C.c_int(v))
    al = C.c_double(0.7)
    lib.dtrsm_(C.c_char_p(b"L"), C.c_char_p(b"L"), C.c_char_p(b"N"), C.c_char_p(b"N"), one(m), one(n), C.byref(al),
               C.c_void_p(a.data_ptr()), one(m), C.c_void_p(b.data_ptr()), one(m))
    torch.cuda.synchronize()
    A = torch.tril(a.T)                                                       # A(i, j), lower
    X = b.T
    res = (A @ X - 0.7 * b0.T).abs().max().item()
    assert res < 1e-12, res


def test_symm_hemm_panel_scheme_on_the_gpu(tmp_path):
    """SYMM / HEMM with the full-expansion limit forced to zero (B200_SYMM_FULL_MB=0, read once per process, hence the
    child process): the symmetric operand is expanded and multiplied in 256-wide panels along the inner dimension,
    both sides, both triangles, NaN in the triangle that must not be read; against the oracle."""
    import subprocess, sys, textwrap
    script = tmp_path / "symm_panels.py"
    script.write_text(textwrap.dedent(f"""
        import sys
        sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})
        import numpy as np
        import openblas_b200 as ob
        from oracle import cpu
        import level3_helpers as L
        oracle = cpu.Oracle()
        call = L.bind(ob.lib())
        rng = np.random.default_rng(4)
        n_exp = 0
        for dtype, herm in ((cpu.D, 0), (cpu.Z, 1), (cpu.CX, 0)):
            cplx = dtype in (cpu.CX, cpu.Z)
            for x in (0, 1):
                for uplo in (0, 1):
                    m, n = 600, 520
                    ka = n if x else m
                    a, b, c0 = L.operand(rng, dtype, ka, ka + 1), L.operand(rng, dtype, n, m + 2), L.operand(rng, dtype, n, m + 3)
                    jj, ii = np.meshgrid(np.arange(ka), np.arange(ka + 1), indexing="ij")
                    a[(ii < jj) if uplo else ((ii > jj) & (ii < ka))] = np.nan
                    alpha, beta = ((0.7 - 0.9j, 1.3 - 1.1j) if cplx else (0.7, 1.3))
                    case = (0, dtype, herm, x, uplo, 0, m, n, 0, ka + 1, m + 2, m + 3, alpha, beta)
                    before = ob.cblas.launch_count()
                    got, want, gauge, K, touched = L.run_case(call, oracle, case, a, b, c0)
                    L.check_case(case, got, want, gauge, K, touched, c0)
                    assert ob.cblas.launch_count() - before == 2 * ((ka + 255) // 256), (ob.cblas.launch_count() - before, ka)
                    n_exp += 1
        print("PANELS OK", n_exp)
    """))
    env = dict(os.environ, B200_SYMM_FULL_MB="0", B200_SYMM_PANEL_MB="1")
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "PANELS OK 12" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])
