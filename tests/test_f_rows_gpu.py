"""GPU parity tests (through the C ABI) of the SURVEY 8 f3 / f4 routines added in round 2: ?gemmt (one
triangle-masked GEMM launch, or the block-column scheme), sbgemv and sbdot (bf16_level12.cu), against the
restatements in oracle/level3_oracle.c (pinned to the reference by tests/test_f_rows_pin.py) and against the
reference's own outputs in tests/golden/f_rows_golden.npz."""
import os

import numpy as np
import pytest

from oracle import cpu
import test_f_rows_pin as F

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EPS32 = 2.0 ** -23


def test_gemmt_reference_golden_vectors(ob, oracle):
    lib = ob.lib()
    g = np.load(os.path.join(ROOT, "tests", "golden", "f_rows_golden.npz"))
    for i in range(int(g["gemmt_count"][0])):
        key = f"gemmt{i}"
        dtype, uplo, ta, tb, m, k, lda, ldb, ldc, cblas, rowmajor = F.colmajor_problem(g[key + "_meta"])
        cplx = dtype in (cpu.CX, cpu.Z)
        alpha, beta = ((0.7 - 0.9j, 1.3 - 1.1j) if cplx else (0.7, 1.3))
        a, b, c0 = g[key + "_a"], g[key + "_b"], g[key + "_c0"]
        got = c0.copy()
        cpu.call_gemmt(lib, dtype, uplo, ta, tb, m, k, alpha, a, lda, b, ldb, beta, got, ldc, cblas=bool(cblas), rowmajor=bool(rowmajor))
        want, gauge, mask = F.gemmt_expected(oracle, dtype, uplo, ta, tb, m, k, alpha, a, lda, b, ldb, beta, c0, ldc, rowmajor)
        F.check_gemmt(dtype, m, k, ldc, got, g[key + "_c"], gauge, mask, c0, key)


@pytest.mark.parametrize("tri_env", ["1", "0"])
@pytest.mark.parametrize("dtype", [cpu.S, cpu.D, cpu.CX, cpu.Z])
def test_gemmt_all_ops_both_schemes_host_and_device(ob, oracle, dtype, tri_env, monkeypatch):
    """m = 300 (several tiles, ragged), every op combination, both triangles; NaN outside the triangle and in the
    padding rows (never read, bytes unchanged); beta == 0 over a NaN triangle; alpha == 0; host and device operands.
    B200_RANKK_TRI=1: one masked launch of the roofline kernel; =0: block columns + tri_merge."""
    import torch
    monkeypatch.setenv("B200_RANKK_TRI", tri_env)
    lib = ob.lib()
    rng = np.random.default_rng(100 + dtype)
    cplx = dtype in (cpu.CX, cpu.Z)
    kernels = set()
    for uplo in (0, 1):
        for ta in range(4 if cplx else 2):
            for tb in range(4 if cplx else 2):
                m, k = 300, 132
                ra, ca = (k, m) if ta & 1 else (m, k)
                rb, cb = (m, k) if tb & 1 else (k, m)
                lda, ldb, ldc = ra + 4, rb + 4, m + 4
                a, b, c0 = F.operand(rng, dtype, ca, lda), F.operand(rng, dtype, cb, ldb), F.operand(rng, dtype, m, ldc)
                mask = F.tri_mask(m, ldc, uplo)
                c0[~mask] = np.nan
                for alpha, beta in (((0.7 - 0.9j, 1.3 - 1.1j) if cplx else (0.7, 1.3)), (1.0, 0.0), (0.0, 0.5)):
                    start = c0.copy()
                    if beta == 0.0:
                        start[mask] = np.nan
                    want, gauge, mk = F.gemmt_expected(oracle, dtype, uplo, ta, tb, m, k, alpha, a, lda, b, ldb, beta, start, ldc, False)
                    got = start.copy()
                    cpu.call_gemmt(lib, dtype, uplo, ta, tb, m, k, alpha, a, lda, b, ldb, beta, got, ldc)
                    F.check_gemmt(dtype, m, k, ldc, got, want, gauge, mk, start, ("host", uplo, ta, tb, alpha))
                    kernels.add(ob.cblas.last_kernel())
                    if (ta + tb + uplo) % 2 == 0:
                        da, db, dc = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), torch.from_numpy(start.copy()).cuda()
                        cpu.call_gemmt(lib, dtype, uplo, ta, tb, m, k, alpha, da.data_ptr(), lda, db.data_ptr(), ldb, beta, dc.data_ptr(), ldc,
                                       cblas=True)
                        F.check_gemmt(dtype, m, k, ldc, dc.cpu().numpy(), want, gauge, mk, start, ("device", uplo, ta, tb, alpha))
    if tri_env == "1":      # the masked launch ran on a roofline kernel, not on the generic one
        assert any(("dmma" in kn or "ffma" in kn or "ws_tma" in kn) for kn in kernels), kernels


def test_gemmt_error_exits_match_the_oracle_table(ob, oracle, capfd):
    """illegal arguments: the library's own xerbla_ prints the reference's message with the position the restated
    checks give; C is not touched"""
    import ctypes as C
    lib = ob.lib()
    buf = np.full(64, 7.0)
    U = {-1: 0, 0: 121, 1: 122}; T = {-1: 0, 0: 111, 1: 112}
    probes = 0
    for rowmajor in (0, 1):
        for (uplo, ta, tb, m, k, lda, ldb, ldc) in ((-1, 0, 0, 2, 2, 2, 2, 2), (0, -1, 0, 2, 2, 2, 2, 2), (0, 0, -1, 2, 2, 2, 2, 2), (0, 0, 0, -1, 2, 2, 2, 2),
                                                    (0, 0, 0, 2, -1, 2, 2, 2), (0, 0, 0, 3, 2, 2, 3, 3), (0, 1, 0, 3, 4, 3, 4, 3), (0, 0, 0, 3, 4, 3, 3, 3),
                                                    (0, 0, 1, 3, 2, 3, 2, 3), (0, 0, 0, 3, 2, 3, 2, 2), (0, 0, 0, 3, 2, 2, 1, 2), (0, 1, 1, 3, 5, 2, 4, 3),
                                                    (0, 0, 0, 3, 2, 3, 1, 3), (0, 1, 1, 3, 5, 5, 2, 3)):
            if rowmajor:
                want = oracle.check_gemmt(1, (1 - uplo) if uplo >= 0 else -1, tb, ta, m, k, ldb, lda, ldc, -1)
            else:
                want = oracle.check_gemmt(0, uplo, ta, tb, m, k, lda, ldb, ldc, -1)
            if want < 0:
                continue                      # legal in this layout (the table is written for column-major)
            probes += 1
            capfd.readouterr()
            lib.cblas_dgemmt(101 if rowmajor else 102, U[uplo], T[ta], T[tb], C.c_int(m), C.c_int(k), C.c_double(1.0), buf.ctypes.data_as(C.c_void_p),
                             C.c_int(lda), buf.ctypes.data_as(C.c_void_p), C.c_int(ldb), C.c_double(0.0), buf.ctypes.data_as(C.c_void_p), C.c_int(ldc))
            C.CDLL(None).fflush(None)
            out = "".join(capfd.readouterr())
            assert f"DGEMMT  parameter number {want:2d}" in out, (rowmajor, uplo, ta, tb, m, k, lda, ldb, ldc, want, out)
    assert probes >= 20 and (buf == 7.0).all()


def test_sbgemv_sbdot_reference_golden_vectors(ob, oracle):
    """the reference's own outputs; the GPU sums in a different order: within 2 (len + 2) eps32 gauge"""
    lib = ob.lib()
    g = np.load(os.path.join(ROOT, "tests", "golden", "f_rows_golden.npz"))
    for i in range(int(g["sbgemv_count"][0])):
        key = f"sbgemv{i}"
        trans, m, n, lda, incx, incy = (int(v) for v in g[key + "_meta"])
        alpha, beta = (float(v) for v in g[key + "_ab"])
        want = g[key + "_y0"].copy()
        gauge = oracle.sbgemv(trans, m, n, alpha, g[key + "_a"], lda, g[key + "_x"], incx, beta, want, incy)
        assert np.array_equal(want, g[key + "_y"])
        for cblas in (False, True):
            y = g[key + "_y0"].copy()
            cpu.call_sbgemv(lib, trans, m, n, alpha, g[key + "_a"], lda, g[key + "_x"], incx, beta, y, incy, cblas=cblas)
            leny, lenx = (n, m) if trans else (m, n)
            step = abs(incy)
            touched = y[::step][::-1] if incy < 0 else y[::step]
            ref = want[::step][::-1] if incy < 0 else want[::step]
            assert (np.abs(touched[:leny] - ref[:leny]) <= 2 * (lenx + 2) * EPS32 * np.maximum(gauge[:leny], 1e-30)).all(), (key, cblas)
            if step > 1:      # the gaps between the elements of y keep their bits
                gap = np.ones(y.size, bool); gap[::step] = False
                assert np.array_equal(y[gap], g[key + "_y0"][gap])
    for i in range(int(g["sbdot_count"][0])):
        key = f"sbdot{i}"
        n, incx, incy = (int(v) for v in g[key + "_meta"])
        _, gauge = oracle.sbdot(n, g[key + "_x"], incx, g[key + "_y"], incy)
        for cblas in (False, True):
            d = cpu.call_sbdot(lib, n, g[key + "_x"], incx, g[key + "_y"], incy, cblas=cblas)
            assert abs(d - float(g[key + "_d"][0])) <= 2 * (n + 2) * EPS32 * max(gauge, 1e-30), (key, cblas, d)


@pytest.mark.parametrize("where", ["host", "device"])
def test_sbgemv_sbdot_large_and_unaligned(ob, oracle, where):
    """shapes that use every slice / pair path of bf16_level12.cu: tall, wide, odd lda (scalar loads), strided and
    reversed vectors, row-major; device operands use the caller's increments in place; run-to-run identical"""
    import torch
    lib = ob.lib()
    rng = np.random.default_rng(55)
    for (m, n, pad) in ((5000, 300, 0), (300, 5000, 1), (1, 4096, 0), (4097, 1, 3), (2048, 2048, 8)):
        for trans in (0, 1):
            for incx, incy in ((1, 1), (2, -3)):
                lda = m + pad
                lenx, leny = (m, n) if trans else (n, m)
                a = oracle.tobf16(rng.random((n, lda), dtype=np.float32) - 0.5)
                x = oracle.tobf16(rng.random(1 + (lenx - 1) * abs(incx), dtype=np.float32) - 0.5)
                y0 = (rng.random(1 + (leny - 1) * abs(incy)) - 0.5).astype(np.float32)
                for alpha, beta in ((0.7, 1.3), (1.0, 0.0)):
                    start = y0.copy()
                    if beta == 0.0:
                        start[::abs(incy)] = np.nan                      # beta == 0 never reads y
                    want = start.copy()
                    gauge = oracle.sbgemv(trans, m, n, alpha, a, lda, x, incx, beta, want, incy)
                    outs = []
                    for _ in range(2):
                        if where == "device":
                            da, dx, dy = torch.from_numpy(a.view(np.int16)).cuda(), torch.from_numpy(x.view(np.int16)).cuda(), torch.from_numpy(start.copy()).cuda()
                            cpu.call_sbgemv(lib, trans, m, n, alpha, da.data_ptr(), lda, dx.data_ptr(), incx, beta, dy.data_ptr(), incy, cblas=True)
                            got = dy.cpu().numpy()
                        else:
                            got = start.copy()
                            cpu.call_sbgemv(lib, trans, m, n, alpha, a, lda, x, incx, beta, got, incy)
                        outs.append(got)
                    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32)), "not deterministic"
                    got = outs[0]
                    sel = slice(None, None, abs(incy))
                    g_ord = gauge[::-1] if incy < 0 else gauge
                    assert (np.abs(got[sel] - want[sel]) <= 2 * (lenx + 2) * EPS32 * np.maximum(g_ord, 1e-30)).all(), (m, n, pad, trans, incx, incy, alpha)
                    gap = np.ones(got.size, bool); gap[sel] = False
                    assert np.array_equal(got[gap], start[gap])
    for n in (1, 255, 4096, 1_000_003):
        for incx, incy in ((1, 1), (3, -2)):
            x = oracle.tobf16(rng.random(1 + (n - 1) * abs(incx), dtype=np.float32) - 0.5)
            y = oracle.tobf16(rng.random(1 + (n - 1) * abs(incy), dtype=np.float32) - 0.5)
            want, gauge = oracle.sbdot(n, x, incx, y, incy)
            if where == "device":
                dx, dy = torch.from_numpy(x.view(np.int16)).cuda(), torch.from_numpy(y.view(np.int16)).cuda()
                d = [cpu.call_sbdot(lib, n, dx.data_ptr(), incx, dy.data_ptr(), incy, cblas=True) for _ in range(2)]
            else:
                d = [cpu.call_sbdot(lib, n, x, incx, y, incy) for _ in range(2)]
            assert d[0] == d[1] and abs(d[0] - want) <= 2 * (n + 2) * EPS32 * max(gauge, 1e-30), (n, incx, incy, d, want)
