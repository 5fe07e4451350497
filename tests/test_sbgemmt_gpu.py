"""GPU parity tests (through the C ABI) of SBGEMMT (interface/sbgemmt.c): one launch of the tcgen05
SBGEMM kernel walking only the tiles of the triangle with a masked epilogue (sbgemm_tcgen05.cu), or the
block-column scheme, against oracle_sbgemmt (pinned bit for bit to the reference by tests/test_sbgemmt_pin.py) and
the reference's own outputs in tests/golden/sbgemmt_golden.npz.  Tolerance: |got - want| <= (k + 2) * 2^-23 *
(|alpha| sum |a||b| + |beta||c|), the north star's fp32-accumulation bound (bf16 products are exact in fp32)."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import cpu
import test_sbgemmt_pin as P

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def check(got, want, gauge, mask, start, m, k, what):
    inside = mask[:, :m]
    d = np.abs(got.astype(np.float64) - want.astype(np.float64))[:, :m]
    assert not np.isnan(got[:, :m][inside]).any(), (what, "NaN inside the triangle")
    ratio = (d[inside] / ((k + 2) * P.EPS32 * np.maximum(gauge[:m, :m][inside], 1e-300))).max() if inside.any() else 0.0
    assert ratio <= 1.0, (what, ratio)
    assert P.same_bits(got[~mask], start[~mask]), (what, "bytes outside the triangle changed")


def test_reference_golden_vectors(ob, oracle):
    lib = ob.lib()
    g = np.load(os.path.join(ROOT, "tests", "golden", "sbgemmt_golden.npz"))
    for i in range(int(g["count"][0])):
        key = f"case{i}"
        uplo, ta, tb, m, k, lda, ldb, ldc, cblas, rowmajor = (int(v) for v in g[key + "_meta"])
        alpha, beta = (float(v) for v in g[key + "_scal"])
        a, b, c0 = g[key + "_a"], g[key + "_b"], g[key + "_c0"]
        got = c0.copy()
        cpu.call_sbgemmt(lib, uplo, ta, tb, m, k, alpha, a, lda, b, ldb, beta, got, ldc, cblas=bool(cblas), rowmajor=bool(rowmajor))
        _, gauge, mask = P.expected(oracle, uplo, ta, tb, m, k, alpha, a, lda, b, ldb, beta, c0, ldc, rowmajor)
        check(got, g[key + "_c"], gauge, mask, c0, m, k, key)
        if k == 0:
            assert P.same_bits(got, c0)


@pytest.mark.parametrize("tri_env", ["1", "0"])
def test_all_ops_both_schemes_host_and_device(ob, oracle, tri_env, monkeypatch):
    """m = 300 and 700 (ragged against 128 x 256 and 256 x 256 tiles; 700 runs on CTA pairs), every op combination,
    both triangles, both ABIs and orders; NaN outside the triangle and in the padding rows; beta == 0 over a NaN
    triangle; alpha == 0; host operands (odd leading dimensions: the repack path) and device operands."""
    import torch
    monkeypatch.setenv("B200_RANKK_TRI", tri_env)
    lib = ob.lib()
    rng = np.random.default_rng(311)
    kernels = set()
    for m, k in ((300, 132), (700, 70)):
        for uplo in (0, 1):
            for ta in range(4):
                for tb in range(4):
                    if m == 700 and (ta > 1 or tb > 1):
                        continue                      # R / C fold to N / T: covered at 300
                    cblas, rowmajor = ((False, False), (True, False), (True, True))[(ta + 2 * tb + uplo) % 3]
                    a, lda, b, ldb, c0, ldc = P.problem(rng, oracle, ta, tb, m, k, rowmajor, pad=(5, 8, 4))
                    mask = P.tri_mask(m, ldc, uplo)
                    c0[~mask] = np.nan
                    for alpha, beta in ((0.7, 1.3), (1.0, 0.0), (0.0, 0.5)):
                        start = c0.copy()
                        if beta == 0.0:
                            start[mask] = np.nan
                        want, gauge, _ = P.expected(oracle, uplo, ta, tb, m, k, alpha, a, lda, b, ldb, beta, start, ldc, rowmajor)
                        got = start.copy()
                        cpu.call_sbgemmt(lib, uplo, ta, tb, m, k, alpha, a, lda, b, ldb, beta, got, ldc, cblas=cblas, rowmajor=rowmajor)
                        check(got, want, gauge, mask, start, m, k, ("host", tri_env, m, uplo, ta, tb, cblas, rowmajor, alpha))
                        if alpha != 0.0:
                            kernels.add(ob.cblas.last_kernel())
                        if (ta + tb + uplo) % 2 == 0:
                            da, db = torch.from_numpy(a.view(np.int16)).cuda(), torch.from_numpy(b.view(np.int16)).cuda()
                            dc = torch.from_numpy(start.copy()).cuda()
                            cpu.call_sbgemmt(lib, uplo, ta, tb, m, k, alpha, da.data_ptr(), lda, db.data_ptr(), ldb, beta, dc.data_ptr(), ldc,
                                             cblas=cblas, rowmajor=rowmajor)
                            torch.cuda.synchronize()
                            check(dc.cpu().numpy(), want, gauge, mask, start, m, k, ("device", tri_env, m, uplo, ta, tb, cblas, rowmajor, alpha))
    if tri_env == "1":
        assert any("tcgen05" in kn for kn in kernels), kernels


def test_large_triangle_equals_the_triangle_of_sbgemm_in_one_launch(ob, oracle):
    """4096 x 4096 x 2048 on the device: the triangle must equal the same triangle of the full SBGEMM (same kernel,
    same k order), the other triangle must keep its bytes, a sample of elements is checked against the oracle's
    arithmetic in float64, and the launch count says it was ONE tcgen05 launch."""
    import torch
    lib = ob.lib()
    m, k = 4096, 2048
    rng = np.random.default_rng(5)
    a32 = (rng.random((k, m), dtype=np.float32) - 0.5)
    b32 = (rng.random((m, k), dtype=np.float32) - 0.5)
    a, b = oracle.tobf16(a32), oracle.tobf16(b32)                 # column-major m x k (lda = m) and k x m (ldb = k)
    da, db = torch.from_numpy(a.view(np.int16)).cuda(), torch.from_numpy(b.view(np.int16)).cuda()
    c0 = torch.from_numpy((rng.random((m, m), dtype=np.float32) - 0.5)).cuda()
    full = c0.clone()
    ob.cblas.sbgemm(102, 111, 111, m, m, k, 0.7, da, m, db, k, 1.3, full, m)
    for uplo in (0, 1):
        tri = c0.clone()
        before = ob.cblas.launch_count()
        cpu.call_sbgemmt(lib, uplo, 0, 0, m, k, 0.7, da.data_ptr(), m, db.data_ptr(), k, 1.3, tri.data_ptr(), m)
        torch.cuda.synchronize()
        assert ob.cblas.launch_count() - before == 1 and "tcgen05" in ob.cblas.last_kernel()
        # torch tensors are row-major views of the column-major buffer: element (i, j) of C sits at [j, i]
        keep = torch.ones(m, m, dtype=torch.bool, device="cuda")
        keep = torch.tril(keep) if uplo == 0 else torch.triu(keep)     # [j, i] with i <= j  <=>  upper triangle of C
        assert torch.equal(tri[~keep].view(torch.int32), c0[~keep].view(torch.int32))
        assert torch.equal(tri[keep].view(torch.int32), full[keep].view(torch.int32))
        got = tri.cpu().numpy()
        af, bf = oracle.bf16to(a).astype(np.float64), oracle.bf16to(b).astype(np.float64)
        c0h = c0.cpu().numpy().astype(np.float64)
        for _ in range(200):
            i, j = int(rng.integers(0, m)), int(rng.integers(0, m))
            if (i > j) if uplo == 0 else (i < j):
                i, j = j, i
            col_a, col_b = af[:, i], bf[j, :]                          # op(A)(i, :) and op(B)(:, j)
            exact = 0.7 * float(col_a @ col_b) + 1.3 * c0h[j, i]
            gauge = 0.7 * float(np.abs(col_a) @ np.abs(col_b)) + 1.3 * abs(c0h[j, i])
            assert abs(got[j, i] - exact) <= (k + 2) * P.EPS32 * gauge, (uplo, i, j)


def test_error_exits_print_the_reference_message(ob, oracle, capfd):
    lib = ob.lib()
    buf = np.full(64, 7.0, dtype=np.float32)
    p = buf.ctypes.data_as(C.c_void_p)
    U = {-1: 0, 0: 121, 1: 122}; T = {-1: 0, 0: 111, 1: 112}
    probes = 0
    for rowmajor in (0, 1):
        for (uplo, ta, tb, m, k, lda, ldb, ldc) in ((-1, 0, 0, 2, 2, 2, 2, 2), (0, -1, 0, 2, 2, 2, 2, 2), (0, 0, -1, 2, 2, 2, 2, 2), (0, 0, 0, -1, 2, 2, 2, 2),
                                                    (0, 0, 0, 2, -1, 2, 2, 2), (0, 0, 0, 3, 2, 2, 3, 3), (0, 1, 0, 3, 4, 3, 4, 3), (0, 0, 0, 3, 4, 3, 3, 3),
                                                    (0, 0, 1, 3, 2, 3, 2, 3), (0, 0, 0, 3, 2, 3, 2, 2), (0, 1, 1, 3, 5, 2, 4, 3), (0, 1, 1, 3, 5, 5, 2, 3)):
            # row-major: the swapped problem, uplo NOT flipped (sbgemmt.c:239-240)
            want = oracle.check_sbgemmt(1, uplo, tb, ta, m, k, ldb, lda, ldc, -1) if rowmajor else oracle.check_sbgemmt(0, uplo, ta, tb, m, k, lda, ldb, ldc, -1)
            if want < 0:
                continue
            probes += 1
            capfd.readouterr()
            lib.cblas_sbgemmt(101 if rowmajor else 102, U[uplo], T[ta], T[tb], C.c_int(m), C.c_int(k), C.c_float(1.0), p, C.c_int(lda), p, C.c_int(ldb),
                              C.c_float(0.0), p, C.c_int(ldc))
            C.CDLL(None).fflush(None)
            out = "".join(capfd.readouterr())
            assert f"SBGEMMT  parameter number {want:2d}" in out, (rowmajor, uplo, ta, tb, m, k, lda, ldb, ldc, want, out)
    assert probes >= 18 and (buf == 7.0).all()
