"""BASELINE.json configurations at FULL size on the GPU, checked through size-independent
properties because a CPU GEMM of these sizes takes minutes to hours:

  * sampled entries of C against extended-precision dot products of the same rows/columns
    (the long-double reference the survey used), within the north-star bound
    |C - C_ref| <= c*k*eps*(|alpha||A||B| + |beta||C|), c = 2;
  * linearity: GEMM(alpha, A, B, beta, C) == alpha*GEMM(1, A, B, 0, .) + beta*C to rounding;
  * padding columns of C (ldc > m) keep their bits.

Configs: (2) DGEMM / SGEMM 16384^3 NN, (3) SBGEMM 8192^3, (5) ZGEMM / CGEMM tall-skinny
65536 x 256 x 65536 with a conjugate-transposed A and beta != 0 (A alone is 64 GiB / 32 GiB)."""
import numpy as np
import pytest

from oracle import cpu

pytestmark = pytest.mark.gpu
C_BOUND = 2.0


def _rand(torch, shape, dtype, gen, dev):
    if dtype.is_complex:
        base = torch.float64 if dtype == torch.complex128 else torch.float32
        return torch.view_as_complex(torch.rand(shape + (2,), generator=gen, device=dev, dtype=base) - 0.5)
    if dtype == torch.bfloat16:
        return (torch.rand(shape, generator=gen, device=dev, dtype=torch.float32) - 0.5).to(torch.bfloat16)
    return torch.rand(shape, generator=gen, device=dev, dtype=dtype) - 0.5


def _sampled_check(torch, ob, code, ta, tb, m, n, k, alpha, beta, tdt, odt, eps, pad=0, nsamples=24):
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev); gen.manual_seed(code * 1000 + m % 997)
    ra, ca = (k, m) if ta & 1 else (m, k)
    rb, cb = (n, k) if tb & 1 else (k, n)
    lda, ldb, ldc = ra, rb, m + pad
    a = _rand(torch, (ca, lda), tdt, gen, dev)
    b = _rand(torch, (cb, ldb), tdt, gen, dev)
    c0 = _rand(torch, (n, ldc), odt, gen, dev)
    if pad:
        c0[:, m:] = -1e10
    c = c0.clone()
    ob.cblas.gemm_any(code, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc)
    torch.cuda.synchronize()
    kern = ob.cblas.last_kernel()
    assert "generic" not in kern, kern
    if pad:
        assert torch.equal(c[:, m:], c0[:, m:]), "padding rows of C changed"
    rng = np.random.default_rng(m + n + k)
    ii = np.concatenate([[0, m - 1], rng.integers(0, m, nsamples)])
    jj = np.concatenate([[0, n - 1], rng.integers(0, n, nsamples)])
    worst = 0.0
    wide = np.clongdouble if tdt.is_complex else np.longdouble
    for i, j in zip(ii, jj):
        i, j = int(i), int(j)
        row = (a[i, :k] if ta & 1 else a[:k, i]).to(torch.complex128 if tdt.is_complex else torch.float64).cpu().numpy().astype(wide)
        col = (b[:k, j] if tb & 1 else b[j, :k]).to(torch.complex128 if tdt.is_complex else torch.float64).cpu().numpy().astype(wide)
        if ta & 2:
            row = np.conj(row)
        if tb & 2:
            col = np.conj(col)
        old = complex(c0[j, i].item()) if odt.is_complex else float(c0[j, i].item())
        got = complex(c[j, i].item()) if odt.is_complex else float(c[j, i].item())
        want = alpha * np.dot(row, col) + (beta * old if beta != 0 else 0)
        gauge = abs(alpha) * float(np.dot(np.abs(row), np.abs(col))) + abs(beta) * abs(old)
        worst = max(worst, float(abs(got - want)) / (k * eps * gauge))
    print(f"{kern}: {m}x{n}x{k} op={ta}{tb} worst sampled ratio {worst:.4f}")
    assert worst <= C_BOUND, (kern, worst)
    return a, b, c0, c, lda, ldb, ldc


def test_dgemm_16384_config2(ob):
    import torch
    m = n = k = 16384
    a, b, c0, c, lda, ldb, ldc = _sampled_check(torch, ob, cpu.D, 0, 0, m, n, k, 1.0, 0.0, torch.float64, torch.float64, 2.0 ** -52)
    # linearity on a column slab: alpha*AB + beta*C0 vs alpha*(AB) + beta*C0 recomputed from c (= AB)
    alpha, beta, w = 0.7, 1.3, 512
    c2 = c0[:w].clone()
    ob.cblas.gemm_any(cpu.D, 0, 0, m, w, k, alpha, a, lda, b, ldb, beta, c2, ldc)
    want = alpha * c[:w] + beta * c0[:w]
    scale = float(want.abs().max())
    assert float((c2 - want).abs().max()) <= 64 * k * 2.0 ** -52 * max(1.0, scale)


def test_sgemm_16384_config2(ob):
    import torch
    _sampled_check(torch, ob, cpu.S, 0, 0, 16384, 16384, 16384, 1.0, 0.0, torch.float32, torch.float32, 2.0 ** -23)


def test_sbgemm_8192_config3(ob):
    import torch
    for ta, tb in ((0, 0), (1, 0), (0, 1), (1, 1)):
        _sampled_check(torch, ob, cpu.SB, ta, tb, 8192, 8192, 8192, 1.0, 0.0, torch.bfloat16, torch.float32, 2.0 ** -23)


def test_zgemm_tall_skinny_config5(ob):
    import torch
    free, _ = torch.cuda.mem_get_info()
    if free < 80 * 2 ** 30:
        pytest.skip("needs ~66 GiB of free HBM")
    # op(A) = A^H (stored k x m), beta != 0, padded ldc
    _sampled_check(torch, ob, cpu.Z, 3, 0, 65536, 256, 65536, 0.7 - 0.9j, 1.3 - 1.1j, torch.complex128, torch.complex128,
                   2.0 ** -52, pad=2, nsamples=10)


def test_cgemm_tall_skinny_config5(ob):
    import torch
    _sampled_check(torch, ob, cpu.CX, 0, 3, 65536, 256, 65536, 0.7 - 0.9j, 1.3 - 1.1j, torch.complex64, torch.complex64,
                   2.0 ** -23, pad=2, nsamples=10)
    # the other tall-skinny orientation: C is the big operand (65536 x 65536 complex64 = 32 GiB)
    _sampled_check(torch, ob, cpu.CX, 3, 0, 65536, 65536, 256, 0.7 - 0.9j, 1.3 - 1.1j, torch.complex64, torch.complex64,
                   2.0 ** -23, pad=0, nsamples=10)
