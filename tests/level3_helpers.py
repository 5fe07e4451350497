"""Shared by the CPU pin tests and the GPU parity tests of the symmetric level-3 family."""
import numpy as np

from oracle import cpu

EPS = {cpu.S: 2.0 ** -23, cpu.D: 2.0 ** -52, cpu.CX: 2.0 ** -23, cpu.Z: 2.0 ** -52}
C_BOUND = 2.0     # |x - y| <= C_BOUND * K * eps * gauge, K = length of the sums (+2 for alpha and beta)


def operand(rng, dtype, cols, ld):
    x = rng.random((cols, ld)) - 0.5
    if dtype in (cpu.CX, cpu.Z):
        x = x + 1j * (rng.random((cols, ld)) - 0.5)
    return x.astype(cpu.NP_IN[dtype])


def run_case(lib_call, oracle, case, a, b, c0):
    """Runs one golden-style case through `lib_call` (cpu.call_symm / cpu.call_rankk bound to a
    library) and through the oracle; returns (got, want, gauge, K, touched mask over (n, ldc))."""
    kind, dtype, herm, x, uplo, trans, m, n, k, lda, ldb, ldc, alpha, beta = case
    got, want = c0.copy(), c0.copy()
    if kind == 0:
        lib_call("symm", dtype, herm, x, uplo, m, n, alpha, a, lda, b, ldb, beta, got, ldc)
        g = oracle.symm(dtype, herm, x, uplo, m, n, alpha, a, lda, b, ldb, beta, want, ldc)
        K = (n if x else m) + 2
        touched = np.zeros(c0.shape, dtype=bool)
        touched[:n, :m] = True
        gauge = np.zeros(c0.shape)
        gauge[:n, :m] = g[:n, :m]
    else:
        lib_call("rankk", dtype, herm, x, uplo, trans, n, k, alpha, a, lda, b if x else a, ldb if x else lda, beta, got, ldc)
        g = oracle.rankk(dtype, herm, x, uplo, trans, n, k, alpha, a, lda, b, ldb, beta, want, ldc)
        K = (2 * k if x else k) + 2
        jj, ii = np.meshgrid(np.arange(c0.shape[0]), np.arange(c0.shape[1]), indexing="ij")   # c0[j, i] = C(i, j)
        touched = (ii < n) & (jj < n) & ((ii >= jj) if uplo else (ii <= jj))
        gauge = np.zeros(c0.shape)
        gauge[:n, :n] = g[:n, :n]
    return got, want, gauge, K, touched


def check_case(case, got, want, gauge, K, touched, c0, c_bound=C_BOUND):
    dtype = case[1]
    # everything outside the window / triangle keeps its bits
    assert np.array_equal(got.view(np.uint8)[np.repeat(~touched, got.itemsize, axis=1)],
                          c0.view(np.uint8)[np.repeat(~touched, c0.itemsize, axis=1)]), ("untouched part changed",) + tuple(case[:9])
    err = np.abs(got.astype(np.complex128) - want.astype(np.complex128))[touched]
    bound = c_bound * K * EPS[dtype] * gauge[touched]
    bad = err > bound + 1e-300
    assert not bad.any(), ("componentwise bound", float((err[bad] / np.maximum(bound[bad], 1e-300)).max())) + tuple(case[:9])
    return float((err / np.maximum(bound, 1e-300)).max()) if err.size else 0.0


def meta_case(row):
    kind, dtype, herm, x, uplo, trans, m, n, k, lda, ldb, ldc = (int(v) for v in row[:12])
    cplx = dtype in (cpu.CX, cpu.Z)
    alpha = complex(row[12], row[13]) if cplx else float(row[12])
    beta = complex(row[14], row[15]) if cplx else float(row[14])
    return (kind, dtype, herm, x, uplo, trans, m, n, k, lda, ldb, ldc, alpha, beta)


def bind(lib):
    def call(which, *args):
        return (cpu.call_symm if which == "symm" else cpu.call_rankk)(lib, *args)
    return call


# ---- TRMM / TRSM ---------------------------------------------------------------------------
TRSM_THRESH = 16.0      # ctest's THRESH on err / (eps * gauge) of the multiplied-back solution


def tri_operand(rng, dtype, ka, lda, uplo, unit):
    """Triangular test matrix as ctest's DMAKE builds it (c_dblat3.f:2084-2087: diagonal + 1); the
    triangle that must not be referenced, and a unit diagonal, are NaN."""
    a = operand(rng, dtype, ka, lda)
    a[np.arange(ka), np.arange(ka)] += 1.0
    jj, ii = np.meshgrid(np.arange(ka), np.arange(lda), indexing="ij")
    a[(ii < jj) if uplo else ((ii > jj) & (ii < ka))] = np.nan
    if unit:
        a[np.arange(ka), np.arange(ka)] = np.nan
    return a


def check_trxm(oracle, lib, case, a, b0, ref_b=None):
    """case = (dtype, solve, side, uplo, trans, unit, m, n, lda, ldb, alpha).  Runs `lib` (or takes
    ref_b as its result), checks padding rows, and the product bound (TRMM) or ctest's residual ratio
    (TRSM).  Returns the ratio."""
    dtype, solve, side, uplo, trans, unit, m, n, lda, ldb, alpha = case
    if ref_b is None:
        got = b0.copy()
        cpu.call_trxm(lib, dtype, solve, side, uplo, trans, unit, m, n, alpha, a, lda, got, ldb)
    else:
        got = ref_b
    assert np.array_equal(got[:, m:].view(np.uint8), b0[:, m:].view(np.uint8)), ("padding rows changed",) + tuple(case[:8])
    assert not np.isnan(got[:, :m]).any(), ("NaN leaked from the unreferenced part of A",) + tuple(case[:8])
    if complex(alpha) == 0:
        assert np.all(got[:, :m] == 0), case[:8]
        return 0.0
    if solve:
        ratio = oracle.trsm_residual(dtype, side, uplo, trans, unit, m, n, alpha, a, lda, b0, ldb, got, ldb)
        assert ratio < TRSM_THRESH, ("ctest residual ratio", ratio) + tuple(case[:8])
        return ratio
    want = b0.copy()
    g = oracle.trxm(dtype, 0, side, uplo, trans, unit, m, n, alpha, a, lda, want, ldb)
    K = (n if side else m) + 1
    err = np.abs(got[:, :m].astype(np.complex128) - want[:, :m].astype(np.complex128))
    bound = C_BOUND * K * EPS[dtype] * g[:n, :m]
    assert np.all(err <= bound + 1e-300), ("componentwise bound", float((err / np.maximum(bound, 1e-300)).max())) + tuple(case[:8])
    return float((err / np.maximum(bound, 1e-300)).max())


def trxm_meta_case(row):
    dtype, solve, side, uplo, trans, unit, m, n, lda, ldb = (int(v) for v in row[:10])
    alpha = complex(row[10], row[11]) if dtype in (cpu.CX, cpu.Z) else float(row[10])
    return (dtype, solve, side, uplo, trans, unit, m, n, lda, ldb, alpha)
