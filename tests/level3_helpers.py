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
