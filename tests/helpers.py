"""Shared test helpers: seeded operands in the reference's storage conventions and the
acceptance thresholds."""
import numpy as np

from oracle import cpu

S, D, CX, Z, SB = cpu.S, cpu.D, cpu.CX, cpu.Z, cpu.SB
ALL_DTYPES = (S, D, CX, Z, SB)
NAMES = cpu.DTYPE_NAMES

# ctest's pass threshold on err/(eps*gauge) (ctest/din3:7)
THRESH = 16.0
# north-star bound |C - C_ref| <= c * k * eps * (|alpha||A||B| + |beta||C|), c stated here
C_BOUND = 2.0
# test/compare_sgemm_sbgemm.c:185-188
SBGEMM_ABS_TOL = 1.0


def ntrans(dtype):
    return 4 if dtype in (CX, Z) else 2


def stored_dims(trans, rows_op, cols_op):
    """(rows, cols) of the stored matrix whose op() is rows_op x cols_op."""
    return (cols_op, rows_op) if trans & 1 else (rows_op, cols_op)


def operand(rng, oracle, dtype, cols, ld, out=False):
    """Column-major storage as a (cols, ld) C-order numpy array, values in (-0.5, 0.5) like
    benchmark/gemm.c:132-140; bf16 inputs are fp32 values rounded by the reference's rule."""
    if dtype == SB and not out:
        return oracle.tobf16(rng.random((cols, ld), dtype=np.float32) - 0.5)
    t = cpu.NP_OUT[dtype] if out else cpu.NP_IN[dtype]
    if dtype in (CX, Z):
        return ((rng.random((cols, ld)) - 0.5) + 1j * (rng.random((cols, ld)) - 0.5)).astype(t)
    return (rng.random((cols, ld)) - 0.5).astype(t)


def problem(rng, oracle, dtype, ta, tb, m, n, k, pad=(1, 2, 3)):
    ra, ca = stored_dims(ta, m, k)
    rb, cb = stored_dims(tb, k, n)
    lda, ldb, ldc = max(1, ra) + pad[0], max(1, rb) + pad[1], max(1, m) + pad[2]
    a = operand(rng, oracle, dtype, max(1, ca), lda)
    b = operand(rng, oracle, dtype, max(1, cb), ldb)
    c = operand(rng, oracle, dtype, max(1, n), ldc, out=True)
    c[:, m:] = -1e10
    return a, lda, b, ldb, c, ldc


def alpha_beta(dtype):
    """The ctest scalar grids (ctest/din3:10-13, zin3:10-13)."""
    if dtype in (CX, Z):
        return [0.0, 1.0, 0.7 - 0.9j], [0.0, 1.0, 1.3 - 1.1j]
    return [0.0, 1.0, 0.7], [0.0, 1.0, 1.3]
