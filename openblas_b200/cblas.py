"""Python mirror of the reference's CBLAS GEMM interface (cblas.h:62-63, 298-309, 444-447), one
function per C symbol, same argument order and meaning, same error behaviour (illegal arguments
are reported through xerbla_ and the call returns without touching C -- interface/gemm.c:473-476).
Every function forwards to the C symbol of the same name in libopenblas_b200.so."""
import ctypes as C

import numpy as np

from ._lib import lib

RowMajor, ColMajor = 101, 102
NoTrans, Trans, ConjTrans, ConjNoTrans = 111, 112, 113, 114

# precision / op codes of the extension entry points (include/openblas_b200.h)
S, D, CX, Z, SB = 0, 1, 2, 3, 4
K_AUTO, K_GENERIC, K_FAST = 0, 1, 2


def addr(x):
    """Address of the first element of a numpy array / torch tensor / int / None."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    raise TypeError(f"cannot take the address of {type(x)}")


def _cscalar(v, ctype):
    arr = (ctype * 2)(v.real, v.imag) if isinstance(v, complex) else (ctype * 2)(float(v), 0.0)
    return arr


def sgemm(order, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc):
    lib().cblas_sgemm(order, ta, tb, m, n, k, alpha, addr(a), lda, addr(b), ldb, beta, addr(c), ldc)


def dgemm(order, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc):
    lib().cblas_dgemm(order, ta, tb, m, n, k, alpha, addr(a), lda, addr(b), ldb, beta, addr(c), ldc)


def sbgemm(order, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc):
    lib().cblas_sbgemm(order, ta, tb, m, n, k, alpha, addr(a), lda, addr(b), ldb, beta, addr(c), ldc)


def cgemm(order, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, _sym="cblas_cgemm"):
    al, be = _cscalar(complex(alpha), C.c_float), _cscalar(complex(beta), C.c_float)
    getattr(lib(), _sym)(order, ta, tb, m, n, k, C.addressof(al), addr(a), lda, addr(b), ldb,
                         C.addressof(be), addr(c), ldc)


def zgemm(order, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, _sym="cblas_zgemm"):
    al, be = _cscalar(complex(alpha), C.c_double), _cscalar(complex(beta), C.c_double)
    getattr(lib(), _sym)(order, ta, tb, m, n, k, C.addressof(al), addr(a), lda, addr(b), ldb,
                         C.addressof(be), addr(c), ldc)


def cgemm3m(*args):
    cgemm(*args, _sym="cblas_cgemm3m")


def zgemm3m(*args):
    zgemm(*args, _sym="cblas_zgemm3m")


GEMM = {S: sgemm, D: dgemm, CX: cgemm, Z: zgemm, SB: sbgemm}
_FORTRAN = {S: "sgemm_", D: "dgemm_", CX: "cgemm_", Z: "zgemm_", SB: "sbgemm_"}
_SCALAR_NP = {S: np.float32, D: np.float64, CX: np.float32, Z: np.float64, SB: np.float32}


def scalar_array(dtype, v):
    if dtype in (CX, Z):
        v = complex(v)
        return np.array([v.real, v.imag], dtype=_SCALAR_NP[dtype])
    return np.array([v], dtype=_SCALAR_NP[dtype])


def fortran_gemm(dtype, transa, transb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc):
    """The Fortran-ABI symbol (?gemm_): every argument by reference, trans as a character."""
    al, be = scalar_array(dtype, alpha), scalar_array(dtype, beta)
    ints = [C.c_int(int(v)) for v in (m, n, k, lda, ldb, ldc)]
    p = lambda x: C.cast(C.byref(x), C.c_void_p)
    getattr(lib(), _FORTRAN[dtype])(transa.encode(), transb.encode(), p(ints[0]), p(ints[1]), p(ints[2]),
                                    al.ctypes.data, addr(a), p(ints[3]), addr(b), p(ints[4]),
                                    be.ctypes.data, addr(c), p(ints[5]))


def gemm_device(dtype, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, stream=None):
    """Extension: column-major GEMM on device pointers, enqueued on `stream` (int handle or
    None) without synchronising; ta/tb are op codes 0..3 (N, T, conj, conj-trans)."""
    al, be = scalar_array(dtype, alpha), scalar_array(dtype, beta)
    err = lib().b200_gemm_async(dtype, ta, tb, m, n, k, al.ctypes.data, addr(a), lda, addr(b), ldb,
                                be.ctypes.data, addr(c), ldc, stream)
    if err:
        raise RuntimeError(f"b200_gemm_async failed: {lib().b200_last_error().decode()} [{err}]")


def gemm_any(dtype, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc):
    """Extension: synchronous column-major GEMM, op codes 0..3, host or device pointers."""
    al, be = scalar_array(dtype, alpha), scalar_array(dtype, beta)
    err = lib().b200_gemm(dtype, ta, tb, m, n, k, al.ctypes.data, addr(a), lda, addr(b), ldb,
                          be.ctypes.data, addr(c), ldc)
    if err:
        raise RuntimeError(f"b200_gemm failed: {lib().b200_last_error().decode()} [{err}]")


def set_kernel(kind):
    lib().b200_set_kernel(kind)


def last_kernel():
    return lib().b200_last_kernel().decode()


def launch_count():
    return int(lib().b200_launch_count())
