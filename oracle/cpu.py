"""TEST INFRASTRUCTURE -- ctypes access to the CPU checkers.

  * `Oracle`     oracle/_build/liboracle.so, the plain-C restatement (gemm_oracle.c)
  * `Reference`  oracle/_ref/<target>/libopenblas_ref.so, the unmodified reference compiled from
                 its own sources (build_ref.py); the target is chosen from the host CPU's flags

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
Nothing under openblas_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

S, D, CX, Z, SB = 0, 1, 2, 3, 4
DTYPE_NAMES = {S: "s", D: "d", CX: "c", Z: "z", SB: "sb"}
NP_IN = {S: np.float32, D: np.float64, CX: np.complex64, Z: np.complex128, SB: np.uint16}
NP_OUT = {S: np.float32, D: np.float64, CX: np.complex64, Z: np.complex128, SB: np.float32}
TRANS_CHAR = "NTRC"
CBLAS_TRANS = {0: 111, 1: 112, 2: 114, 3: 113}


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def scalar_bytes(dtype, value):
    """alpha/beta as the in-memory object the C ABI points at."""
    if dtype in (CX, Z):
        t = np.float32 if dtype == CX else np.float64
        v = complex(value)
        return np.array([v.real, v.imag], dtype=t)
    t = np.float64 if dtype == D else np.float32
    return np.array([value], dtype=t)


def build_oracle():
    so = os.path.join(HERE, "_build", "liboracle.so")
    srcs = [os.path.join(HERE, f) for f in ("gemm_oracle.c", "level3_oracle.c")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", HERE, "_build/liboracle.so"], stdout=subprocess.DEVNULL)
    return so


class Oracle:
    def __init__(self):
        self.lib = C.CDLL(build_oracle())
        L = self.lib
        L.oracle_gemm.restype = C.c_int
        L.oracle_gemm.argtypes = [C.c_int, C.c_int, C.c_int, C.c_long, C.c_long, C.c_long, C.c_void_p,
                                  C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p,
                                  C.c_long, C.c_long, C.c_long]
        L.oracle_gemm_small.restype = C.c_int
        L.oracle_gemm_small.argtypes = L.oracle_gemm.argtypes[:14]
        L.oracle_mmch.restype = C.c_double
        L.oracle_mmch.argtypes = [C.c_int, C.c_int, C.c_int, C.c_long, C.c_long, C.c_long, C.c_void_p,
                                  C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p,
                                  C.c_long, C.c_void_p, C.c_long]
        L.oracle_componentwise_ratio.restype = C.c_double
        L.oracle_componentwise_ratio.argtypes = L.oracle_mmch.argtypes + [C.c_void_p, C.c_long]
        L.oracle_check_args.restype = C.c_int
        L.oracle_check_args.argtypes = [C.c_int, C.c_int] + [C.c_long] * 6 + [C.c_int]
        L.oracle_f32_to_bf16.restype = C.c_uint16
        L.oracle_f32_to_bf16.argtypes = [C.c_float]
        L.oracle_bf16_to_f32.restype = C.c_float
        L.oracle_bf16_to_f32.argtypes = [C.c_uint16]
        L.oracle_tobf16.argtypes = [C.c_long, C.c_void_p, C.c_long, C.c_void_p, C.c_long]
        L.oracle_symm.restype = C.c_int
        L.oracle_symm.argtypes = [C.c_int] * 4 + [C.c_long, C.c_long, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_long,
                                                  C.c_void_p, C.c_void_p, C.c_long, C.c_void_p]
        L.oracle_rankk.restype = C.c_int
        L.oracle_rankk.argtypes = [C.c_int] * 5 + [C.c_long, C.c_long, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_long,
                                                   C.c_void_p, C.c_void_p, C.c_long, C.c_void_p]
        L.oracle_trxm.restype = C.c_int
        L.oracle_trxm.argtypes = [C.c_int] * 6 + [C.c_long, C.c_long, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_void_p]
        L.oracle_trsm_residual.restype = C.c_double
        L.oracle_trsm_residual.argtypes = [C.c_int] * 5 + [C.c_long, C.c_long, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_long,
                                                            C.c_void_p, C.c_long]
        L.oracle_check_trxm.restype = C.c_int
        L.oracle_check_trxm.argtypes = [C.c_int] * 4 + [C.c_long] * 4 + [C.c_int]
        L.oracle_check_symm.restype = C.c_int
        L.oracle_check_symm.argtypes = [C.c_int, C.c_int] + [C.c_long] * 5 + [C.c_int]
        L.oracle_check_rankk.restype = C.c_int
        L.oracle_check_rankk.argtypes = [C.c_int, C.c_int, C.c_int] + [C.c_long] * 5 + [C.c_int]
        L.oracle_bf16to.argtypes = [C.c_long, C.c_void_p, C.c_long, C.c_void_p, C.c_long]
        L.oracle_gemmt.restype = C.c_int
        L.oracle_gemmt.argtypes = [C.c_int] * 4 + [C.c_long, C.c_long, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_long,
                                                   C.c_void_p, C.c_void_p, C.c_long, C.c_void_p]
        L.oracle_check_gemmt.restype = C.c_int
        L.oracle_check_gemmt.argtypes = [C.c_int] * 4 + [C.c_long] * 5 + [C.c_int]
        L.oracle_sbgemv.restype = C.c_int
        L.oracle_sbgemv.argtypes = [C.c_int, C.c_long, C.c_long, C.c_float, C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_float,
                                    C.c_void_p, C.c_long, C.c_void_p]
        L.oracle_sbdot.restype = C.c_double
        L.oracle_sbdot.argtypes = [C.c_long, C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_void_p]
        L.oracle_sbgemmt.restype = C.c_int
        L.oracle_sbgemmt.argtypes = [C.c_int] * 3 + [C.c_long, C.c_long, C.c_float, C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_float,
                                                     C.c_void_p, C.c_long, C.c_void_p]
        L.oracle_check_sbgemmt.restype = C.c_int
        L.oracle_check_sbgemmt.argtypes = [C.c_int] * 4 + [C.c_long] * 5 + [C.c_int]
        L.oracle_check_sbgemv.restype = C.c_int
        L.oracle_check_sbgemv.argtypes = [C.c_int] + [C.c_long] * 5 + [C.c_int]

    def gemm(self, dtype, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, q=0, unroll_m=0, small=False):
        """In-place on c (column-major storage, numpy arrays of the dtype's element type)."""
        al, be = scalar_bytes(dtype, alpha), scalar_bytes(dtype, beta)
        if small:
            r = self.lib.oracle_gemm_small(dtype, ta, tb, m, n, k, _ptr(al), _ptr(a), lda, _ptr(b), ldb,
                                           _ptr(be), _ptr(c), ldc)
        else:
            r = self.lib.oracle_gemm(dtype, ta, tb, m, n, k, _ptr(al), _ptr(a), lda, _ptr(b), ldb,
                                     _ptr(be), _ptr(c), ldc, q, unroll_m)
        assert r == 0
        return c

    def mmch(self, dtype, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c0, ldc0, cc, ldcc):
        al, be = scalar_bytes(dtype, alpha), scalar_bytes(dtype, beta)
        return self.lib.oracle_mmch(dtype, ta, tb, m, n, k, _ptr(al), _ptr(a), lda, _ptr(b), ldb, _ptr(be),
                                    _ptr(c0), ldc0, _ptr(cc), ldcc)

    def ratio(self, dtype, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c0, ldc0, x, ldx, y, ldy):
        al, be = scalar_bytes(dtype, alpha), scalar_bytes(dtype, beta)
        return self.lib.oracle_componentwise_ratio(dtype, ta, tb, m, n, k, _ptr(al), _ptr(a), lda, _ptr(b),
                                                   ldb, _ptr(be), _ptr(c0), ldc0, _ptr(x), ldx, _ptr(y), ldy)

    def check_args(self, ta, tb, m, n, k, lda, ldb, ldc, ok):
        return self.lib.oracle_check_args(ta, tb, m, n, k, lda, ldb, ldc, ok)

    def symm(self, dtype, herm, side, uplo, m, n, alpha, a, lda, b, ldb, beta, c, ldc):
        """SYMM (herm=0) / HEMM (herm=1) in place on c; returns the m x n gauge (column-major, as (n, m) array)."""
        al, be = scalar_bytes(dtype, alpha), scalar_bytes(dtype, beta)
        g = np.zeros((max(n, 1), max(m, 1)), dtype=np.float64)
        assert self.lib.oracle_symm(dtype, herm, side, uplo, m, n, _ptr(al), _ptr(a), lda, _ptr(b), ldb, _ptr(be), _ptr(c),
                                    ldc, _ptr(g)) == 0
        return g

    def rankk(self, dtype, herm, two, uplo, trans, n, k, alpha, a, lda, b, ldb, beta, c, ldc):
        """SYRK/HERK (two=0) or SYR2K/HER2K (two=1) in place on c; returns the n x n gauge (0 outside the triangle)."""
        al, be = scalar_bytes(dtype, alpha), scalar_bytes(dtype, beta)
        g = np.zeros((max(n, 1), max(n, 1)), dtype=np.float64)
        assert self.lib.oracle_rankk(dtype, herm, two, uplo, trans, n, k, _ptr(al), _ptr(a), lda, _ptr(b if two else a),
                                     ldb if two else lda, _ptr(be), _ptr(c), ldc, _ptr(g)) == 0
        return g

    def trxm(self, dtype, solve, side, uplo, trans, unit, m, n, alpha, a, lda, b, ldb):
        """TRMM (solve=0) / TRSM (solve=1) in place on b; returns the m x n gauge of the product form."""
        al = scalar_bytes(dtype, alpha)
        g = np.zeros((max(n, 1), max(m, 1)), dtype=np.float64)
        assert self.lib.oracle_trxm(dtype, solve, side, uplo, trans, unit, m, n, _ptr(al), _ptr(a), lda, _ptr(b), ldb, _ptr(g)) == 0
        return g

    def trsm_residual(self, dtype, side, uplo, trans, unit, m, n, alpha, a, lda, b0, ldb0, x, ldx):
        """ctest's acceptance ratio of a TRSM solution x (c_dblat3.f:1195-1235): passes below 16."""
        al = scalar_bytes(dtype, alpha)
        return self.lib.oracle_trsm_residual(dtype, side, uplo, trans, unit, m, n, _ptr(al), _ptr(a), lda, _ptr(b0), ldb0, _ptr(x), ldx)

    def check_trxm(self, side, uplo, trans, unit, m, n, lda, ldb, ok):
        return self.lib.oracle_check_trxm(side, uplo, trans, unit, m, n, lda, ldb, ok)

    def check_symm(self, side, uplo, m, n, lda, ldb, ldc, ok):
        return self.lib.oracle_check_symm(side, uplo, m, n, lda, ldb, ldc, ok)

    def check_rankk(self, two, uplo, trans, n, k, lda, ldb, ldc, ok):
        return self.lib.oracle_check_rankk(two, uplo, trans, n, k, lda, ldb, ldc, ok)

    def gemmt(self, dtype, uplo, ta, tb, m, k, alpha, a, lda, b, ldb, beta, c, ldc):
        """?GEMMT in place on the uplo triangle of c (column-major); returns the m x m gauge (0 outside the triangle)."""
        al, be = scalar_bytes(dtype, alpha), scalar_bytes(dtype, beta)
        g = np.zeros((max(m, 1), max(m, 1)), dtype=np.float64)
        assert self.lib.oracle_gemmt(dtype, uplo, ta, tb, m, k, _ptr(al), _ptr(a), lda, _ptr(b), ldb, _ptr(be), _ptr(c), ldc, _ptr(g)) == 0
        return g

    def check_gemmt(self, rowmajor, uplo, ta, tb, m, k, lda, ldb, ldc, ok):
        return self.lib.oracle_check_gemmt(rowmajor, uplo, ta, tb, m, k, lda, ldb, ldc, ok)

    def sbgemv(self, trans, m, n, alpha, a, lda, x, incx, beta, y, incy):
        """SBGEMV in place on y; x and y are the arrays as the CALLER holds them (lowest address first; a negative
        increment walks from the far end, as in the interface).  Returns the gauge per output."""
        lenx, leny = (m, n) if trans else (n, m)
        g = np.zeros(max(leny, 1), dtype=np.float64)
        assert self.lib.oracle_sbgemv(trans, m, n, alpha, _ptr(a), lda, _ptr(x), incx, beta, _ptr(y), incy, _ptr(g)) == 0
        return g

    def sbdot(self, n, x, incx, y, incy):
        g = C.c_double(0.0)
        return self.lib.oracle_sbdot(n, _ptr(x), incx, _ptr(y), incy, C.byref(g)), g.value

    def sbgemmt(self, uplo, ta, tb, m, k, alpha, a, lda, b, ldb, beta, c, ldc):
        """SBGEMMT in place on the uplo triangle of the fp32 c (a, b: uint16 bf16 arrays); returns the m x m gauge."""
        g = np.zeros((max(m, 1), max(m, 1)), dtype=np.float64)
        assert self.lib.oracle_sbgemmt(uplo, ta, tb, m, k, alpha, _ptr(a), lda, _ptr(b), ldb, beta, _ptr(c), ldc, _ptr(g)) == 0
        return g

    def check_sbgemmt(self, rowmajor, uplo, ta, tb, m, k, lda, ldb, ldc, ok):
        return self.lib.oracle_check_sbgemmt(rowmajor, uplo, ta, tb, m, k, lda, ldb, ldc, ok)

    def check_sbgemv(self, trans, m, n, lda, incx, incy, ok):
        return self.lib.oracle_check_sbgemv(trans, m, n, lda, incx, incy, ok)

    def tobf16(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty(x.shape, dtype=np.uint16)
        self.lib.oracle_tobf16(x.size, _ptr(x), 1, _ptr(out), 1)
        return out

    def bf16to(self, x):
        x = np.ascontiguousarray(x, dtype=np.uint16)
        out = np.empty(x.shape, dtype=np.float32)
        self.lib.oracle_bf16to(x.size, _ptr(x), 1, _ptr(out), 1)
        return out


def cpu_flags():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("flags"):
                return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def best_target(flags=None):
    """Most capable reference build this host can execute (oracle/ref_recipe/*.recipe)."""
    f = cpu_flags() if flags is None else flags
    if {"amx_bf16", "amx_tile", "avx512_bf16", "avx512f", "avx512vl"} <= f:
        return "sapphirerapids"
    if {"avx512f", "avx512vl", "avx512dq", "avx512bw", "avx512cd"} <= f:
        return "skylakex"
    if {"avx2", "fma"} <= f:
        return "haswell"
    return "generic"


def ref_path(target=None):
    return os.path.join(HERE, "_ref", target or best_target(), "libopenblas_ref.so")


def have_reference(target=None):
    return os.path.exists(ref_path(target))


def _real_scalar(dtype, value):
    return np.array([float(np.real(value))], dtype=np.float32 if dtype in (S, CX) else np.float64)


def call_symm(lib, dtype, herm, side, uplo, m, n, alpha, a, lda, b, ldb, beta, c, ldc):
    """?symm_ / ?hemm_ of ANY library exporting the Fortran ABI (common_interface.h:554-605): the
    reference (oracle/_ref) in the pin tests, the library under test in the GPU tests."""
    fn = getattr(lib, DTYPE_NAMES[dtype] + ("hemm_" if herm else "symm_"))
    al, be = scalar_bytes(dtype, alpha), scalar_bytes(dtype, beta)
    i = lambda v: C.byref(C.c_int(int(v)))
    fn(C.c_char_p(b"LR"[side:side + 1]), C.c_char_p(b"UL"[uplo:uplo + 1]), i(m), i(n), _ptr(al), _ptr(a), i(lda), _ptr(b),
       i(ldb), _ptr(be), _ptr(c), i(ldc))
    return c


def call_rankk(lib, dtype, herm, two, uplo, trans, n, k, alpha, a, lda, b, ldb, beta, c, ldc):
    """?syrk_ / ?herk_ / ?syr2k_ / ?her2k_ (common_interface.h:574-626).  HERK takes real alpha and
    beta, HER2K complex alpha and real beta."""
    name = DTYPE_NAMES[dtype] + {(0, 0): "syrk_", (1, 0): "herk_", (0, 1): "syr2k_", (1, 1): "her2k_"}[(herm, two)]
    fn = getattr(lib, name)
    al = _real_scalar(dtype, alpha) if (herm and not two) else scalar_bytes(dtype, alpha)
    be = _real_scalar(dtype, beta) if herm else scalar_bytes(dtype, beta)
    i = lambda v: C.byref(C.c_int(int(v)))
    tch = (b"NC" if herm else b"NT")[trans:trans + 1]
    args = [C.c_char_p(b"UL"[uplo:uplo + 1]), C.c_char_p(tch), i(n), i(k), _ptr(al), _ptr(a), i(lda)]
    if two:
        args += [_ptr(b), i(ldb)]
    fn(*(args + [_ptr(be), _ptr(c), i(ldc)]))
    return c


def call_trxm(lib, dtype, solve, side, uplo, trans, unit, m, n, alpha, a, lda, b, ldb):
    """?trmm_ / ?trsm_ (common_interface.h:530-552) of any library; trans 0..3 = N, T, R, C."""
    fn = getattr(lib, DTYPE_NAMES[dtype] + ("trsm_" if solve else "trmm_"))
    al = scalar_bytes(dtype, alpha)
    i = lambda v: C.byref(C.c_int(int(v)))
    fn(C.c_char_p(b"LR"[side:side + 1]), C.c_char_p(b"UL"[uplo:uplo + 1]), C.c_char_p(b"NTRC"[trans:trans + 1]),
       C.c_char_p(b"NU"[unit:unit + 1]), i(m), i(n), _ptr(al), _ptr(a), i(lda), _ptr(b), i(ldb))
    return b


def call_gemmt(lib, dtype, uplo, ta, tb, m, k, alpha, a, lda, b, ldb, beta, c, ldc, cblas=False, rowmajor=False):
    """?gemmt_ (common_interface.h:506-513) or cblas_?gemmt (cblas.h:311-318) of any library."""
    al, be = scalar_bytes(dtype, alpha), scalar_bytes(dtype, beta)
    P = lambda v: C.c_void_p(v) if isinstance(v, int) else _ptr(v)        # device addresses come as ints
    if cblas:
        fn = getattr(lib, "cblas_" + DTYPE_NAMES[dtype] + "gemmt")
        if dtype in (CX, Z):
            sa, sb = _ptr(al), _ptr(be)
        else:
            t = C.c_double if dtype == D else C.c_float
            sa, sb = t(float(np.real(alpha))), t(float(np.real(beta)))
        fn(101 if rowmajor else 102, 121 if uplo == 0 else 122, CBLAS_TRANS[ta], CBLAS_TRANS[tb], C.c_int(m), C.c_int(k), sa,
           P(a), C.c_int(lda), P(b), C.c_int(ldb), sb, P(c), C.c_int(ldc))
        return c
    fn = getattr(lib, DTYPE_NAMES[dtype] + "gemmt_")
    i = lambda v: C.byref(C.c_int(int(v)))
    fn(C.c_char_p(b"UL"[uplo:uplo + 1]), C.c_char_p(TRANS_CHAR[ta].encode()), C.c_char_p(TRANS_CHAR[tb].encode()), i(m), i(k),
       _ptr(al), P(a), i(lda), P(b), i(ldb), _ptr(be), P(c), i(ldc))
    return c


def call_sbgemmt(lib, uplo, ta, tb, m, k, alpha, a, lda, b, ldb, beta, c, ldc, cblas=False, rowmajor=False):
    """sbgemmt_ or cblas_sbgemmt (interface/sbgemmt.c:47-52,143-149; neither is declared in cblas.h) of any library."""
    P = lambda v: C.c_void_p(v) if isinstance(v, int) else _ptr(v)
    if cblas:
        lib.cblas_sbgemmt(101 if rowmajor else 102, 121 if uplo == 0 else 122, CBLAS_TRANS[ta], CBLAS_TRANS[tb], C.c_int(m), C.c_int(k),
                          C.c_float(alpha), P(a), C.c_int(lda), P(b), C.c_int(ldb), C.c_float(beta), P(c), C.c_int(ldc))
        return c
    i = lambda v: C.byref(C.c_int(int(v)))
    lib.sbgemmt_(C.c_char_p(b"UL"[uplo:uplo + 1]), C.c_char_p(TRANS_CHAR[ta].encode()), C.c_char_p(TRANS_CHAR[tb].encode()), i(m), i(k),
                 C.byref(C.c_float(alpha)), P(a), i(lda), P(b), i(ldb), C.byref(C.c_float(beta)), P(c), i(ldc))
    return c


def call_sbgemv(lib, trans, m, n, alpha, a, lda, x, incx, beta, y, incy, cblas=False, rowmajor=False):
    """sbgemv_ (common_interface.h:258) or cblas_sbgemv (cblas.h:442) of any library; a, x, y: addresses or arrays."""
    P = lambda v: C.c_void_p(v) if isinstance(v, int) else _ptr(v)
    if cblas:
        lib.cblas_sbgemv(101 if rowmajor else 102, CBLAS_TRANS[trans], C.c_int(m), C.c_int(n), C.c_float(alpha), P(a), C.c_int(lda),
                         P(x), C.c_int(incx), C.c_float(beta), P(y), C.c_int(incy))
        return y
    i = lambda v: C.byref(C.c_int(int(v)))
    lib.sbgemv_(C.c_char_p(TRANS_CHAR[trans].encode()), i(m), i(n), C.byref(C.c_float(alpha)), P(a), i(lda), P(x), i(incx),
                C.byref(C.c_float(beta)), P(y), i(incy))
    return y


def call_sbdot(lib, n, x, incx, y, incy, cblas=False):
    """sbdot_ (common_interface.h:62) or cblas_sbdot (cblas.h:441) of any library."""
    P = lambda v: C.c_void_p(v) if isinstance(v, int) else _ptr(v)
    if cblas:
        lib.cblas_sbdot.restype = C.c_float
        return float(lib.cblas_sbdot(C.c_int(n), P(x), C.c_int(incx), P(y), C.c_int(incy)))
    lib.sbdot_.restype = C.c_float
    i = lambda v: C.byref(C.c_int(int(v)))
    return float(lib.sbdot_(i(n), P(x), i(incx), P(y), i(incy)))


class Reference:
    """The reference library's own C ABI (cblas.h:298-307,444-445; common_interface.h:484-497)."""

    def __init__(self, target=None):
        self.target = target or best_target()
        # DEEPBIND: the reference must bind its internal calls (goto_set_num_threads, xerbla_, ...) to
        # itself even when a library exporting the same names is already loaded in the process
        self.lib = C.CDLL(ref_path(self.target), mode=os.RTLD_LOCAL | os.RTLD_DEEPBIND)
        self.lib.openblas_get_config.restype = C.c_char_p
        self.lib.openblas_get_corename.restype = C.c_char_p
        self.lib.openblas_get_num_threads.restype = C.c_int

    def config(self):
        return self.lib.openblas_get_config().decode()

    def set_threads(self, n):
        self.lib.openblas_set_num_threads(int(n))

    def threads(self):
        return self.lib.openblas_get_num_threads()

    def gemm(self, dtype, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc):
        """Column-major, through the Fortran symbol (what benchmark/gemm.c calls)."""
        fn = getattr(self.lib, DTYPE_NAMES[dtype] + "gemm_")
        al, be = scalar_bytes(dtype, alpha), scalar_bytes(dtype, beta)
        i = lambda v: C.byref(C.c_int(int(v)))
        fn(C.c_char_p(TRANS_CHAR[ta].encode()), C.c_char_p(TRANS_CHAR[tb].encode()), i(m), i(n), i(k),
           _ptr(al), _ptr(a), i(lda), _ptr(b), i(ldb), _ptr(be), _ptr(c), i(ldc))
        return c
