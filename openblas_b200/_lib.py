"""ctypes binding of libopenblas_b200.so -- the same symbols a C caller of the reference uses
(include/openblas_b200.h).  Loading fails loudly when the library has not been built: there is
no Python or CPU fallback behind this module."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OPENBLAS_B200_LIB") or os.path.join(_HERE, "lib", "libopenblas_b200.so")   # override: experiments only


class LibraryMissing(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            f"{LIB_PATH} is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C openblas_b200/csrc`). openblas_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    i, i64, vp, f, d = C.c_int, C.c_int64, C.c_void_p, C.c_float, C.c_double
    for name, scal in (("cblas_sgemm", f), ("cblas_dgemm", d), ("cblas_sbgemm", f)):
        fn = getattr(lib, name)
        fn.restype = None
        fn.argtypes = [i, i, i, i, i, i, scal, vp, i, vp, i, scal, vp, i]
    for name in ("cblas_cgemm", "cblas_zgemm", "cblas_cgemm3m", "cblas_zgemm3m"):
        fn = getattr(lib, name)
        fn.restype = None
        fn.argtypes = [i, i, i, i, i, i, vp, vp, i, vp, i, vp, vp, i]
    for name in ("sgemm_", "dgemm_", "cgemm_", "zgemm_", "sbgemm_", "cgemm3m_", "zgemm3m_"):
        fn = getattr(lib, name)
        fn.restype = None
        fn.argtypes = [C.c_char_p, C.c_char_p] + [vp] * 11
    for name in ("cblas_sgemm_batch", "cblas_dgemm_batch", "cblas_cgemm_batch", "cblas_zgemm_batch",
                 "cblas_sbgemm_batch"):
        fn = getattr(lib, name)
        fn.restype = None
        fn.argtypes = [i] + [vp] * 13 + [i, vp]
    for name in ("cblas_sbstobf16", "cblas_sbdtobf16", "cblas_sbf16tos", "cblas_dbf16tod"):
        fn = getattr(lib, name)
        fn.restype = None
        fn.argtypes = [i, vp, i, vp, i]
    lib.b200_gemm.restype = i
    lib.b200_gemm.argtypes = [i, i, i, i64, i64, i64, vp, vp, i64, vp, i64, vp, vp, i64]
    lib.b200_gemm_async.restype = i
    lib.b200_gemm_async.argtypes = lib.b200_gemm.argtypes + [vp]
    lib.b200_set_kernel.argtypes = [i]
    lib.b200_get_kernel.restype = i
    lib.b200_launch_count.restype = C.c_uint64
    lib.b200_last_kernel.restype = C.c_char_p
    lib.b200_last_error.restype = C.c_char_p
    lib.b200_version.restype = C.c_char_p
    lib.b200_init.restype = i
    lib.b200_init.argtypes = [i]
    lib.b200_host_alloc.restype = vp
    lib.b200_host_alloc.argtypes = [C.c_size_t]
    lib.b200_host_free.argtypes = [vp]
    lib.openblas_get_config.restype = C.c_char_p
    lib.openblas_get_corename.restype = C.c_char_p
    lib.openblas_get_num_threads.restype = i
    lib.openblas_get_num_procs.restype = i
    lib.openblas_get_parallel.restype = i
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib
