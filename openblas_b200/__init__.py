"""openblas_b200 -- B200-native GEMM behind the OpenBLAS C ABI.

The product is the shared library `openblas_b200/lib/libopenblas_b200.so` (sources under
`openblas_b200/csrc/`, public header `include/openblas_b200.h`).  This package is the host-side
mirror of the reference's CBLAS interface for Python callers and tests:

    from openblas_b200 import cblas
    cblas.dgemm(cblas.ColMajor, cblas.NoTrans, cblas.NoTrans, m, n, k, 1.0, A, lda, B, ldb, 0.0, C, ldc)

Array arguments may be numpy arrays (host memory), torch tensors (host, pinned or CUDA) or raw
integer addresses; the library itself decides per pointer whether to stage it.
"""
from . import cblas  # noqa: F401
from ._lib import LIB_PATH, LibraryMissing, lib  # noqa: F401

__version__ = "0.1.0"
