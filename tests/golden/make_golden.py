#!/usr/bin/env python3
"""Writes tests/golden/gemm_golden.npz and bf16_golden.npz: fixtures produced by the
REFERENCE itself (run in the authoring container, where /root/reference is mounted and
oracle/_ref has been built from it by oracle/build_ref.py).

  gemm_golden.npz   seeded inputs and the outputs of the reference's GENERIC-target build
                    (deterministic, single thread) for every precision, every op combination,
                    ragged shapes, padded leading dimensions and the ctest alpha/beta values.
                    The oracle must reproduce these bit for bit; the GPU path within the bound.
  bf16_golden.npz   fp32 -> bf16 -> fp32 conversions by the reference's sbstobf16_/sbf16tos_.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cpu  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def make_operand(rng, dtype, cols, ld, oracle):
    if dtype == cpu.SB:
        return oracle.tobf16(rng.random((cols, ld), dtype=np.float32) - 0.5)
    if dtype in (cpu.CX, cpu.Z):
        x = (rng.random((cols, ld)) - 0.5) + 1j * (rng.random((cols, ld)) - 0.5)
    else:
        x = rng.random((cols, ld)) - 0.5
    return x.astype(cpu.NP_IN[dtype])


def cases():
    shapes = [(1, 1, 1), (2, 3, 5), (7, 5, 3), (9, 35, 2), (35, 9, 7), (17, 13, 70), (5, 9, 150), (31, 33, 37)]
    i = 0
    for dtype in (cpu.S, cpu.D, cpu.CX, cpu.Z, cpu.SB):
        ntr = 4 if dtype in (cpu.CX, cpu.Z) else 2
        ab = [(0.0, 0.0), (1.0, 0.0), (0.7, 1.3), (0.0, 1.3), (1.0, 1.0)]
        if dtype in (cpu.CX, cpu.Z):
            ab = [(0.0, 0.0), (1.0, 0.0), (0.7 - 0.9j, 1.3 - 1.1j), (0.0, 1.3 - 1.1j)]
        for ta in range(ntr):
            for tb in range(ntr):
                for rep in range(2):
                    m, n, k = shapes[i % len(shapes)]
                    alpha, beta = ab[(i // 2) % len(ab)] if rep else ab[2]
                    i += 1
                    yield dtype, ta, tb, m, n, k, alpha, beta


def main():
    oracle = cpu.Oracle()
    ref = cpu.Reference("generic")
    ref.set_threads(1)
    rng = np.random.default_rng(20261017)
    out = {}
    meta = []
    for idx, (dtype, ta, tb, m, n, k, alpha, beta) in enumerate(cases()):
        ra, ca = (k, m) if ta & 1 else (m, k)
        rb, cb = (n, k) if tb & 1 else (k, n)
        lda, ldb, ldc = ra + 1, rb + 1, m + 1
        a = make_operand(rng, dtype, ca, lda, oracle)
        b = make_operand(rng, dtype, cb, ldb, oracle)
        c0 = make_operand(rng, cpu.S if dtype == cpu.SB else dtype, n, ldc, oracle).astype(cpu.NP_OUT[dtype])
        c0[:, m:] = -1e10  # rogue padding rows, as ctest does (c_dblat3.f:2106)
        c = c0.copy()
        ref.gemm(dtype, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc)
        out[f"a{idx}"], out[f"b{idx}"], out[f"c0_{idx}"], out[f"c{idx}"] = a, b, c0, c
        meta.append([dtype, ta, tb, m, n, k, lda, ldb, ldc, complex(alpha).real, complex(alpha).imag,
                     complex(beta).real, complex(beta).imag])
    out["meta"] = np.array(meta, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "gemm_golden.npz"), **out)
    print("gemm_golden.npz:", len(meta), "cases")

    # bf16 conversion vectors through the reference's Fortran-ABI helpers
    lib = ref.lib
    special = np.array([0.0, -0.0, 1.0, -1.0, 1.00390625, 1.01171875, 3.140625, 1e-40, -1e-40, 65504.0,
                        3.3895314e38, np.inf, -np.inf, np.nan, 0.1, 0.2, 0.3, 1.0 / 3.0], dtype=np.float32)
    x = np.concatenate([special, (rng.random(4096, dtype=np.float32) - 0.5) * 8,
                        rng.standard_normal(1024).astype(np.float32) * 1e-3])
    n = C.c_int(x.size)
    one = C.c_int(1)
    h = np.zeros(x.size, dtype=np.uint16)
    lib.sbstobf16_(C.byref(n), x.ctypes.data_as(C.c_void_p), C.byref(one), h.ctypes.data_as(C.c_void_p), C.byref(one))
    back = np.zeros(x.size, dtype=np.float32)
    lib.sbf16tos_(C.byref(n), h.ctypes.data_as(C.c_void_p), C.byref(one), back.ctypes.data_as(C.c_void_p), C.byref(one))
    np.savez_compressed(os.path.join(OUT, "bf16_golden.npz"), x=x, bf16=h, back=back)
    print("bf16_golden.npz:", x.size, "values")


if __name__ == "__main__":
    main()
