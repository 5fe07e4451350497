#!/usr/bin/env python3
"""SBGEMMT against SBGEMM on device-resident operands (CUDA events around `reps` back-to-back calls after warm-up):
m x m x k with k = m, NN and TN, both triangles.  One JSON line per point: ms, TFLOP/s counted on the triangle's
m (m + 1) k flops, and the ratio to the full SBGEMM's time (0.5 + 1 / (2 tiles) is the floor of a tile walk)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import openblas_b200 as ob  # noqa: E402
from oracle import cpu  # noqa: E402


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    lib = ob.lib()
    g = torch.Generator(device="cuda").manual_seed(1)
    for m in (2048, 4096, 8192, 16384):
        k = m
        a = (torch.rand(m * k, device="cuda", generator=g) - 0.5).to(torch.bfloat16)
        b = (torch.rand(m * k, device="cuda", generator=g) - 0.5).to(torch.bfloat16)
        c = torch.zeros(m * m, device="cuda", dtype=torch.float32)
        reps = 20 if m <= 8192 else 5
        for ta, tb, name in ((0, 0, "NN"), (1, 0, "TN")):
            lda = k if ta else m
            full = timed(lambda: ob.cblas.sbgemm(102, 112 if ta else 111, 111, m, m, k, 1.0, a, lda, b, k, 0.0, c, m), reps)
            for uplo in (0, 1):
                t = timed(lambda: cpu.call_sbgemmt(lib, uplo, ta, tb, m, k, 1.0, a.data_ptr(), lda, b.data_ptr(), k, 0.0, c.data_ptr(), m), reps)
                print(json.dumps({"routine": "sbgemmt", "m": m, "k": k, "ops": name, "uplo": "UL"[uplo], "ms": t, "kernel": ob.cblas.last_kernel(),
                                  "tflops_triangle": m * (m + 1) * k / t / 1e9, "sbgemm_ms": full, "sbgemm_tflops": 2.0 * m * m * k / full / 1e9,
                                  "time_vs_sbgemm": t / full}), flush=True)
        del a, b, c


if __name__ == "__main__":
    main()
