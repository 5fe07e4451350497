"""GPU parity tests of the round-2 kernels and runtime paths, through the C ABI, against the oracle:
the warp-specialised TMA SGEMM / CGEMM (sgemm_ws.cuh), the grouped gemm_batch kernel, the 1-D grid of
the generic kernel, library shutdown / re-initialisation."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import cpu
from helpers import alpha_beta, ntrans, problem
from test_gemm_gpu import CB, check

pytestmark = pytest.mark.gpu


def _pad4(dtype, ta, tb, m, n, k):
    """paddings that make lda, ldb multiples of 4 elements (TMA: 16-byte pitches) but not of the tile"""
    ra, rb = (k if ta & 1 else m), (n if tb & 1 else k)
    return (4 + (-ra) % 4, 8 + (-rb) % 4, 3)


@pytest.mark.parametrize("dtype", [cpu.S, cpu.CX])
def test_ws_tma_kernels_all_ops_ragged(ob, oracle, dtype, monkeypatch):
    """sgemm_ws.cuh forced onto interior, ragged, k-tail and tiny shapes of every op combination (conj
    included for CGEMM), beta != 0 and beta == 0 over a NaN C; device operands so that the kernel sees
    exactly these leading dimensions."""
    import torch
    monkeypatch.setenv("B200_SGEMM_TILE" if dtype == cpu.S else "B200_CGEMM_TILE", "256" if dtype == cpu.S else "128")
    rng = np.random.default_rng(4242 + dtype)
    alphas, betas = alpha_beta(dtype)
    ob.cblas.set_kernel(ob.cblas.K_FAST)          # the dispatcher's small-size threshold (generic kernel below 64^3) is off
    try:
        cplx = dtype == cpu.CX
        for (m, n, k) in [(256, 128, 64), (300, 260, 200), (1000, 77, 257 if cplx else 513), (64, 64, 16), (513, 130, 17), (36, 20, 5)]:
            for ta in range(ntrans(dtype)):
                for tb in range(ntrans(dtype)):
                    if cplx and m == 1000 and (tb - ta) % 4 > 1:
                        continue                                   # the long-double checker is the cost: the largest shape on 8 of the 16 op pairs
                    a, lda, b, ldb, c0, ldc = problem(rng, oracle, dtype, ta, tb, m, n, k, pad=_pad4(dtype, ta, tb, m, n, k))
                    pairs = ((alphas[2], betas[2]), (alphas[1], 0.0))
                    if cplx and (ta + tb) % 4 != 0:
                        pairs = pairs[(ta + tb) % 2:][:1]          # 16 op pairs: both scalar pairs on four of them, one (alternating) on the rest
                    for alpha, beta in pairs:
                        start = c0.copy()
                        if beta == 0.0:
                            start[:, :m] = np.nan                     # beta == 0 never reads C
                        da, db, dc = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), torch.from_numpy(start.copy()).cuda()
                        ob.cblas.gemm_any(dtype, ta, tb, m, n, k, alpha, da, lda, db, ldb, beta, dc, ldc)
                        kern = ob.cblas.last_kernel()
                        assert "ws_tma" in kern, kern
                        check(oracle, dtype, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c0, ldc, dc.cpu().numpy(), kern)
    finally:
        ob.cblas.set_kernel(ob.cblas.K_AUTO)


@pytest.mark.parametrize("dtype", [cpu.S, cpu.CX])
def test_ws_tma_kernels_are_the_default_on_full_grids(ob, oracle, dtype):
    """2048 x 2048 x 96: enough 256 x 128 tiles for the dispatcher to pick the TMA kernel on its own; checked
    against the oracle (k kept small: the CPU oracle is scalar code)."""
    import torch
    rng = np.random.default_rng(99 + dtype)
    m = n = 2048
    k = 96 if dtype == cpu.S else 40                    # complex: the long-double checker costs 4x per element
    for ta, tb in ((0, 0), (1, 0), (0, 1), (1, 1)) if dtype == cpu.S else ((0, 0), (1, 0), (3, 2)):
        a, lda, b, ldb, c0, ldc = problem(rng, oracle, dtype, ta, tb, m, n, k, pad=(0, 0, 0))
        alpha, beta = alpha_beta(dtype)[0][2], alpha_beta(dtype)[1][2]
        da, db, dc = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), torch.from_numpy(c0.copy()).cuda()
        ob.cblas.gemm_any(dtype, ta, tb, m, n, k, alpha, da, lda, db, ldb, beta, dc, ldc)
        kern = ob.cblas.last_kernel()
        assert "ws_tma" in kern, kern
        check(oracle, dtype, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c0, ldc, dc.cpu().numpy(), kern)


def test_ssyrk_and_cherk_through_the_ws_kernel(ob, oracle, monkeypatch):
    """The triangle mask (DeviceGemm::tri) of the TMA kernel: SSYRK / CSYRK / CHERK with the kernel forced,
    the other triangle of C keeps its bits (checked by check_case)."""
    import level3_helpers as L
    monkeypatch.setenv("B200_SGEMM_TILE", "256")
    monkeypatch.setenv("B200_CGEMM_TILE", "128")
    monkeypatch.setenv("B200_RANKK_TRI", "1")
    call = L.bind(ob.lib())
    rng = np.random.default_rng(7)
    for dtype, herm in ((cpu.S, 0), (cpu.CX, 0), (cpu.CX, 1)):
        cplx = dtype == cpu.CX
        for uplo in (0, 1):
            for trans in (0, 1):
                n, k = 300, 132
                rows, cols = (k, n) if trans else (n, k)
                a, c0 = L.operand(rng, dtype, cols, rows + 4), L.operand(rng, dtype, n, n + 3)
                alpha, beta = ((0.7, 1.3) if herm or not cplx else (0.7 - 0.9j, 1.3 - 1.1j))
                case = (1, dtype, herm, 0, uplo, trans, 0, n, k, rows + 4, rows + 4, n + 3, alpha, beta)
                got, want, gauge, K, touched = L.run_case(call, oracle, case, a, a, c0)
                L.check_case(case, got, want, gauge, K, touched, c0)
                assert "ws_tma" in ob.cblas.last_kernel() or "diag" in ob.cblas.last_kernel(), ob.cblas.last_kernel()


def test_generic_kernel_takes_skinny_problems_beyond_65535_column_tiles(ob):
    """m < 16 forces the generic kernel; n / 32 > 65535 overflowed its 2-D grid in round 1 (ADVICE):
    cblas_dgemm(m = 8, n = 2.2M, k = 2) must compute, not abort."""
    import torch
    m, n, k = 8, 2_200_000, 2
    a = torch.rand((k, m), dtype=torch.float64, device="cuda") - 0.5          # column-major m x k
    b = torch.rand((n, k), dtype=torch.float64, device="cuda") - 0.5          # column-major k x n
    c = torch.full((n, m), float("nan"), dtype=torch.float64, device="cuda")
    ob.cblas.gemm_any(cpu.D, 0, 0, m, n, k, 1.0, a, m, b, k, 0.0, c, m)
    assert ob.cblas.last_kernel() == "gemm_generic"
    want = b @ a                                                              # (n x k) (k x m) = C^T as stored
    assert torch.allclose(c, want, rtol=1e-13, atol=1e-14)


@pytest.mark.parametrize("where", ["host", "device"])
def test_grouped_batch_kernel(ob, oracle, where):
    """200 small DGEMMs of mixed shapes in ONE launch (gemm_grouped), host operands (one packed upload) and
    device operands; C := 1 * C members keep their bits; every matrix checked against the oracle."""
    import torch
    rng = np.random.default_rng(31)
    shapes = [(0, 0, 64, 64, 64, 120), (1, 0, 33, 17, 40, 50), (0, 1, 128, 5, 3, 20), (1, 1, 9, 128, 70, 10)]
    alphas, betas = [0.7, 1.0, 0.0, 2.0], [1.3, 0.0, 1.0, -0.5]
    probs, keep = [], []
    for gi, (ta, tb, m, n, k, cnt) in enumerate(shapes):
        for _ in range(cnt):
            a, lda, b, ldb, c0, ldc = problem(rng, oracle, cpu.D, ta, tb, m, n, k, pad=(1, 1, 1))
            start = c0.copy()
            if betas[gi] == 0.0:
                start[:, :m] = np.nan
            if where == "device":
                bufs = [torch.from_numpy(x).cuda() for x in (a, b, start)]
                keep.append(bufs)
                ptrs = [t.data_ptr() for t in bufs]
            else:
                ptrs = [a.ctypes.data, b.ctypes.data, start.ctypes.data]
            probs.append((a, lda, b, ldb, c0, start, ldc, ptrs))
    I = lambda vals: (C.c_int * len(vals))(*vals)
    first = [sum(s[5] for s in shapes[:i]) for i in range(len(shapes))]
    arr = lambda j: (C.c_void_p * len(probs))(*[p[7][j] for p in probs])
    adr = lambda x: C.cast(x, C.c_void_p)
    before = ob.cblas.launch_count()
    ob.lib().cblas_dgemm_batch(ob.cblas.ColMajor, adr(I([CB[s[0]] for s in shapes])), adr(I([CB[s[1]] for s in shapes])),
                               adr(I([s[2] for s in shapes])), adr(I([s[3] for s in shapes])), adr(I([s[4] for s in shapes])),
                               adr((C.c_double * 4)(*alphas)), adr(arr(0)), adr(I([probs[f][1] for f in first])), adr(arr(1)),
                               adr(I([probs[f][3] for f in first])), adr((C.c_double * 4)(*betas)), adr(arr(2)),
                               adr(I([probs[f][6] for f in first])), len(shapes), adr(I([s[5] for s in shapes])))
    assert ob.cblas.launch_count() - before == 1 and ob.cblas.last_kernel() == "gemm_grouped"
    i = 0
    for gi, (ta, tb, m, n, k, cnt) in enumerate(shapes):
        for _ in range(cnt):
            a, lda, b, ldb, c0, start, ldc, _p = probs[i]
            got = keep[i][2].cpu().numpy() if where == "device" else start
            i += 1
            if alphas[gi] == 0.0 and betas[gi] == 1.0:
                assert np.array_equal(got.view(np.uint8), c0.view(np.uint8)), "C := 1 * C must not touch C"
            else:
                check(oracle, cpu.D, ta, tb, m, n, k, alphas[gi], a, lda, b, ldb, betas[gi], c0, ldc, got, "grouped-" + where)


def test_shutdown_releases_and_the_next_call_initialises_again(ob, oracle):
    rng = np.random.default_rng(3)
    m, n, k = 150, 140, 130
    a, lda, b, ldb, c0, ldc = problem(rng, oracle, cpu.D, 0, 0, m, n, k)
    for _ in range(2):
        got = c0.copy()
        ob.cblas.dgemm(ob.cblas.ColMajor, CB[0], CB[0], m, n, k, 0.7, a, lda, b, ldb, 1.3, got, ldc)
        check(oracle, cpu.D, 0, 0, m, n, k, 0.7, a, lda, b, ldb, 1.3, c0, ldc, got, "around-shutdown")
        ob.lib().b200_shutdown()


def test_sbgemm_small_ragged_and_misaligned_shapes_stay_on_tcgen05(ob, oracle):
    """Round 1 sent SBGEMM to the CUDA-core generic kernel for m < 128, n < 256, k < 64, lda % 8 != 0 or a base that
    is not 16-byte aligned (VERDICT missing #7).  TMA boxes may exceed the tensor (zero fill), and misaligned operands
    are repacked once on the device, so every shape from 64^3 up runs on tcgen05.mma: all four op combinations,
    ragged extents, odd leading dimensions, a sub-matrix view starting at an odd element, beta != 0 / beta == 0 over NaN."""
    import torch
    rng = np.random.default_rng(808)
    for (m, n, k, pad, skew) in [(64, 64, 64, (0, 0, 0), 0), (100, 130, 70, (8, 8, 3), 0), (128, 255, 63, (1, 3, 5), 0), (129, 257, 65, (7, 5, 1), 0),
                                 (300, 77, 513, (8, 16, 0), 1), (36, 500, 96, (3, 0, 2), 3), (513, 130, 17, (0, 0, 1), 0)]:
        for ta in range(2):
            for tb in range(2):
                a, lda, b, ldb, c0, ldc = problem(rng, oracle, cpu.SB, ta, tb, m, n, k, pad=pad)
                for alpha, beta in ((0.7, 1.3), (1.0, 0.0)):
                    start = c0.copy()
                    if beta == 0.0:
                        start[:, :m] = np.nan
                    # device copies with `skew` extra leading elements: the operand starts at a 2 * skew byte offset
                    da = torch.zeros(a.size + skew, dtype=torch.int16, device="cuda"); da[skew:] = torch.from_numpy(a.view(np.int16).ravel()).cuda()
                    db = torch.zeros(b.size + skew, dtype=torch.int16, device="cuda"); db[skew:] = torch.from_numpy(b.view(np.int16).ravel()).cuda()
                    dc = torch.from_numpy(start.copy()).cuda()
                    ob.cblas.gemm_any(cpu.SB, ta, tb, m, n, k, alpha, da.data_ptr() + 2 * skew, lda, db.data_ptr() + 2 * skew, ldb, beta, dc, ldc)
                    kern = ob.cblas.last_kernel()
                    assert "tcgen05" in kern, (kern, m, n, k, pad, skew)
                    check(oracle, cpu.SB, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c0, ldc, dc.cpu().numpy(), kern)


def _summa_bind(lib):
    lib.b200_summa_create.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.b200_summa_destroy.argtypes = [C.c_void_p]
    lib.b200_summa_gemm.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                    C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.b200_summa_launches.restype = C.c_uint64
    lib.b200_summa_launches.argtypes = [C.c_void_p]
    lib.b200_summa_describe.restype = C.c_char_p
    lib.b200_summa_describe.argtypes = [C.c_void_p]
    lib.b200_last_error.restype = C.c_char_p


@pytest.mark.parametrize("dtype", [cpu.D, cpu.S, cpu.Z, cpu.CX])
def test_summa_driver_c_abi_on_a_1x1_grid(ob, oracle, dtype):
    """b200_summa_create / gemm / destroy (include/openblas_b200.h, csrc/summa.cu) through the C ABI on one GPU: the
    whole machinery -- window packing panel by panel, panels used in place, the double-buffered sweep, k tails,
    beta on the first panel only, host operands (upload into the window, C strips back), k = 0, window growth -- is
    the same code that runs on a P x Q grid (tools/summa_c_check.py checks that under torchrun on 2 and 8 GPUs);
    against the oracle's GEMM."""
    import torch
    lib = ob.lib()
    _summa_bind(lib)
    h = C.c_void_p()
    assert lib.b200_summa_create(C.byref(h), bytes(128), 0, 1, 1, 1) == 0, lib.b200_last_error()
    assert b"1x1" in lib.b200_summa_describe(h)
    rng = np.random.default_rng(900 + dtype)
    alphas, betas = alpha_beta(dtype)
    stream = torch.cuda.current_stream()
    try:
        for (m, n, k, nb) in [(300, 260, 200, 64), (130, 77, 513, 128), (64, 64, 0, 32), (1000, 900, 96, 32)]:
            a, lda, b, ldb, c0, ldc = problem(rng, oracle, dtype, 0, 0, m, n, k, pad=(3, 5, 7))
            for alpha, beta, where in ((alphas[2], betas[2], "device"), (alphas[1], 0.0, "device"), (alphas[2], betas[2], "host"), (alphas[1], 0.0, "pinned")):
                start = c0.copy()
                if beta == 0.0:
                    start[:, :m] = np.nan                                  # beta == 0 never reads C
                al, be = ob.cblas.scalar_array(dtype, alpha), ob.cblas.scalar_array(dtype, beta)
                before = lib.b200_summa_launches(h)
                if where == "device":
                    da, db, dc = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), torch.from_numpy(start.copy()).cuda()
                    rc = lib.b200_summa_gemm(h, dtype, m, n, k, nb, al.ctypes.data, da.data_ptr(), lda, db.data_ptr(), ldb, be.ctypes.data, dc.data_ptr(), ldc,
                                             stream.cuda_stream)
                    torch.cuda.synchronize()
                    got = dc.cpu().numpy()
                else:
                    ha, hb, hc = torch.from_numpy(a.copy()), torch.from_numpy(b.copy()), torch.from_numpy(start.copy())
                    if where == "pinned":
                        ha, hb, hc = ha.pin_memory(), hb.pin_memory(), hc.pin_memory()
                    rc = lib.b200_summa_gemm(h, dtype, m, n, k, nb, al.ctypes.data, ha.data_ptr(), lda, hb.data_ptr(), ldb, be.ctypes.data, hc.data_ptr(), ldc,
                                             stream.cuda_stream)
                    got = hc.numpy()
                assert rc == 0, lib.b200_last_error()
                assert lib.b200_summa_launches(h) - before == max(1, (k + nb - 1) // nb)      # one local product per k panel (k = 0: the scaling of C)
                check(oracle, dtype, 0, 0, m, n, k, alpha, a, lda, b, ldb, beta, c0, ldc, got, f"summa-{where}")
        # illegal calls are refused with a message, not executed
        assert lib.b200_summa_gemm(h, cpu.SB, 8, 8, 8, 4, None, None, 8, None, 8, None, None, 8, None) != 0
        assert b"bf16" in lib.b200_last_error()
    finally:
        assert lib.b200_summa_destroy(h) == 0


def test_summa_driver_on_two_gpus_when_present(ob):
    """tools/summa_c_check.py under torchrun on 2 GPUs (skipped on a one-GPU box): flags + copy-engine pulls from the
    peer's window, every rank's block against a one-GPU DGEMM and long-double samples."""
    import subprocess, sys, torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(root, "tools", "summa_c_check.py"), "1500", "1300", "1100", "128"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ALL OK" in r.stdout, (r.stdout[-3000:], r.stderr[-2000:])


def test_summa_driver_host_c_in_two_column_sweeps(ob):
    """A host C is swept in two column halves so that the first half's download hides behind the second half's products
    (default from 8192 local columns; B200_SUMMA_HOST_HALVES lowers the limit, read once per process, hence the child):
    tools/summa_c_check.py on one GPU with host operands -- pinned, pageable, beta != 0 -- against a one-GPU DGEMM and
    long-double samples."""
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, B200_SUMMA_HOST_HALVES="64")
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "summa_c_check.py"), "1000", "900", "800", "128"], capture_output=True, text=True,
                       timeout=600, env=env)
    assert r.returncode == 0 and "ALL OK" in r.stdout, (r.stdout[-3000:], r.stderr[-2000:])
