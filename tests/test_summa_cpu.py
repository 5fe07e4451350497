"""CPU tests of the multi-GPU layer's HOST logic (openblas_b200/summa.py) with 2, 4 and 8 processes on the
gloo backend: grid shape, block-cyclic ownership maps, the panel schedule, and a full SUMMA sweep
whose local product is done by the CPU oracle (test-only injection; on GPUs it is the library's
b200_gemm_async).  The distributed result must equal the single-process oracle result."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openblas_b200 import summa  # noqa: E402


def test_grid_shapes_follow_divide_rule():
    # gemm_thread_mn.c:43-61 divide_rule: squarest grid, more columns than rows
    assert summa.grid_shape(1) == (1, 1)
    assert summa.grid_shape(2) == (1, 2)
    assert summa.grid_shape(4) == (2, 2)
    assert summa.grid_shape(8) == (2, 4)
    assert summa.grid_shape(6) == (2, 3)


def test_block_cyclic_ownership_partitions_every_index():
    for n, nb, p in [(10, 3, 2), (32768, 2048, 4), (17, 5, 3), (4, 8, 2), (0, 4, 2)]:
        seen = []
        for ip in range(p):
            idx = summa.local_index_map(n, nb, ip, p)
            assert len(idx) == summa.numroc(n, nb, ip, p)
            seen += idx
        assert sorted(seen) == list(range(n))


def test_panel_schedule_has_one_owner_per_panel():
    k, nb, P, Q = 23, 4, 2, 3
    steps = summa.panel_schedule(k, nb, P, Q)
    assert sum(w for _, w, *_ in steps) == k
    for k0, w, qa, ca, pb, rb in steps:
        a_cols = summa.local_index_map(k, nb, qa, Q)
        b_rows = summa.local_index_map(k, nb, pb, P)
        assert a_cols[ca:ca + w] == list(range(k0, k0 + w))
        assert b_rows[rb:rb + w] == list(range(k0, k0 + w))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, m, n, k, nb, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import cpu
    oracle = cpu.Oracle()
    rng = np.random.default_rng(42)          # same global matrices on every rank
    A = torch.from_numpy(rng.random((k, m)) - 0.5)   # (cols, rows) storage of m x k
    B = torch.from_numpy(rng.random((n, k)) - 0.5)
    C = torch.from_numpy(rng.random((n, m)) - 0.5)
    grid = summa.make_grid(world, rank)
    a_loc = summa.scatter_from_global(A, nb, grid, rows_by="p", cols_by="q")
    b_loc = summa.scatter_from_global(B, nb, grid, rows_by="p", cols_by="q")
    c_loc = summa.scatter_from_global(C, nb, grid, rows_by="p", cols_by="q")

    def local_gemm(mm, nn, kk, alpha, a, lda, b, ldb, beta, c, ldc, stream):
        oracle.gemm(cpu.D, 0, 0, mm, nn, kk, alpha, a.numpy(), lda, b.numpy(), ldb, beta, c.numpy(), ldc)

    sm = summa.Summa(grid, m, n, k, nb, torch.float64, "cpu", local_gemm)
    assert (sm.m_loc, sm.n_loc) == (c_loc.shape[1], c_loc.shape[0])
    sm.run(0.7, a_loc, b_loc, 1.3, c_loc)
    want = C.numpy().copy()
    oracle.gemm(cpu.D, 0, 0, m, n, k, 0.7, A.numpy(), m, B.numpy(), k, 1.3, want, m)
    ri = summa.local_index_map(m, nb, grid.p, grid.P)
    ci = summa.local_index_map(n, nb, grid.q, grid.Q)
    ref_loc = want[np.ix_(ci, ri)]
    err = float(np.max(np.abs(ref_loc - c_loc.numpy()))) if ref_loc.size else 0.0
    out[rank] = (err, sm.launches, len(sm.steps))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,shape", [(2, (37, 29, 23, 4)), (2, (16, 16, 16, 8)), (2, (5, 40, 9, 3)),
                                         (4, (37, 29, 23, 4)), (4, (9, 7, 30, 2)), (8, (41, 53, 19, 3))])
def test_summa_gloo_matches_oracle(world, shape):
    """1 x 2, 2 x 2 and 2 x 4 process grids (the grids bench.py runs at N = 2, 4, 8): both broadcast
    directions, ranks that own no block of a short dimension, ragged last blocks."""
    m, n, k, nb = shape
    port = _free_port()
    mgr = mp.get_context("spawn").Manager()      # never fork this (by now multi-threaded) test process
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, m, n, k, nb, out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        err, launches, steps = out[rank]
        assert err < 1e-12, (rank, err)
        assert launches == steps == (k + nb - 1) // nb


def test_c_driver_layout_functions_agree_with_the_python_mirror():
    """b200_summa_numroc / b200_summa_schedule / b200_summa_grid (csrc/summa.cu; no CUDA call behind them) against
    numroc / panel_schedule / GRID of openblas_b200/summa.py, which the gloo sweeps above exercise."""
    import ctypes as C
    import openblas_b200 as ob
    from openblas_b200 import summa
    L = ob.lib()
    L.b200_summa_numroc.restype = C.c_int64
    L.b200_summa_numroc.argtypes = [C.c_int64, C.c_int64, C.c_int, C.c_int]
    L.b200_summa_schedule.restype = C.c_int64
    L.b200_summa_schedule.argtypes = [C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int64] + [C.c_void_p] * 6
    for n in (0, 1, 5, 100, 1000, 32768, 65537):
        for nb in (1, 7, 64, 2048):
            for P in (1, 2, 3, 4, 8):
                got = [L.b200_summa_numroc(n, nb, i, P) for i in range(P)]
                assert got == [summa.numroc(n, nb, i, P) for i in range(P)] and sum(got) == n
                assert got == [len(summa.local_index_map(n, nb, i, P)) for i in range(P)]
    for world, want in summa.GRID.items():
        P, Q = C.c_int(), C.c_int()
        L.b200_summa_grid(world, C.byref(P), C.byref(Q))
        assert (P.value, Q.value) == want
    for (k, nb, P, Q) in ((0, 4, 2, 2), (10, 4, 2, 4), (32768, 4096, 2, 4), (1000, 64, 3, 5), (7, 8, 1, 2)):
        steps = summa.panel_schedule(k, nb, P, Q)
        cap = len(steps) + 2
        k0, w, al, bl = ((C.c_int64 * cap)() for _ in range(4))
        ao, bo = (C.c_int * cap)(), (C.c_int * cap)()
        cnt = L.b200_summa_schedule(k, nb, P, Q, cap, k0, w, ao, al, bo, bl)
        assert cnt == len(steps)
        assert [(k0[i], w[i], ao[i], al[i], bo[i], bl[i]) for i in range(cnt)] == steps


def test_c_driver_rejects_bad_handles_and_grids_before_touching_cuda():
    """b200_summa_create / b200_summa_gemm / b200_summa_describe with arguments that can never work: an error code and
    the reason in b200_last_error, no handle, no CUDA call (this runs without a GPU)."""
    import ctypes as C
    import openblas_b200 as ob
    L = ob.lib()
    L.b200_summa_create.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.b200_summa_describe.restype = C.c_char_p
    L.b200_summa_describe.argtypes = [C.c_void_p]
    L.b200_summa_destroy.argtypes = [C.c_void_p]
    L.b200_last_error.restype = C.c_char_p
    ident = bytes(128)
    for rank, world, P, Q in ((0, 4, 2, 3), (4, 4, 2, 2), (-1, 2, 1, 2), (0, 0, 0, 0), (0, 128, 8, 16), (0, 2, -1, -2)):
        h = C.c_void_p(0x1234)
        assert L.b200_summa_create(C.byref(h), ident, rank, world, P, Q) != 0
        assert h.value is None and b"bad rank / world / grid" in L.b200_last_error()
    h = C.c_void_p(0x1234)
    assert L.b200_summa_create(C.byref(h), None, 0, 2, 1, 2) != 0 and h.value is None and b"unique id" in L.b200_last_error()
    assert L.b200_summa_create(None, ident, 0, 1, 1, 1) != 0
    assert L.b200_summa_destroy(None) == 0
    assert b"no handle" in L.b200_summa_describe(None)
