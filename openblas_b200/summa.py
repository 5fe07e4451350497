"""2-D SUMMA over the GPUs of one box: the multi-GPU counterpart of the reference's threaded
level-3 driver.

What it replaces:
  driver/level3/level3_thread.c:804-862   choice of an nthreads_m x nthreads_n grid
  driver/level3/gemm_thread_mn.c:43-61    divide_rule[] (8 workers -> 2 x 4)
  level3_thread.c:357-451                 each worker packs a B panel, PUBLISHES it to the workers
                                          of its group through job[].working flags, double
                                          buffered (DIVIDE_RATE 2)
Here a worker is one process driving one GPU.  C is stationary and distributed 2-D
block-cyclically (block NB) over a P x Q grid; A and B are distributed the same way.  For each
k panel the owning grid column broadcasts its slice of A along grid rows and the owning grid
row broadcasts its slice of B along grid columns (NCCL over NVLink, on a communication stream,
double buffered) while the local GEMM  C_loc += A_panel * B_panel  of the PREVIOUS panel runs on
the compute stream through the library's device entry point (b200_gemm_async -> the DMMA /
FFMA / tcgen05 kernels).  No reduction is needed: k is never split across GPUs.

Matrices are column-major.  A local matrix is held as a 2-D torch tensor of shape (cols, rows)
(C-contiguous), i.e. element (i, j) is t[j, i] and the leading dimension is t.shape[1].

torch.distributed is plumbing only (process groups, NCCL broadcast); the arithmetic is ours.
"""
import os
from dataclasses import dataclass

import torch
import torch.distributed as dist

# grid shapes: squarest grid with P <= Q, cf. divide_rule[] in gemm_thread_mn.c:43-61
GRID = {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (2, 4), 16: (4, 4)}


def grid_shape(world):
    forced = os.environ.get("B200_SUMMA_GRID")      # e.g. "2x1": testing / odd topologies
    if forced:
        P, Q = (int(x) for x in forced.lower().split("x"))
        assert P * Q == world, f"B200_SUMMA_GRID={forced} does not match world size {world}"
        return P, Q
    if world in GRID:
        return GRID[world]
    p = int(world ** 0.5)
    while world % p:
        p -= 1
    return p, world // p


def numroc(n, nb, iproc, nprocs):
    """Number of rows/cols of a block-cyclically distributed dimension owned by process iproc
    (ScaLAPACK's NUMROC with source process 0)."""
    nblocks = n // nb
    base = (nblocks // nprocs) * nb
    extra = nblocks % nprocs
    if iproc < extra:
        base += nb
    elif iproc == extra:
        base += n % nb
    return base


def local_index_map(n, nb, iproc, nprocs):
    """Global indices, in local order, of the entries process iproc owns."""
    idx = []
    for b0 in range(iproc * nb, n, nb * nprocs):
        idx.extend(range(b0, min(b0 + nb, n)))
    return idx


@dataclass
class Grid:
    P: int
    Q: int
    rank: int
    p: int            # my grid row
    q: int            # my grid column
    row_group: object  # ranks sharing my grid row (vary q)
    col_group: object  # ranks sharing my grid column (vary p)
    row_ranks: list
    col_ranks: list


def make_grid(world=None, rank=None):
    """Row-major rank -> (p, q) mapping; creates the row and column communicators."""
    world = dist.get_world_size() if world is None else world
    rank = dist.get_rank() if rank is None else rank
    P, Q = grid_shape(world)
    p, q = rank // Q, rank % Q
    row_group = col_group = None
    row_ranks = col_ranks = None
    for pp in range(P):
        ranks = [pp * Q + qq for qq in range(Q)]
        g = dist.new_group(ranks) if world > 1 else None
        if pp == p:
            row_group, row_ranks = g, ranks
    for qq in range(Q):
        ranks = [pp * Q + qq for pp in range(P)]
        g = dist.new_group(ranks) if world > 1 else None
        if qq == q:
            col_group, col_ranks = g, ranks
    return Grid(P, Q, rank, p, q, row_group, col_group, row_ranks, col_ranks)


def panel_schedule(k, nb, P, Q):
    """The k panels of one SUMMA sweep: (k0, width, owner grid column of the A slice, local
    column offset there, owner grid row of the B slice, local row offset there).  A panel never
    crosses a distribution block, so each slice has exactly one owner."""
    steps = []
    for k0 in range(0, k, nb):
        w = min(nb, k - k0)
        blk = k0 // nb
        steps.append((k0, w, blk % Q, (blk // Q) * nb, blk % P, (blk // P) * nb))
    return steps


class Summa:
    """C_loc <- alpha * sum_k A_panel(k) * B_panel(k) + beta * C_loc on a P x Q grid.

    local_gemm(m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, stream) performs the column-major
    NN product on this rank; on GPUs it is openblas_b200.cblas.gemm_device.  Tests inject a CPU
    function to exercise the schedule under the gloo backend.
    """

    def __init__(self, grid, m, n, k, nb, dtype, device, local_gemm, comm_stream=None, compute_stream=None):
        self.g, self.m, self.n, self.k, self.nb = grid, m, n, k, nb
        self.dtype, self.device, self.local_gemm = dtype, device, local_gemm
        self.m_loc = numroc(m, nb, grid.p, grid.P)
        self.n_loc = numroc(n, nb, grid.q, grid.Q)
        self.ka_loc = numroc(k, nb, grid.q, grid.Q)   # columns of A held here
        self.kb_loc = numroc(k, nb, grid.p, grid.P)   # rows of B held here
        self.steps = panel_schedule(k, nb, grid.P, grid.Q)
        self.cuda = torch.device(device).type == "cuda"
        # double-buffered panel landing zones (DIVIDE_RATE 2 in level3_thread.c:44-46)
        self.a_buf = [torch.empty((nb, max(1, self.m_loc)), dtype=dtype, device=device) for _ in range(2)]
        self.b_buf = [torch.empty((max(1, self.n_loc), nb), dtype=dtype, device=device) for _ in range(2)]
        if self.cuda:
            self.comm_stream = comm_stream or torch.cuda.Stream(device=device)
            self.compute_stream = compute_stream or torch.cuda.current_stream(device)
            self.ready = [torch.cuda.Event() for _ in range(2)]     # panel landed
            self.consumed = [torch.cuda.Event() for _ in range(2)]  # panel buffer free again
        self.launches = 0

    def _post_panel(self, step, slot, a_loc, b_loc):
        """Owner packs, everybody receives panel `step` into buffer `slot`."""
        g = self.g
        k0, w, qa, ca, pb, rb = self.steps[step]
        a_pan = self.a_buf[slot][:w]           # (w, m_loc): w columns of A
        b_pan = self.b_buf[slot][:, :w]        # (n_loc, w) view: w rows of B, ld = nb
        if g.q == qa:
            a_pan.copy_(a_loc[ca:ca + w])
        if g.p == pb:
            b_pan.copy_(b_loc[:, rb:rb + w])
        if g.Q > 1:
            dist.broadcast(a_pan, src=g.row_ranks[qa], group=g.row_group)
        if g.P > 1:
            # NCCL wants a dense buffer: broadcast the whole (n_loc, nb) slot when the view is strided
            if w == self.nb:
                dist.broadcast(self.b_buf[slot], src=g.col_ranks[pb], group=g.col_group)
            else:
                dense = b_pan.contiguous()
                dist.broadcast(dense, src=g.col_ranks[pb], group=g.col_group)
                b_pan.copy_(dense)
        return w

    def run(self, alpha, a_loc, b_loc, beta, c_loc):
        """a_loc: (ka_loc, m_loc), b_loc: (n_loc, kb_loc), c_loc: (n_loc, m_loc) torch tensors."""
        nsteps = len(self.steps)
        if self.m_loc == 0 or self.n_loc == 0:
            # still take part in the broadcasts of my row / column
            for s in range(nsteps):
                self._post_panel(s, s % 2, a_loc, b_loc)
            return
        if not self.cuda:
            for s in range(nsteps):
                w = self._post_panel(s, s % 2, a_loc, b_loc)
                self.local_gemm(self.m_loc, self.n_loc, w, alpha, self.a_buf[s % 2], self.m_loc, self.b_buf[s % 2],
                                self.nb, beta if s == 0 else 1.0, c_loc, self.m_loc, None)
                self.launches += 1
            if nsteps == 0:
                self.local_gemm(self.m_loc, self.n_loc, 0, alpha, self.a_buf[0], self.m_loc, self.b_buf[0], self.nb,
                                beta, c_loc, self.m_loc, None)
            return

        cs, ms = self.comm_stream, self.compute_stream
        cs.wait_stream(ms)   # operands written on the compute stream are visible to the packer
        with torch.cuda.stream(cs):
            self._post_panel(0, 0, a_loc, b_loc)
            self.ready[0].record(cs)
        for s in range(nsteps):
            slot = s % 2
            if s + 1 < nsteps:
                nslot = (s + 1) % 2
                with torch.cuda.stream(cs):
                    if s >= 1:
                        cs.wait_event(self.consumed[nslot])   # GEMM of step s-1 released this buffer
                    self._post_panel(s + 1, nslot, a_loc, b_loc)
                    self.ready[nslot].record(cs)
            ms.wait_event(self.ready[slot])
            w = self.steps[s][1]
            self.local_gemm(self.m_loc, self.n_loc, w, alpha, self.a_buf[slot], self.m_loc, self.b_buf[slot], self.nb,
                            beta if s == 0 else 1.0, c_loc, self.m_loc, ms.cuda_stream)
            self.launches += 1
            self.consumed[slot].record(ms)
        if nsteps == 0:
            self.local_gemm(self.m_loc, self.n_loc, 0, alpha, self.a_buf[0], self.m_loc, self.b_buf[0], self.nb, beta,
                            c_loc, self.m_loc, ms.cuda_stream)
        ms.wait_stream(cs)


def scatter_from_global(full, nb, grid, rows_by, cols_by):
    """Take this rank's block-cyclic piece of a global column-major matrix given as a (cols, rows)
    tensor.  rows_by / cols_by: ('p' | 'q') which grid coordinate distributes that dimension."""
    coord = {"p": (grid.p, grid.P), "q": (grid.q, grid.Q)}
    ri = local_index_map(full.shape[1], nb, *coord[rows_by])
    ci = local_index_map(full.shape[0], nb, *coord[cols_by])
    ri_t = torch.tensor(ri, dtype=torch.long, device=full.device)
    ci_t = torch.tensor(ci, dtype=torch.long, device=full.device)
    return full.index_select(0, ci_t).index_select(1, ri_t).contiguous()


# ----------------------------------------------------------------------------------------------------------
# The driver inside the library (csrc/summa.cu) through its C symbols; the class above is its host-side mirror
# (same grid, same distribution, same schedule -- tests/test_summa_cpu.py checks them against each other) and
# the path that runs under gloo on CPU.
class CSumma:
    """b200_summa_* of libopenblas_b200.so: SUMMA over peer windows (CUDA IPC + copy engines + stream memory
    operations), NCCL bootstrapped from an id this class distributes with torch.distributed."""

    def __init__(self, world=None, rank=None, device=None):
        import ctypes as C
        from ._lib import lib
        self.C, self.lib = C, lib()
        L = self.lib
        L.b200_summa_create.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.b200_summa_destroy.argtypes = [C.c_void_p]
        L.b200_summa_gemm.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                      C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.b200_summa_launches.restype = C.c_uint64
        L.b200_summa_launches.argtypes = [C.c_void_p]
        L.b200_summa_describe.restype = C.c_char_p
        L.b200_summa_describe.argtypes = [C.c_void_p]
        L.b200_last_error.restype = C.c_char_p
        self.world = dist.get_world_size() if world is None else world
        self.rank = dist.get_rank() if rank is None else rank
        self.P, self.Q = grid_shape(self.world)
        self.p, self.q = self.rank // self.Q, self.rank % self.Q
        ident = torch.zeros(128, dtype=torch.uint8)
        if self.world > 1:
            if self.rank == 0:
                buf = (C.c_char * 128)()
                if L.b200_summa_unique_id(buf):
                    raise RuntimeError(L.b200_last_error().decode())
                ident = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
            dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
            on_dev = ident.to(dev) if dist.get_backend() == "nccl" else ident
            dist.broadcast(on_dev, src=0)
            ident = on_dev.cpu()
        raw = bytes(ident.numpy().tobytes())
        self.handle = C.c_void_p()
        if L.b200_summa_create(C.byref(self.handle), raw, self.rank, self.world, self.P, self.Q):
            raise RuntimeError(L.b200_last_error().decode())

    def local_shapes(self, m, n, k, nb):
        """(m_loc, n_loc, ka_loc, kb_loc) of this rank"""
        return (numroc(m, nb, self.p, self.P), numroc(n, nb, self.q, self.Q), numroc(k, nb, self.q, self.Q), numroc(k, nb, self.p, self.P))

    def gemm(self, dtype_code, m, n, k, nb, alpha, a_loc, lda, b_loc, ldb, beta, c_loc, ldc, stream=None):
        """a_loc / b_loc / c_loc: torch tensors (device, or pinned / pageable host) or raw addresses"""
        from .cblas import addr, scalar_array
        al, be = scalar_array(dtype_code, alpha), scalar_array(dtype_code, beta)
        if self.lib.b200_summa_gemm(self.handle, dtype_code, m, n, k, nb, al.ctypes.data, addr(a_loc), lda, addr(b_loc), ldb, be.ctypes.data,
                                    addr(c_loc), ldc, stream):
            raise RuntimeError(self.lib.b200_last_error().decode())

    def launches(self):
        return int(self.lib.b200_summa_launches(self.handle))

    def describe(self):
        return self.lib.b200_summa_describe(self.handle).decode()

    def close(self):
        if self.handle:
            self.lib.b200_summa_destroy(self.handle)
            self.handle = None


def hashed_entries(tag, gi, gj):
    """Deterministic pseudo-random entries in [-0.5, 0.5) of the GLOBAL matrix `tag` at rows gi x columns gj (int64
    torch tensors, any device): every rank can regenerate any row or column of A and B without communication, which
    is what lets each rank check sampled entries of its C block against long-double dot products.  Returns a
    (len(gj), len(gi)) float64 tensor (column-major storage of the piece)."""
    x = (gi[None, :] * 2654435761 + gj[:, None] * 40503 + tag * 7919) & 0xFFFFFFFF
    x = x ^ (x >> 16)
    x = (x * 0x45D9F3B) & 0xFFFFFFFF
    x = x ^ (x >> 16)
    return x.to(torch.float64) * (1.0 / 4294967296.0) - 0.5
