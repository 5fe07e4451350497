"""CPU tests of the library's HOST logic with no GPU: tests/hostsim/build.sh compiles the unchanged
interface_*.c and runtime.cu (with g++, -DB200_HOSTSIM: smaller size thresholds so that every staging
path is reached by matrices a CPU can multiply) against a host-memory stand-in for the CUDA runtime
(cuda_shim.cpp: "device" memory is poisoned with NaN bytes) and CPU stand-ins for the kernel launchers
(sim_kernels.cpp: every GEMM launch is the oracle, which is bit-identical to the reference's GENERIC
build).  What is exercised is everything between the C ABI and the kernel launch: argument
normalisation, pointer classification, packing / strided copies / pinned slot rings, the panel
pipeline, batch staging, the dispatcher's fallbacks, the block-column and masked-triangle schemes of
the SYRK family, the TRMM / TRSM recursion and its operand offsets.  The real kernels are covered by
the -m gpu tests."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import cpu
import level3_helpers as L
from helpers import ALL_DTYPES, NAMES, alpha_beta, ntrans, problem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sim():
    out = subprocess.check_output([os.path.join(ROOT, "tests", "hostsim", "build.sh")], text=True).strip().splitlines()[-1]
    lib = C.CDLL(out)
    lib.b200_last_kernel.restype = C.c_char_p
    lib.b200_launch_count.restype = C.c_uint64
    lib.hostsim_copy_count.restype = C.c_size_t
    lib.hostsim_device_alloc.restype = C.c_void_p
    lib.hostsim_device_alloc.argtypes = [C.c_size_t]
    lib.hostsim_pinned_alloc.restype = C.c_void_p
    lib.hostsim_pinned_alloc.argtypes = [C.c_size_t]
    lib.hostsim_free.argtypes = [C.c_void_p]
    return lib


def fgemm(lib, dtype, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc):
    fn = getattr(lib, cpu.DTYPE_NAMES[dtype] + "gemm_")
    al, be = cpu.scalar_bytes(dtype, alpha), cpu.scalar_bytes(dtype, beta)
    i = lambda v: C.byref(C.c_int(int(v)))
    P = lambda x: C.c_void_p(x) if isinstance(x, int) else x.ctypes.data_as(C.c_void_p)      # a bare int would be passed as a 32-bit int
    fn(C.c_char_p(cpu.TRANS_CHAR[ta].encode()), C.c_char_p(cpu.TRANS_CHAR[tb].encode()), i(m), i(n), i(k), cpu._ptr(al), P(a), i(lda),
       P(b), i(ldb), cpu._ptr(be), P(c), i(ldc))


def test_gemm_golden_vectors_bitwise_through_the_host_path(sim, golden):
    """All 88 reference-generated GEMM cases: host path + oracle must reproduce the REFERENCE's bytes."""
    for idx, row in enumerate(golden["meta"]):
        dtype, ta, tb, m, n, k, lda, ldb, ldc = (int(v) for v in row[:9])
        cplx = dtype in (cpu.CX, cpu.Z)
        alpha = complex(row[9], row[10]) if cplx else row[9]
        beta = complex(row[11], row[12]) if cplx else row[11]
        got = golden[f"c0_{idx}"].copy()
        fgemm(sim, dtype, ta, tb, m, n, k, alpha, golden[f"a{idx}"], lda, golden[f"b{idx}"], ldb, beta, got, ldc)
        assert np.array_equal(got.view(np.uint8), golden[f"c{idx}"].view(np.uint8)), (idx, NAMES[dtype], ta, tb, m, n, k)


@pytest.mark.parametrize("dtype", [cpu.D, cpu.Z, cpu.S, cpu.SB])
def test_every_staging_path_equals_the_oracle_bitwise(sim, oracle, dtype):
    """Small (one packed block), middle (strided copies, pinned slot ring for pageable memory) and
    pipelined (A row panels / B column panels / C blocks) host paths, every op combination, beta == 0
    over NaN and beta != 0, pageable, pinned and device operands.  Each element's k order is the same
    whatever the blocking of m and n, so the result must equal one oracle call on the whole problem."""
    rng = np.random.default_rng(40 + dtype)
    es_out = np.dtype(cpu.NP_OUT[dtype]).itemsize
    shapes = [(9, 7, 5), (70, 50, 33), (150, 130, 40), (300, 280, 12)]
    seen_multi = False
    for (m, n, k) in shapes:
        for ta in range(ntrans(dtype)):
            for tb in range(ntrans(dtype)):
                for beta_zero in (False, True):
                    a, lda, b, ldb, c0, ldc = problem(rng, oracle, dtype, ta, tb, m, n, k, pad=(1, 2, 3))
                    alpha, beta = alpha_beta(dtype)[0][2], (0.0 if beta_zero else alpha_beta(dtype)[1][2])
                    if beta_zero:
                        c0[:, :m] = np.nan
                    want = c0.copy()
                    oracle.gemm(dtype, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, want, ldc)
                    before = sim.b200_launch_count()
                    got = c0.copy()
                    fgemm(sim, dtype, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, got, ldc)
                    seen_multi |= sim.b200_launch_count() - before > 1
                    assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), (NAMES[dtype], ta, tb, m, n, k, beta_zero)
    assert seen_multi, "the panel pipeline (several block GEMMs per call) was never taken"
    # pinned and device operands: same answers
    m, n, k = 150, 130, 40
    a, lda, b, ldb, c0, ldc = problem(rng, oracle, dtype, 0, 1, m, n, k, pad=(0, 0, 0))
    alpha, beta = alpha_beta(dtype)[0][2], alpha_beta(dtype)[1][2]
    want = c0.copy()
    oracle.gemm(dtype, 0, 1, m, n, k, alpha, a, lda, b, ldb, beta, want, ldc)
    for alloc in (sim.hostsim_pinned_alloc, sim.hostsim_device_alloc):
        bufs = []
        for x in (a, b, c0):
            p = alloc(x.nbytes)
            C.memmove(p, x.ctypes.data, x.nbytes)
            bufs.append(p)
        fgemm(sim, dtype, 0, 1, m, n, k, alpha, bufs[0], lda, bufs[1], ldb, beta, bufs[2], ldc)
        got = np.empty_like(c0)
        C.memmove(got.ctypes.data, bufs[2], got.nbytes)
        assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), alloc
        for p in bufs:
            sim.hostsim_free(p)
    assert es_out in (4, 8, 16)


@pytest.mark.parametrize("one_big", [False, True])
def test_gemm_batch_is_staged_in_one_upload_and_one_download(sim, oracle, one_big):
    rng = np.random.default_rng(9)
    groups = [(0, 1, 12, 9, 20, 3), (1, 0, 40, 33, 8, 2), (0, 0, 130 if one_big else 30, 20, 16, 4), (1, 1, 7, 5, 3, 4)]
    alphas, betas = [0.7, 1.0, -0.4, 2.0], [1.3, 0.0, 1.0, 0.0]
    cb = {0: 111, 1: 112}
    probs = []
    for gi, (ta, tb, m, n, k, cnt) in enumerate(groups):
        for _ in range(cnt):
            a, lda, b, ldb, c0, ldc = problem(rng, oracle, cpu.D, ta, tb, m, n, k, pad=(1, 1, 1))
            got = c0.copy()
            if betas[gi] == 0.0:
                got[:, :m] = np.nan
            probs.append((a, lda, b, ldb, c0, got, ldc))
    ints = lambda v: (C.c_int * len(v))(*v)
    first = [sum(g[5] for g in groups[:i]) for i in range(len(groups))]
    ptrs = lambda j: (C.c_void_p * len(probs))(*[p[j].ctypes.data for p in probs])
    copies, launches = sim.hostsim_copy_count(), sim.b200_launch_count()
    sim.cblas_dgemm_batch(102, ints([cb[g[0]] for g in groups]), ints([cb[g[1]] for g in groups]), ints([g[2] for g in groups]),
                          ints([g[3] for g in groups]), ints([g[4] for g in groups]), (C.c_double * 4)(*alphas), ptrs(0),
                          ints([probs[f][1] for f in first]), ptrs(2), ints([probs[f][3] for f in first]), (C.c_double * 4)(*betas),
                          ptrs(5), ints([probs[f][6] for f in first]), len(groups), ints([g[5] for g in groups]))
    assert sim.hostsim_copy_count() - copies == 2, "a packed batch is one H2D (problem list included) and one D2H"
    if one_big:     # a matrix beyond 128 rows: every matrix gets the kernel the dispatcher picks for it
        assert sim.b200_launch_count() - launches == len(probs)
    else:
        assert sim.b200_launch_count() - launches == 1, "small matrices of one precision: one grouped launch for the whole batch"
    i = 0
    for gi, (ta, tb, m, n, k, cnt) in enumerate(groups):
        for _ in range(cnt):
            a, lda, b, ldb, c0, got, ldc = probs[i]
            i += 1
            want = c0.copy()
            oracle.gemm(cpu.D, ta, tb, m, n, k, alphas[gi], a, lda, b, ldb, betas[gi], want, ldc)
            assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), (gi, m, n, k)


def test_level3_family_golden_vectors_through_the_host_path(sim, oracle):
    """The 144 + 192 reference-generated cases of SYMM/HEMM/SYRK/HERK/SYR2K/HER2K and TRMM/TRSM: operand
    expansion, block columns, triangle merge, recursion order and offsets, with the oracle as GEMM."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "level3_golden.npz"))
    call = L.bind(sim)
    for idx, row in enumerate(g["meta"]):
        case = L.meta_case(row)
        a, b, c0, ref_c = g[f"a{idx}"], g[f"b{idx}"], g[f"c0_{idx}"], g[f"c{idx}"]
        got, _, gauge, K, touched = L.run_case(call, oracle, case, a, b if b.size else a, c0)
        L.check_case(case, got, ref_c, gauge, K, touched, c0)
    g = np.load(os.path.join(ROOT, "tests", "golden", "trxm_golden.npz"))
    for idx, row in enumerate(g["meta"]):
        L.check_trxm(oracle, sim, L.trxm_meta_case(row), g[f"a{idx}"], g[f"b0_{idx}"])


@pytest.mark.parametrize("rankk_tri", ["1", "0"])
@pytest.mark.parametrize("dtype", [cpu.D, cpu.CX])
def test_rank_k_schemes_and_symm_at_block_crossing_sizes(sim, oracle, dtype, rankk_tri, monkeypatch):
    """n = 300 crosses the 128-wide block columns of the fallback scheme; B200_RANKK_TRI selects it or
    the masked-triangle GEMM.  NaN fills whatever must not be read or written."""
    monkeypatch.setenv("B200_RANKK_TRI", rankk_tri)
    call = L.bind(sim)
    rng = np.random.default_rng(300 + dtype)
    cplx = dtype in (cpu.CX, cpu.Z)
    for herm in ((0, 1) if cplx else (0,)):
        for x in (0, 1):
            for uplo in (0, 1):
                m, n = 90, 70
                ka = n if x else m
                a, b, c0 = L.operand(rng, dtype, ka, ka + 1), L.operand(rng, dtype, n, m + 2), L.operand(rng, dtype, n, m + 3)
                jj, ii = np.meshgrid(np.arange(ka), np.arange(ka + 1), indexing="ij")
                a[(ii < jj) if uplo else ((ii > jj) & (ii < ka))] = np.nan
                alpha, beta = ((0.7 - 0.9j, 1.3 - 1.1j) if cplx else (0.7, 1.3))
                case = (0, dtype, herm, x, uplo, 0, m, n, 0, ka + 1, m + 2, m + 3, alpha, beta)
                before = sim.b200_launch_count()
                got, want, gauge, K, touched = L.run_case(call, oracle, case, a, b, c0)
                L.check_case(case, got, want, gauge, K, touched, c0)
                # the symmetric operand is bigger than the (host-simulation) full-expansion limit: expanded and multiplied panel by panel
                assert sim.b200_launch_count() - before >= 6, sim.b200_launch_count() - before
                for trans in (0, 1):
                    for (nn, k, beta_zero) in [(300, 24, False), (140, 33, True)]:
                        rows, cols = (k, nn) if trans else (nn, k)
                        a, b = L.operand(rng, dtype, cols, rows + 2), L.operand(rng, dtype, cols, rows + 4)
                        c0 = L.operand(rng, dtype, nn, nn + 3)
                        jj, ii = np.meshgrid(np.arange(nn), np.arange(nn + 3), indexing="ij")
                        if beta_zero:
                            c0[:, :nn] = np.nan
                        else:
                            c0[((ii < jj) if uplo else (ii > jj)) & (ii < nn)] = np.nan
                        al = 0.7 if (herm and not x) or not cplx else 0.7 - 0.9j
                        be = 0.0 if beta_zero else (1.3 if herm or not cplx else 1.3 - 1.1j)
                        case = (1, dtype, herm, x, uplo, trans, nn, nn, k, rows + 2, rows + 4, nn + 3, al, be)
                        got, want, gauge, K, touched = L.run_case(call, oracle, case, a, b, c0)
                        L.check_case(case, got, want, gauge, K, touched, c0)
                        if nn == 300:
                            assert (sim.b200_last_kernel() == b"sim_tri_merge") == (rankk_tri == "0"), sim.b200_last_kernel()


@pytest.mark.parametrize("dtype", [cpu.D, cpu.Z])
def test_trxm_recursion_offsets(sim, oracle, dtype):
    """Triangles of 150 and 203 rows (three / four levels of the recursive split, ragged last block),
    every side / uplo / trans / diag combination, through the real recursion with simulated kernels."""
    rng = np.random.default_rng(900 + dtype)
    cplx = dtype in (cpu.CX, cpu.Z)
    alpha = (0.7 - 0.9j) if cplx else 0.7
    for solve in (0, 1):
        for side in (0, 1):
            for uplo in (0, 1):
                for trans in range(4 if cplx else 2):
                    for unit in (0, 1):
                        m, n = (150, 20) if (trans + unit) % 2 == 0 else (17, 203)
                        ka = n if side else m
                        a = L.tri_operand(rng, dtype, ka, ka + 1, uplo, unit)
                        if solve:
                            off = ~np.eye(ka, ka + 1, dtype=bool)
                            a[off] *= 4.0 / ka
                        b0 = L.operand(rng, dtype, n, m + 2)
                        L.check_trxm(oracle, sim, (dtype, solve, side, uplo, trans, unit, m, n, ka + 1, m + 2, alpha), a, b0)


def test_row_major_cblas_entry_points(sim):
    """Row-major normalisation of every family against numpy on the logical matrices."""
    rng = np.random.default_rng(5)
    m, n, k = 23, 17, 29
    A, B, Cm = rng.random((m, k)) - 0.5, rng.random((k, n)) - 0.5, rng.random((m, n)) - 0.5
    got = Cm.copy()
    P = lambda x: x.ctypes.data_as(C.c_void_p)
    sim.cblas_dgemm(101, 111, 111, m, n, k, C.c_double(0.7), P(A), k, P(B), n, C.c_double(1.3), P(got), n)
    assert np.allclose(got, 0.7 * A @ B + 1.3 * Cm, rtol=0, atol=1e-13)
    At = np.ascontiguousarray(A.T)      # k x m row-major, used transposed
    got = Cm.copy()
    sim.cblas_dgemm(101, 112, 111, m, n, k, C.c_double(0.7), P(At), m, P(B), n, C.c_double(0.0), P(got), n)
    assert np.allclose(got, 0.7 * A @ B, rtol=0, atol=1e-13)
    # ZHER2K row-major upper, NoTrans: conj(alpha) handling of interface/syr2k.c:305-311
    nn, kk = 21, 13
    Az = rng.random((nn, kk)) - 0.5 + 1j * (rng.random((nn, kk)) - 0.5)
    Bz = rng.random((nn, kk)) - 0.5 + 1j * (rng.random((nn, kk)) - 0.5)
    C0 = rng.random((nn, nn)) - 0.5 + 1j * (rng.random((nn, nn)) - 0.5)
    alpha = 0.7 - 0.9j
    got = C0.copy()
    al = np.array([alpha.real, alpha.imag])
    sim.cblas_zher2k(101, 121, 111, nn, kk, P(al), P(Az), kk, P(Bz), kk, C.c_double(1.3), P(got), nn)
    full = alpha * Az @ Bz.conj().T + np.conj(alpha) * Bz @ Az.conj().T + 1.3 * C0
    up = np.triu(np.ones((nn, nn), dtype=bool))
    assert np.allclose(got[up], np.where(np.eye(nn, dtype=bool), full.real + 0j, full)[up], rtol=0, atol=1e-13)
    assert np.array_equal(got[~up], C0[~up])
    # DTRSM row-major, Right, Lower, Trans, NonUnit: X * L^T = alpha * B
    Lm = np.tril(rng.random((n, n)) - 0.5) * 0.2 + np.eye(n)
    Bm = rng.random((m, n)) - 0.5
    X = Bm.copy()
    sim.cblas_dtrsm(101, 142, 122, 112, 131, m, n, C.c_double(0.7), P(Lm), n, P(X), n)
    assert np.allclose(X @ Lm.T, 0.7 * Bm, rtol=0, atol=1e-12)
    # DSYMM row-major, Left, Upper
    S = rng.random((m, m)) - 0.5
    S = np.triu(S) + np.triu(S, 1).T
    Su = np.triu(S) + np.tril(np.full((m, m), np.nan), -1)
    Bs = rng.random((m, n)) - 0.5
    got = Cm.copy()
    sim.cblas_dsymm(101, 141, 121, m, n, C.c_double(0.7), P(Su), m, P(Bs), n, C.c_double(1.3), P(got), n)
    assert np.allclose(got, 0.7 * S @ Bs + 1.3 * Cm, rtol=0, atol=1e-13)


def test_bf16_conversions_through_the_host_path(sim):
    g = np.load(os.path.join(ROOT, "tests", "golden", "bf16_golden.npz"))
    x = g["x"]
    h = np.zeros(x.size, dtype=np.uint16)
    sim.cblas_sbstobf16(x.size, x.ctypes.data_as(C.c_void_p), 1, h.ctypes.data_as(C.c_void_p), 1)
    assert np.array_equal(h, g["bf16"])
    h2 = np.full(2 * 50, 0xdead, dtype=np.uint16)
    sim.cblas_sbstobf16(50, x.ctypes.data_as(C.c_void_p), 1, h2.ctypes.data_as(C.c_void_p), -2)
    assert np.array_equal(h2[::2][::-1], g["bf16"][:50]) and np.all(h2[1::2] == 0xdead)
    # one element per call (the in-place pinned path of run_convert_on_context), all four directions
    for i in range(0, x.size, max(1, x.size // 64)):
        one, one_back, d_in, d_out, d_back = np.zeros(1, np.uint16), np.zeros(1, np.float32), x[i:i + 1].astype(np.float64), np.zeros(1, np.uint16), np.zeros(1, np.float64)
        sim.cblas_sbstobf16(1, C.c_void_p(x[i:i + 1].ctypes.data), 1, C.c_void_p(one.ctypes.data), 1)
        sim.cblas_sbf16tos(1, C.c_void_p(one.ctypes.data), 1, C.c_void_p(one_back.ctypes.data), 1)
        sim.cblas_sbdtobf16(1, C.c_void_p(d_in.ctypes.data), 1, C.c_void_p(d_out.ctypes.data), 1)
        sim.cblas_dbf16tod(1, C.c_void_p(d_out.ctypes.data), 1, C.c_void_p(d_back.ctypes.data), 1)
        assert one[0] == g["bf16"][i] and one_back.view(np.uint32)[0] == g["back"].view(np.uint32)[i]
        assert d_out[0] == g["bf16"][i] and np.float32(d_back[0]).view(np.uint32) == g["back"].view(np.uint32)[i]


def _build_and_run_thread_test(sanitize):
    env = dict(os.environ)
    if sanitize:
        env["HOSTSIM_SANITIZE"] = sanitize
    libpath = subprocess.check_output([os.path.join(ROOT, "tests", "hostsim", "build.sh")], text=True, env=env).strip().splitlines()[-1]
    libdir = os.path.dirname(libpath)
    exe = os.path.join(libdir, "thread_test")
    flags = [f"-fsanitize={sanitize}"] if sanitize else []
    subprocess.check_call(["g++"] + flags + ["-std=c++17", "-O1", "-g", f"-I{ROOT}/include", os.path.join(ROOT, "tests", "hostsim", "thread_test.cpp"),
                           "-o", exe, f"-L{libdir}", "-lopenblas_b200_hostsim", f"-Wl,-rpath,{libdir}", "-lpthread"])
    return subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600,
                          env=dict(env, TSAN_OPTIONS="halt_on_error=0"))


def test_concurrent_callers_identical_and_race_free():
    """cpp_thread_test/dgemm_thread_safety.cpp on the host path: 8 threads, GEMM / SYRK / TRSM /
    GEMM_BATCH over the packed, strided and pipelined staging paths, byte-identical results, and no
    data race reported by ThreadSanitizer in the context pool, lazy initialisation, launch counter or
    pinned-slot pool (falls back to an uninstrumented run where TSan cannot start)."""
    r = _build_and_run_thread_test("thread")
    if r.returncode != 0 and ("unexpected memory mapping" in r.stdout or "THREADS" not in r.stdout):
        r = _build_and_run_thread_test("")
    assert r.returncode == 0, r.stdout[-3000:]
    assert "THREADS OK" in r.stdout, r.stdout[-3000:]
    assert "WARNING: ThreadSanitizer" not in r.stdout, r.stdout[-6000:]


def test_extension_api_and_kernel_forcing(sim, oracle):
    """Part 2 of the header: b200_gemm on host / pinned / device pointers, b200_gemm_async on device
    pointers, b200_set_kernel: GENERIC always works, FAST reports cudaErrorNotSupported (801) for a
    problem the roofline kernels do not take instead of silently running something else."""
    sim.b200_gemm.restype = C.c_int
    sim.b200_gemm_async.restype = C.c_int
    sim.b200_host_alloc.restype = C.c_void_p
    sim.b200_host_alloc.argtypes = [C.c_size_t]
    sim.b200_host_free.argtypes = [C.c_void_p]
    sim.b200_last_error.restype = C.c_char_p
    rng = np.random.default_rng(2)
    m, n, k = 96, 80, 64
    a, lda, b, ldb, c0, ldc = problem(rng, oracle, cpu.D, 0, 0, m, n, k, pad=(0, 0, 0))
    want = c0.copy()
    oracle.gemm(cpu.D, 0, 0, m, n, k, 0.7, a, lda, b, ldb, 1.3, want, ldc)
    al, be = np.array([0.7]), np.array([1.3])
    i64 = C.c_int64
    P = lambda x: C.c_void_p(x) if isinstance(x, int) else x.ctypes.data_as(C.c_void_p)
    args = lambda pa, pb, pc: (1, 0, 0, i64(m), i64(n), i64(k), P(al), P(pa), i64(lda), P(pb), i64(ldb), P(be), P(pc), i64(ldc))
    got = c0.copy()
    assert sim.b200_gemm(*args(a, b, got)) == 0
    assert np.array_equal(got, want)      # (host operands of this size take the panel pipeline: 32 x 32 blocks -> generic kernel)
    # pinned memory from the library's own allocator, then device memory through the async entry point
    for alloc, free, use_async in ((sim.b200_host_alloc, sim.b200_host_free, False), (sim.hostsim_device_alloc, sim.hostsim_free, True)):
        bufs = []
        for x in (a, b, c0):
            p = alloc(x.nbytes)
            C.memmove(p, x.ctypes.data, x.nbytes)
            bufs.append(p)
        if use_async:
            assert sim.b200_gemm_async(*args(*bufs), None) == 0
        else:
            assert sim.b200_gemm(*args(*bufs)) == 0
        got = np.empty_like(c0)
        C.memmove(got.ctypes.data, bufs[2], got.nbytes)
        assert np.array_equal(got, want)
        if use_async:                   # device operands are used in place: one launch of the roofline kernel's stand-in
            assert sim.b200_last_kernel().startswith(b"sim_dgemm"), sim.b200_last_kernel()
        for p in bufs:
            free(p)
    try:
        sim.b200_set_kernel(1)                       # GENERIC
        got = c0.copy()
        dev = []
        for x in (a, b, c0):
            p = sim.hostsim_device_alloc(x.nbytes)
            C.memmove(p, x.ctypes.data, x.nbytes)
            dev.append(p)
        assert sim.b200_gemm(*args(*dev)) == 0 and sim.b200_last_kernel() == b"sim_generic"
        C.memmove(got.ctypes.data, dev[2], got.nbytes)
        assert np.array_equal(got, want)
        for p in dev:
            sim.hostsim_free(p)
        sim.b200_set_kernel(2)                       # FAST: a ZGEMM the DMMA kernel refuses (device operands only 8-byte aligned)
        assert sim.b200_get_kernel() == 2
        za = (rng.random((k, m)) + 1j * rng.random((k, m))).astype(np.complex128)
        zb = (rng.random((n, k)) + 1j * rng.random((n, k))).astype(np.complex128)
        zc = np.zeros((n, m), dtype=np.complex128)
        dev = []
        for x in (za, zb, zc):
            p = sim.hostsim_device_alloc(x.nbytes + 16)
            C.memmove(p + 8, x.ctypes.data, x.nbytes)
            dev.append(p)
        za1, zb0 = np.array([1.0, 0.0]), np.array([0.0, 0.0])
        rc = sim.b200_gemm(3, 0, 0, i64(m), i64(n), i64(k), P(za1), C.c_void_p(dev[0] + 8), i64(m), C.c_void_p(dev[1] + 8), i64(k), P(zb0),
                           C.c_void_p(dev[2] + 8), i64(m))
        assert rc == 801, rc
        for p in dev:
            sim.hostsim_free(p)
        # an SBGEMM below the old tcgen05 minimums (m < 128) is taken by the fast path now
        ha, hb = oracle.tobf16(rng.random((k, m), dtype=np.float32)), oracle.tobf16(rng.random((n, k), dtype=np.float32))
        hc = np.zeros((n, m), dtype=np.float32)
        fa, fb = np.array([1.0], dtype=np.float32), np.array([0.0], dtype=np.float32)
        rc = sim.b200_gemm(4, 0, 0, i64(m), i64(n), i64(k), P(fa), P(ha), i64(m), P(hb), i64(k), P(fb), P(hc), i64(m))
        assert rc == 0 and sim.b200_last_kernel() == b"sim_sbgemm", (rc, sim.b200_last_kernel())
    finally:
        sim.b200_set_kernel(0)


def test_fork_after_first_use_fails_loudly(tmp_path):
    """DESIGN 'known deviations': the CUDA context does not survive fork(); a GEMM in a child forked
    after the parent's first call must abort with an explanation, not compute on a dead context."""
    lib = subprocess.check_output([os.path.join(ROOT, "tests", "hostsim", "build.sh")], text=True).strip().splitlines()[-1]
    script = tmp_path / "fork_child.py"
    script.write_text(f'''
import ctypes as C, os, sys
import numpy as np
lib = C.CDLL({lib!r})
a = np.ones((64, 64)); c = np.zeros((64, 64))
def call():
    lib.cblas_dgemm(102, 111, 111, 64, 64, 64, C.c_double(1.0), a.ctypes.data_as(C.c_void_p), 64, a.ctypes.data_as(C.c_void_p), 64,
                    C.c_double(0.0), c.ctypes.data_as(C.c_void_p), 64)
call()
assert c[0, 0] == 64.0
pid = os.fork()
if pid == 0:
    call()
    os._exit(0)
_, status = os.waitpid(pid, 0)
print("child signal", os.WTERMSIG(status) if os.WIFSIGNALED(status) else 0, "exit", os.WEXITSTATUS(status) if os.WIFEXITED(status) else -1)
''')
    r = subprocess.run(["python", str(script)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert "child signal 6" in r.stdout, r.stdout            # SIGABRT from b200_fatal
    assert "cannot be used after fork()" in r.stdout, r.stdout


@pytest.mark.parametrize("m", [48, 136])
def test_gemm_batch_on_device_operands_moves_no_matrix(sim, oracle, m):
    """Device-resident batch: small matrices -> the problem list goes up (one small copy) and ONE grouped
    launch runs; beyond 128 rows -> no copy at all, one launch per matrix."""
    rng = np.random.default_rng(21)
    n, k, cnt = 40, 36, 5
    hosts, devs = [], []
    for _ in range(cnt):
        a, lda, b, ldb, c0, ldc = problem(rng, oracle, cpu.D, 0, 1, m, n, k, pad=(2, 2, 2))
        ptrs = []
        for x in (a, b, c0):
            p = sim.hostsim_device_alloc(x.nbytes)
            C.memmove(p, x.ctypes.data, x.nbytes)
            ptrs.append(p)
        hosts.append((a, lda, b, ldb, c0, ldc))
        devs.append(ptrs)
    ints = lambda v: (C.c_int * len(v))(*v)
    vp = lambda j: (C.c_void_p * cnt)(*[d[j] for d in devs])
    copies, launches = sim.hostsim_copy_count(), sim.b200_launch_count()
    sim.cblas_dgemm_batch(102, ints([111]), ints([112]), ints([m]), ints([n]), ints([k]), (C.c_double * 1)(0.7), vp(0), ints([hosts[0][1]]),
                          vp(1), ints([hosts[0][3]]), (C.c_double * 1)(1.3), vp(2), ints([hosts[0][5]]), 1, ints([cnt]))
    if m <= 128:
        assert sim.hostsim_copy_count() - copies == 1 and sim.b200_launch_count() - launches == 1
    else:
        assert sim.hostsim_copy_count() == copies and sim.b200_launch_count() - launches == cnt
    for (a, lda, b, ldb, c0, ldc), d in zip(hosts, devs):
        want = c0.copy()
        oracle.gemm(cpu.D, 0, 1, m, n, k, 0.7, a, lda, b, ldb, 1.3, want, ldc)
        got = np.empty_like(c0)
        C.memmove(got.ctypes.data, d[2], got.nbytes)
        assert np.array_equal(got.view(np.uint8), want.view(np.uint8))
        for p in d:
            sim.hostsim_free(p)


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_randomised_gemm_staging_stress(sim, oracle, seed):
    """Random GEMMs around every threshold of the host path: extents from sets that straddle the packed /
    strided / pipelined limits and the panel edges (0 included), random leading-dimension padding,
    alpha and beta from {0, 1, other}, every operand independently pageable, pinned or device memory,
    all five precisions.  Result = one oracle call, bit for bit; padding rows keep their bytes."""
    rng = np.random.default_rng(1000 + seed)
    dims = [0, 1, 2, 31, 32, 33, 63, 64, 65, 96, 127, 128, 129, 200, 257]
    ks = [0, 1, 5, 33, 64]
    allocs = [None, sim.hostsim_pinned_alloc, sim.hostsim_device_alloc]
    for it in range(60):
        dtype = int(rng.integers(0, 5))
        m, n, k = int(rng.choice(dims)), int(rng.choice(dims)), int(rng.choice(ks))
        if m * n * max(k, 1) > 3_000_000:
            k = min(k, 5)
        ta, tb = int(rng.integers(0, ntrans(dtype))), int(rng.integers(0, ntrans(dtype)))
        pad = tuple(int(x) for x in rng.integers(0, 4, size=3))
        a, lda, b, ldb, c0, ldc = problem(rng, oracle, dtype, ta, tb, m, n, k, pad=pad)
        alphas, betas = alpha_beta(dtype)
        alpha, beta = alphas[int(rng.integers(0, 3))], betas[int(rng.integers(0, 3))]
        if complex(beta) == 0:
            c0[:, :m] = np.nan
        want = c0.copy()
        oracle.gemm(dtype, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, want, ldc)
        kinds = [allocs[int(rng.integers(0, 3))] for _ in range(3)]
        ptrs, owned = [], []
        for x, alloc in zip((a, b, c0.copy()), kinds):
            if alloc is None:
                ptrs.append(x)
            else:
                p = alloc(max(x.nbytes, 1))
                C.memmove(p, x.ctypes.data, x.nbytes)
                ptrs.append(p)
                owned.append(p)
        fgemm(sim, dtype, ta, tb, m, n, k, alpha, ptrs[0], lda, ptrs[1], ldb, beta, ptrs[2], ldc)
        if isinstance(ptrs[2], int):
            got = np.empty_like(c0)
            C.memmove(got.ctypes.data, ptrs[2], got.nbytes)
        else:
            got = ptrs[2]
        ctx = (seed, it, NAMES[dtype], ta, tb, m, n, k, pad, alpha, beta, [k_ is not None and k_ is sim.hostsim_device_alloc for k_ in kinds])
        assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), ctx
        for p in owned:
            sim.hostsim_free(p)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_randomised_level3_stress(sim, oracle, seed, monkeypatch):
    """Random SYMM/HEMM, SYRK/HERK, SYR2K/HER2K, TRMM/TRSM calls: extents around the 64-wide recursion
    blocks and the 128-wide block columns (1 included), alpha / beta from {0, 1, other}, both rank-k
    schemes, against the oracle within the level-3 bounds, untouched parts bit-identical."""
    rng = np.random.default_rng(2000 + seed)
    call = L.bind(sim)
    dims = [1, 2, 31, 63, 64, 65, 127, 128, 129, 200]
    for it in range(40):
        monkeypatch.setenv("B200_RANKK_TRI", str(int(rng.integers(0, 2))))
        dtype = int(rng.integers(0, 4))
        cplx = dtype in (cpu.CX, cpu.Z)
        fam = int(rng.integers(0, 3))
        m, n, k = int(rng.choice(dims)), int(rng.choice(dims)), int(rng.choice([0, 1, 7, 40]))
        scal = [0.0, 1.0, (0.7 - 0.9j) if cplx else 0.7]
        alpha, beta = scal[int(rng.integers(0, 3))], scal[int(rng.integers(0, 3))]
        uplo = int(rng.integers(0, 2))
        if fam == 0:
            herm, side = (int(rng.integers(0, 2)) if cplx else 0), int(rng.integers(0, 2))
            ka = n if side else m
            pa, pb, pc = (int(x) for x in rng.integers(0, 3, size=3))
            a, b, c0 = L.operand(rng, dtype, ka, ka + pa), L.operand(rng, dtype, n, m + pb), L.operand(rng, dtype, n, m + pc)
            jj, ii = np.meshgrid(np.arange(ka), np.arange(ka + pa), indexing="ij")
            a[(ii < jj) if uplo else ((ii > jj) & (ii < ka))] = np.nan
            if complex(beta) == 0:
                c0[:, :m] = np.nan
            case = (0, dtype, herm, side, uplo, 0, m, n, 0, ka + pa, m + pb, m + pc, alpha, beta)
            got, want, gauge, K, touched = L.run_case(call, oracle, case, a, b, c0)
            L.check_case(case, got, want, gauge, K, touched, c0)
        elif fam == 1:
            herm, two, trans = (int(rng.integers(0, 2)) if cplx else 0), int(rng.integers(0, 2)), int(rng.integers(0, 2))
            rows, cols = (k, n) if trans else (n, k)
            pa, pb, pc = (int(x) for x in rng.integers(0, 3, size=3))
            a, b = L.operand(rng, dtype, max(cols, 1), max(rows, 1) + pa), L.operand(rng, dtype, max(cols, 1), max(rows, 1) + pb)
            c0 = L.operand(rng, dtype, n, n + pc)
            al = complex(alpha).real if (herm and not two) else alpha
            be = complex(beta).real if herm else beta
            jj, ii = np.meshgrid(np.arange(n), np.arange(n + pc), indexing="ij")
            if complex(be) == 0:
                c0[:, :n] = np.nan
            else:
                c0[((ii < jj) if uplo else (ii > jj)) & (ii < n)] = np.nan
            case = (1, dtype, herm, two, uplo, trans, n, n, k, max(rows, 1) + pa, max(rows, 1) + pb, n + pc, al, be)
            got, want, gauge, K, touched = L.run_case(call, oracle, case, a, b, c0)
            if complex(be) == 1 and (k == 0 or complex(al) == 0):
                assert np.array_equal(got.view(np.uint8), c0.view(np.uint8))
            else:
                L.check_case(case, got, want, gauge, K, touched, c0)
        else:
            solve, side, trans, unit = int(rng.integers(0, 2)), int(rng.integers(0, 2)), int(rng.integers(0, 4 if cplx else 2)), int(rng.integers(0, 2))
            ka = n if side else m
            a = L.tri_operand(rng, dtype, ka, ka + 1, uplo, unit)
            if solve and ka > 1:
                off = ~np.eye(ka, ka + 1, dtype=bool)
                a[off] *= min(1.0, 4.0 / ka)
            b0 = L.operand(rng, dtype, n, m + 2)
            L.check_trxm(oracle, sim, (dtype, solve, side, uplo, trans, unit, m, n, ka + 1, m + 2, alpha), a, b0)


def test_gemmt_sbgemv_sbdot_golden_vectors_through_the_host_path(sim, oracle):
    """The reference-generated cases of tests/golden/f_rows_golden.npz through interface_level3.c / interface_gemm.c
    and the staging of runtime_level3.inl / runtime_bf16.inl: flag decoding, the row-major swap, operand
    shapes and offsets, increments (gathered / scattered on the host), alpha == 0 and beta == 0 paths."""
    import test_f_rows_pin as F
    g = np.load(os.path.join(ROOT, "tests", "golden", "f_rows_golden.npz"))
    for i in range(int(g["gemmt_count"][0])):
        key = f"gemmt{i}"
        dtype, uplo, ta, tb, m, k, lda, ldb, ldc, cblas, rowmajor = F.colmajor_problem(g[key + "_meta"])
        cplx = dtype in (cpu.CX, cpu.Z)
        alpha, beta = ((0.7 - 0.9j, 1.3 - 1.1j) if cplx else (0.7, 1.3))
        a, b, c0 = g[key + "_a"], g[key + "_b"], g[key + "_c0"]
        got = c0.copy()
        cpu.call_gemmt(sim, dtype, uplo, ta, tb, m, k, alpha, a, lda, b, ldb, beta, got, ldc, cblas=bool(cblas), rowmajor=bool(rowmajor))
        assert np.array_equal(a, g[key + "_a"]) and np.array_equal(b, g[key + "_b"])         # inputs are const here
        want, gauge, mask = F.gemmt_expected(oracle, dtype, uplo, ta, tb, m, k, alpha, a, lda, b, ldb, beta, c0, ldc, rowmajor)
        F.check_gemmt(dtype, m, k, ldc, got, want, gauge, mask, c0, key)                    # against the restatement
        F.check_gemmt(dtype, m, k, ldc, got, g[key + "_c"], gauge, mask, c0, key + " vs reference")
    for i in range(int(g["sbgemv_count"][0])):
        key = f"sbgemv{i}"
        trans, m, n, lda, incx, incy = (int(v) for v in g[key + "_meta"])
        alpha, beta = (float(v) for v in g[key + "_ab"])
        for cblas in (False, True):
            y = g[key + "_y0"].copy()
            cpu.call_sbgemv(sim, trans, m, n, alpha, g[key + "_a"], lda, g[key + "_x"], incx, beta, y, incy, cblas=cblas)
            assert np.array_equal(y.view(np.uint32), g[key + "_y"].view(np.uint32)), (key, cblas)   # stand-in kernel = oracle: bit for bit
    for i in range(int(g["sbdot_count"][0])):
        key = f"sbdot{i}"
        n, incx, incy = (int(v) for v in g[key + "_meta"])
        for cblas in (False, True):
            assert np.float32(cpu.call_sbdot(sim, n, g[key + "_x"], incx, g[key + "_y"], incy, cblas=cblas)) == g[key + "_d"][0], (key, cblas)


def test_gemmt_block_column_scheme_and_device_operands(sim, oracle, monkeypatch):
    """GEMMT through the block-column fallback (B200_RANKK_TRI=0) at a size that crosses block boundaries, host and
    "device" operands, NaN outside the triangle."""
    import test_f_rows_pin as F
    rng = np.random.default_rng(12)
    for tri_env in ("1", "0"):
        monkeypatch.setenv("B200_RANKK_TRI", tri_env)
        for dtype in (cpu.D, cpu.CX):
            cplx = dtype == cpu.CX
            for uplo in (0, 1):
                for ta, tb in ((0, 0), (1, 0), (0, 1), (1, 1)) + (((3, 2), (2, 3)) if cplx else ()):
                    m, k = 150, 37
                    ra, ca = (k, m) if ta & 1 else (m, k)
                    rb, cb = (m, k) if tb & 1 else (k, m)
                    lda, ldb, ldc = ra + 3, rb + 1, m + 2
                    a, b, c0 = F.operand(rng, dtype, ca, lda), F.operand(rng, dtype, cb, ldb), F.operand(rng, dtype, m, ldc)
                    mask = F.tri_mask(m, ldc, uplo)
                    c0[~mask] = np.nan
                    alpha, beta = ((0.7 - 0.9j, 1.3 - 1.1j) if cplx else (0.7, 1.3))
                    got = c0.copy()
                    cpu.call_gemmt(sim, dtype, uplo, ta, tb, m, k, alpha, a, lda, b, ldb, beta, got, ldc)
                    want, gauge, mk = F.gemmt_expected(oracle, dtype, uplo, ta, tb, m, k, alpha, a, lda, b, ldb, beta, c0, ldc, False)
                    F.check_gemmt(dtype, m, k, ldc, got, want, gauge, mk, c0, (tri_env, dtype, uplo, ta, tb))


def test_sbgemv_sbdot_on_device_operands_use_the_callers_increments(sim, oracle):
    """"device" operands (hostsim allocations) are used in place: no staging copies, the kernel stand-in sees the
    caller's increments (negative ones walk from the far end); mixed host / device operands stage only the host ones."""
    rng = np.random.default_rng(21)
    m, n, lda = 37, 23, 40
    a = oracle.tobf16(rng.random((n, lda), dtype=np.float32) - 0.5)
    for trans in (0, 1):
        lenx, leny = (m, n) if trans else (n, m)
        for incx, incy in ((1, 1), (-2, 3), (3, -1)):
            x = oracle.tobf16(rng.random(1 + (lenx - 1) * abs(incx), dtype=np.float32) - 0.5)
            y0 = (rng.random(1 + (leny - 1) * abs(incy)) - 0.5).astype(np.float32)
            want = y0.copy()
            oracle.sbgemv(trans, m, n, 0.7, a, lda, x, incx, 1.3, want, incy)
            for mix in ("all", "a only", "y only"):
                da = sim.hostsim_device_alloc(a.nbytes); C.memmove(da, a.ctypes.data, a.nbytes)
                dx = sim.hostsim_device_alloc(x.nbytes); C.memmove(dx, x.ctypes.data, x.nbytes)
                dy = sim.hostsim_device_alloc(y0.nbytes); C.memmove(dy, y0.ctypes.data, y0.nbytes)
                hy = y0.copy()
                pa = da if mix in ("all", "a only") else a
                px = dx if mix == "all" else x
                py = dy if mix in ("all", "y only") else hy
                before = sim.hostsim_copy_count()
                cpu.call_sbgemv(sim, trans, m, n, 0.7, pa, lda, px, incx, 1.3, py, incy, cblas=(mix == "all"))
                copies = sim.hostsim_copy_count() - before
                got = np.empty_like(y0)
                if mix in ("all", "y only"):
                    C.memmove(got.ctypes.data, dy, got.nbytes)
                else:
                    got = hy
                assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (trans, incx, incy, mix)
                if mix == "all":
                    assert copies == 0, copies
                for p in (da, dx, dy):
                    sim.hostsim_free(p)
    x = oracle.tobf16(rng.random(50, dtype=np.float32) - 0.5)
    dx = sim.hostsim_device_alloc(x.nbytes); C.memmove(dx, x.ctypes.data, x.nbytes)
    assert np.float32(cpu.call_sbdot(sim, 25, dx, 2, x, -1)) == np.float32(oracle.sbdot(25, x, 2, x, -1)[0])
    sim.hostsim_free(dx)


# ---------------------------------------------------------------- the SUMMA driver's host logic on a 1 x 1 grid
@pytest.mark.parametrize("dtype", [cpu.D, cpu.S, cpu.Z, cpu.CX])
def test_summa_c_driver_on_a_1x1_grid_over_the_stand_in(sim, oracle, dtype, monkeypatch):
    """csrc/summa.cu compiled into the stand-in build: window packing (A dense, B panel-major), the k-panel schedule with
    ragged last panels, host and "device" operands, a host C swept in one pass or in two column halves, beta == 0 over a
    NaN C, k == 0, empty local pieces -- against the oracle (each panel is one simulated launch with beta = 1 after the
    first, so the comparison is by the k * eps bound).  Peers, flags and IPC are the multi-GPU runs' business."""
    # read once by the stand-in build's driver, at its first call: a local C of >= 64 columns is swept in two halves
    # (restored afterwards, so the real library in the same process keeps its default)
    monkeypatch.setenv("B200_SUMMA_HOST_HALVES", "64")
    L = sim
    L.b200_summa_create.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.b200_summa_destroy.argtypes = [C.c_void_p]
    L.b200_summa_gemm.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                  C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    L.b200_summa_launches.restype = C.c_uint64
    L.b200_summa_launches.argtypes = [C.c_void_p]
    L.b200_last_error.restype = C.c_char_p
    h = C.c_void_p()
    assert L.b200_summa_create(C.byref(h), None, 0, 1, 1, 1) == 0, L.b200_last_error()
    rng = np.random.default_rng(600 + dtype)
    cplx = dtype in (cpu.CX, cpu.Z)
    t = cpu.NP_OUT[dtype]
    try:
        for (m, n, k, nb) in ((130, 90, 70, 32), (40, 200, 33, 16), (64, 64, 64, 64), (50, 40, 0, 8), (33, 17, 5, 64), (0, 9, 4, 4), (7, 0, 4, 4)):
            for where in ("host", "device"):
                for alpha, beta in (((0.7 - 0.9j, 1.3 - 1.1j) if cplx else (0.7, 1.3)), (1.0, 0.0)):
                    lda, ldb, ldc = max(m, 1) + 3, max(k, 1) + 2, max(m, 1) + 1
                    a, b, c0 = L_operand(rng, t, cplx, max(k, 1), lda), L_operand(rng, t, cplx, max(n, 1), ldb), L_operand(rng, t, cplx, max(n, 1), ldc)
                    want = c0.copy()
                    if m > 0 and n > 0:
                        oracle.gemm(dtype, 0, 0, m, n, k, alpha, a, lda, b, ldb, beta, want, ldc)
                    start = c0.copy()
                    if beta == 0.0 and m > 0 and n > 0:
                        start[:n, :m] = np.nan                       # beta == 0 never reads C
                    al, be = cpu.scalar_bytes(dtype, alpha), cpu.scalar_bytes(dtype, beta)
                    before = L.b200_summa_launches(h)
                    if where == "host":
                        got = start.copy()
                        rc = L.b200_summa_gemm(h, dtype, m, n, k, nb, al.ctypes.data, a.ctypes.data, lda, b.ctypes.data, ldb, be.ctypes.data, got.ctypes.data, ldc, None)
                    else:
                        bufs = []
                        for arr in (a, b, start):
                            p = sim.hostsim_device_alloc(arr.nbytes)
                            C.memmove(p, arr.ctypes.data, arr.nbytes)
                            bufs.append(p)
                        rc = L.b200_summa_gemm(h, dtype, m, n, k, nb, al.ctypes.data, bufs[0], lda, bufs[1], ldb, be.ctypes.data, bufs[2], ldc, None)
                        got = np.empty_like(start)
                        C.memmove(got.ctypes.data, bufs[2], start.nbytes)
                        for p in bufs:
                            sim.hostsim_free(p)
                    assert rc == 0, L.b200_last_error()
                    if m > 0 and n > 0:
                        steps = (k + nb - 1) // nb if k > 0 else 1
                        sweeps = 2 if (where == "host" and n >= 64 and steps >= 2) else 1
                        assert L.b200_summa_launches(h) - before == steps * sweeps, (m, n, k, nb, where)
                        ratio = oracle.ratio(dtype, 0, 0, m, n, k, alpha, a, lda, b, ldb, beta, c0, ldc, got, ldc, want, ldc)
                        assert ratio <= 2.0, (m, n, k, nb, where, alpha, ratio)
                    # rows beyond m (and everything when the local piece is empty) keep their bytes
                    pad = np.ones(start.shape, dtype=bool)
                    if m > 0 and n > 0:
                        pad[:n, :m] = False
                    assert np.array_equal(got.view(np.uint8)[np.repeat(pad, got.itemsize, axis=1)], start.view(np.uint8)[np.repeat(pad, got.itemsize, axis=1)]), \
                        (m, n, k, nb, where, "bytes outside the local piece changed")
    finally:
        assert L.b200_summa_destroy(h) == 0


def L_operand(rng, t, cplx, cols, ld):
    x = rng.random((cols, ld)) - 0.5
    if cplx:
        x = x + 1j * (rng.random((cols, ld)) - 0.5)
    return x.astype(t)


@pytest.mark.parametrize("dtype", [cpu.Z, cpu.CX])
def test_gemm3m_runs_three_real_products(sim, oracle, dtype, monkeypatch):
    """?gemm3m_ above the size limit (8 in this build, 512 in the product): split3 of both operands, three REAL GEMMs,
    combine3 -- for every op pair (conjugation = sign of the imaginary plane, transposition = the real GEMM's own),
    ragged shapes, beta == 0 over a NaN C, host and "device" operands; accepted the way the reference's 3M ctest driver
    accepts it (ctest/c_zblat3c_3m.c: err / (eps * gauge) < 16 with the |re| + |im| gauge).  Below the limit, and through
    the plain ?gemm_ entry points, nothing changes."""
    sim.b200_last_kernel.restype = C.c_char_p
    rng = np.random.default_rng(1200 + dtype)
    name = cpu.DTYPE_NAMES[dtype] + "gemm3m_"
    i = lambda v: C.byref(C.c_int(int(v)))
    P = lambda x: C.c_void_p(x) if isinstance(x, int) else x.ctypes.data_as(C.c_void_p)

    def gemm3m(ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc):
        al, be = cpu.scalar_bytes(dtype, alpha), cpu.scalar_bytes(dtype, beta)
        getattr(sim, name)(C.c_char_p(cpu.TRANS_CHAR[ta].encode()), C.c_char_p(cpu.TRANS_CHAR[tb].encode()), i(m), i(n), i(k), cpu._ptr(al), P(a), i(lda),
                           P(b), i(ldb), cpu._ptr(be), P(c), i(ldc))

    for (m, n, k) in ((40, 33, 29), (9, 64, 8), (17, 8, 50)):
        for ta in range(4):
            for tb in range(4):
                a, lda, b, ldb, c0, ldc = problem(rng, oracle, dtype, ta, tb, m, n, k, pad=(3, 2, 5))
                for alpha, beta in ((0.7 - 0.9j, 1.3 - 1.1j), (1.0, 0.0)):
                    start = c0.copy()
                    if beta == 0.0:
                        start[:, :m] = np.nan
                    got = start.copy()
                    before = sim.b200_launch_count()
                    if (ta + tb) % 2 == 0:
                        gemm3m(ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, got, ldc)
                    else:
                        bufs = []
                        for arr in (a, b, start):
                            p = sim.hostsim_device_alloc(arr.nbytes)
                            C.memmove(p, arr.ctypes.data, arr.nbytes)
                            bufs.append(p)
                        gemm3m(ta, tb, m, n, k, alpha, bufs[0], lda, bufs[1], ldb, beta, bufs[2], ldc)
                        C.memmove(got.ctypes.data, bufs[2], got.nbytes)
                        for p in bufs:
                            sim.hostsim_free(p)
                    assert sim.b200_launch_count() - before == 6 and sim.b200_last_kernel() == b"sim_combine3"
                    assert oracle.mmch(dtype, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c0, ldc, got, ldc) < 16.0, (m, n, k, ta, tb, alpha)
                    assert np.array_equal(got[:, m:].view(np.uint8), start[:, m:].view(np.uint8))      # padding rows keep their bytes
    # a workspace limit that does not hold the whole product: k goes through in chunks (first chunk with the caller's
    # beta, the others accumulate), every op pair; a limit below the C planes alone: the 4-multiply kernel
    m, n, k = 40, 33, 29
    rs = 8 if dtype == cpu.Z else 4
    pitch = lambda rows: -(-rows * rs // 128) * 128                      # bytes per plane column, as in gemm3m_on_device
    plane = lambda rows, cols: -(-pitch(rows) * cols // 256) * 256
    chunked = 0
    for ta in range(4):
        for tb in range(4):
            a, lda, b, ldb, c0, ldc = problem(rng, oracle, dtype, ta, tb, m, n, k, pad=(2, 3, 1))
            need = lambda kc: 3 * ((plane(kc, m) if ta & 1 else plane(m, kc)) + (plane(n, kc) if tb & 1 else plane(kc, n)) + plane(m, n))
            monkeypatch.setenv("B200_3M_WORKSPACE_BYTES", str(need(8)))       # chunks of 8 always fit; whether 16 or all 29 do depends on the pitch rounding
            got = c0.copy()
            before = sim.b200_launch_count()
            gemm3m(ta, tb, m, n, k, 0.7 - 0.9j, a, lda, b, ldb, 1.3 - 1.1j, got, ldc)
            launches = sim.b200_launch_count() - before
            assert sim.b200_last_kernel() == b"sim_combine3" and launches % 6 == 0, (ta, tb, launches)
            expect = 1 if need(29) <= need(8) else (2 if need(16) <= need(8) else 4)
            assert launches // 6 == expect, (ta, tb, launches, expect)
            chunked += launches // 6 > 1
            assert oracle.mmch(dtype, ta, tb, m, n, k, 0.7 - 0.9j, a, lda, b, ldb, 1.3 - 1.1j, c0, ldc, got, ldc) < 16.0, (ta, tb, "chunked")
    assert chunked >= 8
    monkeypatch.setenv("B200_3M_WORKSPACE_BYTES", "1000")
    got = c0.copy()
    gemm3m(3, 3, m, n, k, 0.7 - 0.9j, a, lda, b, ldb, 1.3 - 1.1j, got, ldc)
    assert sim.b200_last_kernel() != b"sim_combine3"
    monkeypatch.delenv("B200_3M_WORKSPACE_BYTES")
    # below the limit: the 4-multiply kernel, bit for bit the GEMM result; and ?gemm_ itself never takes the 3M path
    a, lda, b, ldb, c0, ldc = problem(rng, oracle, dtype, 0, 3, 7, 30, 30, pad=(1, 1, 1))
    got, want = c0.copy(), c0.copy()
    gemm3m(0, 3, 7, 30, 30, 0.7 - 0.9j, a, lda, b, ldb, 1.3 - 1.1j, got, ldc)
    assert sim.b200_last_kernel() != b"sim_combine3"
    fgemm(sim, dtype, 0, 3, 7, 30, 30, 0.7 - 0.9j, a, lda, b, ldb, 1.3 - 1.1j, want, ldc)
    assert np.array_equal(got.view(np.uint8), want.view(np.uint8))
    a, lda, b, ldb, c0, ldc = problem(rng, oracle, dtype, 1, 0, 40, 33, 29, pad=(1, 1, 1))
    got = c0.copy()
    fgemm(sim, dtype, 1, 0, 40, 33, 29, 0.7 - 0.9j, a, lda, b, ldb, 1.3 - 1.1j, got, ldc)
    assert sim.b200_last_kernel() != b"sim_combine3"
