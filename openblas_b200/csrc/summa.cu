/*
 * summa.cu -- the multi-GPU driver INSIDE the library: 2-D block-cyclic SUMMA of C over a P x Q grid of
 * processes (one per GPU), reached through C symbols (b200_summa_*, include/openblas_b200.h).
 *
 * What it replaces:
 *   driver/level3/level3_thread.c:804-862   choice of the nthreads_m x nthreads_n grid
 *   driver/level3/gemm_thread_mn.c:43-61    divide_rule[] (8 workers -> 2 x 4)
 *   level3_thread.c:219-532  inner_thread   every worker packs its share of B ONCE into a buffer the others
 *                                           read, announces it through job[].working flags and waits on the
 *                                           others' flags, double buffered (DIVIDE_RATE 2)
 * The same structure, one process per GPU.  Every rank packs its local pieces of A and B once per call into a
 * WINDOW (cudaMalloc'd, opened in the peers through CUDA IPC), panel by panel, and ANNOUNCES each panel by
 * writing a counter into the peers' control blocks with a stream memory operation (cuStreamWriteValue32: no SM
 * involved).  For each k panel a rank waits on its own control block (cuStreamWaitValue32), PULLS the slice it
 * needs out of the owner's window with the copy engines over NVLink (cudaMemcpyAsync on peer-mapped memory) into
 * one of two landing buffers, while the local product of the previous panel runs on the compute stream.  Panels
 * the rank owns itself are used in place.  Nothing on this path needs an SM besides the GEMM itself -- an NCCL
 * broadcast kernel, by contrast, cannot co-reside with the persistent 148-CTA DGEMM (221 KB of shared memory per
 * SM) and only runs when the GEMM drains, which serialises transfer and compute (round 1: 0.88 efficiency at 8
 * GPUs).  k is never split across GPUs: no reduction, results are deterministic.
 *
 * NCCL (dlopen'ed libnccl.so.2 -- the BLAS library does not link it) bootstraps the process group from a unique
 * id the caller distributes, carries the IPC handles, and is the alternative transport
 * (B200_SUMMA_TRANSPORT=nccl: ncclBroadcast of the panels on row / column communicators, double buffered on a
 * communication stream) and the alternative synchronisation (B200_SUMMA_SYNC=nccl: all-reduce barriers).
 *
 * Operands may be device or host pointers: host pieces of A and B are uploaded straight into the window panel by
 * panel (so the first product starts when the first panels have landed), a host C is computed in a device copy
 * whose last update is cut into column strips that are downloaded while the next strip is computed.
 */
#include <cuda_runtime.h>
#include <cuda.h>
#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <cstddef>
#include <cstring>
#include <vector>
#include "gemm_common.cuh"
#include "shim.h"

namespace b200 {
int summa_local_gemm(int dtype, int64_t m, int64_t n, int64_t k, const void *alpha, const void *a, int64_t lda, const void *b,
                     int64_t ldb, const void *beta, void *c, int64_t ldc, cudaStream_t stream);   /* runtime.cu */
void summa_set_error(const char *msg);                                                           /* runtime.cu */
}

namespace {

/* ---- the slice of NCCL used here, resolved at run time ------------------------------------------- */
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt8 = 0, ncclInt32 = 2 };
enum { ncclSum = 0 };
struct Nccl {
  void *so = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t *, void *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool load() {
    if (so) return true;
    const char *names[] = {getenv("B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      if (!n) continue;
      so = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (so) break;
    }
    if (!so) return false;
#define SYM(field, name) *(void **)(&field) = dlsym(so, name); if (!field) return false;
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommSplit, "ncclCommSplit")
    SYM(CommDestroy, "ncclCommDestroy") SYM(Broadcast, "ncclBroadcast") SYM(AllReduce, "ncclAllReduce")
    SYM(AllGather, "ncclAllGather") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    return true;
  }
};
Nccl g_nccl;

/* ---- stream memory operations (driver API, resolved through the runtime) --------------------------- */
typedef CUresult (*StreamValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
StreamValue32Fn g_write32 = nullptr, g_wait32 = nullptr;
bool load_memops() {
  if (g_write32 && g_wait32) return true;
  void *p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return false;
  g_write32 = (StreamValue32Fn)p;
  if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return false;
  g_wait32 = (StreamValue32Fn)p;
  return true;
}

thread_local char g_msg[512];
int fail(const char *what, const char *detail) {
  snprintf(g_msg, sizeof g_msg, "b200_summa: %s: %s", what, detail);
  b200::summa_set_error(g_msg);
  return 1;
}
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(#call, cudaGetErrorString(e_)); } while (0)
#define NC(call) do { ncclResult_t r_ = (call); if (r_ != 0) return fail(#call, g_nccl.GetErrorString(r_)); } while (0)
#define DR(call) do { CUresult r_ = (call); if (r_ != CUDA_SUCCESS) { char b_[32]; snprintf(b_, sizeof b_, "CUresult %d", (int)r_); return fail(#call, b_); } } while (0)

inline size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

enum { TRANSPORT_NCCL = 1, TRANSPORT_PULL = 2 };
enum { SYNC_NCCL = 1, SYNC_MEMOPS = 2 };

/* control block: counters the peers write into (stream memory operations), indexed by the writer's rank */
struct Control {
  static const int kMaxRanks = 64;
  uint32_t packed_a[kMaxRanks];   /* writer has packed this many A panels (running count, see b200_summa::base) */
  uint32_t packed_b[kMaxRanks];
  uint32_t done_a[kMaxRanks];     /* writer has finished pulling A panels from me up to this count */
  uint32_t done_b[kMaxRanks];
};

}  // namespace

/* ScaLAPACK's NUMROC with source process 0: rows / columns of a block-cyclic dimension owned by iproc */
extern "C" B200_EXPORT int64_t b200_summa_numroc(int64_t n, int64_t nb, int iproc, int nprocs) {
  const int64_t nblocks = n / nb;
  int64_t base = (nblocks / nprocs) * nb;
  const int64_t extra = nblocks % nprocs;
  if (iproc < extra) base += nb;
  else if (iproc == extra) base += n % nb;
  return base;
}

/* The k panels of one sweep (a panel never crosses a distribution block, so each slice has one owner):
 * step s covers k0 = s * nb, width min(nb, k - k0); its A slice lives in grid column s % Q at local column
 * (s / Q) * nb, its B slice in grid row s % P at local row (s / P) * nb.  Written for tests / callers that want the
 * schedule; out arrays hold `steps` entries each (any may be NULL).  Returns the number of steps. */
extern "C" B200_EXPORT int64_t b200_summa_schedule(int64_t k, int64_t nb, int P, int Q, int64_t capacity, int64_t *k0, int64_t *width,
                                                   int *a_owner_col, int64_t *a_local_col, int *b_owner_row, int64_t *b_local_row) {
  if (k <= 0 || nb <= 0) return 0;
  const int64_t steps = (k + nb - 1) / nb;
  for (int64_t s = 0; s < steps && s < capacity; s++) {
    if (k0) k0[s] = s * nb;
    if (width) width[s] = k - s * nb < nb ? k - s * nb : nb;
    if (a_owner_col) a_owner_col[s] = (int)(s % Q);
    if (a_local_col) a_local_col[s] = (s / Q) * nb;
    if (b_owner_row) b_owner_row[s] = (int)(s % P);
    if (b_local_row) b_local_row[s] = (s / P) * nb;
  }
  return steps;
}

/* squarest grid with P <= Q (gemm_thread_mn.c:43-61 divide_rule[]: 2 -> 1x2, 4 -> 2x2, 8 -> 2x4, 16 -> 4x4) */
extern "C" B200_EXPORT void b200_summa_grid(int world, int *P, int *Q) {
  int p = 1;
  for (int d = 1; d * d <= world; d++) if (world % d == 0) p = d;
  *P = p; *Q = world / p;
}

struct b200_summa {
  int rank = 0, world = 1, P = 1, Q = 1, p = 0, q = 0, device = 0;
  int transport = TRANSPORT_PULL, sync = SYNC_MEMOPS;
  ncclComm_t comm = nullptr, row_comm = nullptr, col_comm = nullptr;
  cudaStream_t s_pack = nullptr, s_a = nullptr, s_b = nullptr, s_out = nullptr;
  cudaEvent_t ev_entry = nullptr, ev_packed = nullptr, ev_exit = nullptr;
  cudaEvent_t ev_a[2] = {nullptr, nullptr}, ev_b[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
  cudaEvent_t ev_strip[2] = {nullptr, nullptr};
  cudaEvent_t ev_cup = nullptr;                   /* a host C has been uploaded */
  std::vector<cudaEvent_t> ev_pack_a, ev_pack_b;  /* my local panel j is in the window (for the products that use it in place) */
  int *flag = nullptr;                            /* device int for the NCCL barriers */
  Control *ctl = nullptr;                         /* my control block (device memory, IPC-exported) */
  std::vector<Control *> peer_ctl;                /* by rank; [rank] = ctl */
  char *win = nullptr; size_t win_bytes = 0;      /* my window: [A pieces | B pieces, panel-major] */
  std::vector<char *> peer_win;                   /* by rank; [rank] = win; only grid-row / grid-column peers are opened */
  char *land = nullptr; size_t land_bytes = 0;    /* two landing slots of (A panel + B panel) */
  char *cdev = nullptr; size_t cdev_bytes = 0;    /* device copy of a host C */
  uint32_t count = 0;                             /* running panel counter: identical on every rank (same call sequence) */
  uint32_t last_end = 0;                          /* value of `count` at the end of the previous call (0 = none yet) */
  uint64_t launches = 0, calls = 0;
  int rank_of(int pp, int qq) const { return pp * Q + qq; }
};

namespace {

int nccl_barrier(b200_summa *h, cudaStream_t s) {
  if (h->world == 1) return 0;
  NC(g_nccl.AllReduce(h->flag, h->flag, 1, ncclInt32, ncclSum, h->comm, s));
  return 0;
}

bool is_host_pointer(const void *p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return true; }
  return !(at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged);
}

struct Handles { cudaIpcMemHandle_t win, ctl; };

/* all-gather `mine` (POD) from every rank through NCCL */
template <class T> int gather_pod(b200_summa *h, const T &mine, std::vector<T> &all, cudaStream_t s) {
  all.assign((size_t)h->world, mine);
  if (h->world == 1) return 0;
  char *d = nullptr;
  const size_t sz = round_up(sizeof(T), 16);
  CU(cudaMalloc((void **)&d, sz * (size_t)(h->world + 1)));
  CU(cudaMemcpyAsync(d + sz * (size_t)h->world, &mine, sizeof(T), cudaMemcpyHostToDevice, s));
  NC(g_nccl.AllGather(d + sz * (size_t)h->world, d, sz, ncclInt8, h->comm, s));
  std::vector<char> host(sz * (size_t)h->world);
  CU(cudaMemcpyAsync(host.data(), d, host.size(), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  CU(cudaFree(d));
  for (int r = 0; r < h->world; r++) memcpy(&all[(size_t)r], host.data() + sz * (size_t)r, sizeof(T));
  return 0;
}

bool is_peer(const b200_summa *h, int r) { return r != h->rank && (r / h->Q == h->p || r % h->Q == h->q); }

int close_window(b200_summa *h) {
  for (int r = 0; r < (int)h->peer_win.size(); r++)
    if (r != h->rank && h->peer_win[(size_t)r]) { CU(cudaIpcCloseMemHandle(h->peer_win[(size_t)r])); h->peer_win[(size_t)r] = nullptr; }
  return 0;
}

/* (re)create the window when this call needs a bigger one.  `need` is computed from the GLOBAL problem (the largest
 * local piece of any rank), so every rank takes the same decision without talking.  Collective. */
int ensure_window(b200_summa *h, size_t need, cudaStream_t s) {
  if (need <= h->win_bytes) return 0;
  /* nobody may still be pulling from the old window: drain my streams, then meet everyone */
  CU(cudaStreamSynchronize(h->s_a)); CU(cudaStreamSynchronize(h->s_b)); CU(cudaStreamSynchronize(h->s_pack));
  if (nccl_barrier(h, s)) return 1;
  CU(cudaStreamSynchronize(s));
  if (close_window(h)) return 1;
  if (nccl_barrier(h, s)) return 1;
  CU(cudaStreamSynchronize(s));
  if (h->win) { CU(cudaFree(h->win)); h->win = nullptr; h->win_bytes = 0; }
  const size_t want = round_up(need + need / 16, (size_t)2 << 20);
  CU(cudaMalloc((void **)&h->win, want));
  h->win_bytes = want;
  h->peer_win.assign((size_t)h->world, nullptr);
  h->peer_win[(size_t)h->rank] = h->win;
  if (h->world > 1 && h->transport == TRANSPORT_PULL) {
    cudaIpcMemHandle_t mine;
    CU(cudaIpcGetMemHandle(&mine, h->win));
    std::vector<cudaIpcMemHandle_t> all;
    if (gather_pod(h, mine, all, s)) return 1;
    for (int r = 0; r < h->world; r++)
      if (is_peer(h, r)) CU(cudaIpcOpenMemHandle((void **)&h->peer_win[(size_t)r], all[(size_t)r], cudaIpcMemLazyEnablePeerAccess));
  }
  return 0;
}

int ensure_buffer(char **buf, size_t *have, size_t need) {
  if (need <= *have) return 0;
  if (*buf) { CU(cudaDeviceSynchronize()); CU(cudaFree(*buf)); *buf = nullptr; *have = 0; }
  CU(cudaMalloc((void **)buf, need));
  *have = need;
  return 0;
}

/* column-major block copy, any direction, on stream s */
int copy2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height, cudaStream_t s) {
  if (width == 0 || height == 0) return 0;
  if (dpitch == width && spitch == width) CU(cudaMemcpyAsync(dst, src, width * height, cudaMemcpyDefault, s));
  else CU(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyDefault, s));
  return 0;
}

int signal_peers(b200_summa *h, cudaStream_t s, size_t field_offset, bool along_row, uint32_t value) {
  if (h->sync != SYNC_MEMOPS) return 0;
  const int n = along_row ? h->Q : h->P;
  for (int i = 0; i < n; i++) {
    const int r = along_row ? h->rank_of(h->p, i) : h->rank_of(i, h->q);
    if (r == h->rank) continue;
    char *slot = (char *)h->peer_ctl[(size_t)r] + field_offset + sizeof(uint32_t) * (size_t)h->rank;
    DR(g_write32((CUstream)s, (CUdeviceptr)slot, value, 0));
  }
  return 0;
}
int wait_for(b200_summa *h, cudaStream_t s, size_t field_offset, int writer, uint32_t value) {
  if (h->sync != SYNC_MEMOPS || writer == h->rank) return 0;
  char *slot = (char *)h->ctl + field_offset + sizeof(uint32_t) * (size_t)writer;
  DR(g_wait32((CUstream)s, (CUdeviceptr)slot, value, CU_STREAM_WAIT_VALUE_GEQ));
  return 0;
}

}  // namespace

extern "C" {

/* 128 bytes the caller hands to every rank (MPI_Bcast, torch.distributed, a file ...): NCCL's unique id */
B200_EXPORT int b200_summa_unique_id(void *id128) {
  if (!g_nccl.load()) return fail("dlopen", "libnccl.so.2 not found (set B200_NCCL_LIB)");
  ncclUniqueId id;
  NC(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof id);
  return 0;
}

B200_EXPORT int b200_summa_destroy(b200_summa *h);

/* Collective over `world` processes, each bound to its own GPU (the current device of the calling thread). */
static int summa_setup(b200_summa *h, const void *id128, int rank, int world);

/* local release of a handle whose creation failed half way: no collective call (the peers may have failed elsewhere),
 * so NCCL communicators that already exist are left to process exit */
static void summa_abandon(b200_summa *h) {
  close_window(h);
  for (int r = 0; r < (int)h->peer_ctl.size(); r++)
    if (r != h->rank && h->peer_ctl[(size_t)r]) cudaIpcCloseMemHandle(h->peer_ctl[(size_t)r]);
  if (h->ctl) cudaFree(h->ctl);
  if (h->flag) cudaFree(h->flag);
  cudaEvent_t evs[] = {h->ev_cup, h->ev_entry, h->ev_packed, h->ev_exit, h->ev_a[0], h->ev_a[1], h->ev_b[0], h->ev_b[1], h->ev_free[0], h->ev_free[1],
                       h->ev_strip[0], h->ev_strip[1]};
  for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
  cudaStream_t ss[] = {h->s_pack, h->s_a, h->s_b, h->s_out};
  for (cudaStream_t st : ss) if (st) cudaStreamDestroy(st);
  cudaGetLastError();
  delete h;
}

B200_EXPORT int b200_summa_create(b200_summa **out, const void *id128, int rank, int world, int P, int Q) {
  if (!out) return fail("b200_summa_create", "no place for the handle");
  *out = nullptr;
  if (world < 1 || rank < 0 || rank >= world || P < 1 || Q < 1 || P * Q != world || world > Control::kMaxRanks) return fail("b200_summa_create", "bad rank / world / grid");
  if (world > 1 && !id128) return fail("b200_summa_create", "no unique id (b200_summa_unique_id on one rank, distributed to all)");
  b200_summa *h = new b200_summa();
  h->rank = rank; h->world = world; h->P = P; h->Q = Q; h->p = rank / Q; h->q = rank % Q;
  if (summa_setup(h, id128, rank, world)) { summa_abandon(h); return 1; }   /* the thread's error text says what failed */
  *out = h;
  return 0;
}

static int summa_setup(b200_summa *h, const void *id128, int rank, int world) {
  CU(cudaGetDevice(&h->device));
  if (b200_init(h->device) != 0) return 1;
  const char *tv = getenv("B200_SUMMA_TRANSPORT"), *sv = getenv("B200_SUMMA_SYNC");
  h->transport = (tv && !strcmp(tv, "nccl")) ? TRANSPORT_NCCL : TRANSPORT_PULL;
  h->sync = (sv && !strcmp(sv, "nccl")) ? SYNC_NCCL : SYNC_MEMOPS;
  if (h->sync == SYNC_MEMOPS && !load_memops()) h->sync = SYNC_NCCL;
  if (h->transport == TRANSPORT_NCCL) h->sync = SYNC_NCCL;       /* broadcasts are collective: no flags needed */
  CU(cudaStreamCreateWithFlags(&h->s_pack, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&h->s_a, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&h->s_b, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
  cudaEvent_t *evs[] = {&h->ev_cup, &h->ev_entry, &h->ev_packed, &h->ev_exit, &h->ev_a[0], &h->ev_a[1], &h->ev_b[0], &h->ev_b[1], &h->ev_free[0],
                        &h->ev_free[1], &h->ev_strip[0], &h->ev_strip[1]};
  for (cudaEvent_t *e : evs) CU(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  CU(cudaMalloc((void **)&h->flag, sizeof(int)));
  CU(cudaMemset(h->flag, 0, sizeof(int)));
  CU(cudaMalloc((void **)&h->ctl, sizeof(Control)));
  CU(cudaMemset(h->ctl, 0, sizeof(Control)));
  CU(cudaDeviceSynchronize());
  h->peer_ctl.assign((size_t)world, nullptr);
  h->peer_ctl[(size_t)rank] = h->ctl;
  h->peer_win.assign((size_t)world, nullptr);
  if (world > 1) {
    if (!g_nccl.load()) return fail("dlopen", "libnccl.so.2 not found (set B200_NCCL_LIB)");
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    NC(g_nccl.CommInitRank(&h->comm, world, id, rank));
    if (h->transport == TRANSPORT_NCCL) {
      NC(g_nccl.CommSplit(h->comm, h->p, h->q, &h->row_comm, nullptr));
      NC(g_nccl.CommSplit(h->comm, h->q, h->p, &h->col_comm, nullptr));
    }
    if (h->sync == SYNC_MEMOPS) {
      cudaIpcMemHandle_t mine;
      CU(cudaIpcGetMemHandle(&mine, h->ctl));
      std::vector<cudaIpcMemHandle_t> all;
      if (gather_pod(h, mine, all, h->s_pack)) return 1;
      for (int r = 0; r < world; r++)
        if (is_peer(h, r)) CU(cudaIpcOpenMemHandle((void **)&h->peer_ctl[(size_t)r], all[(size_t)r], cudaIpcMemLazyEnablePeerAccess));
    }
    if (nccl_barrier(h, h->s_pack)) return 1;
    CU(cudaStreamSynchronize(h->s_pack));
  }
  return 0;
}

B200_EXPORT int b200_summa_destroy(b200_summa *h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  if (h->world > 1 && h->comm) { nccl_barrier(h, h->s_pack); cudaStreamSynchronize(h->s_pack); }
  close_window(h);
  for (int r = 0; r < (int)h->peer_ctl.size(); r++)
    if (r != h->rank && h->peer_ctl[(size_t)r]) cudaIpcCloseMemHandle(h->peer_ctl[(size_t)r]);
  if (h->world > 1 && h->comm) { nccl_barrier(h, h->s_pack); cudaStreamSynchronize(h->s_pack); }
  if (h->win) cudaFree(h->win);
  if (h->land) cudaFree(h->land);
  if (h->cdev) cudaFree(h->cdev);
  if (h->ctl) cudaFree(h->ctl);
  if (h->flag) cudaFree(h->flag);
  if (h->row_comm) g_nccl.CommDestroy(h->row_comm);
  if (h->col_comm) g_nccl.CommDestroy(h->col_comm);
  if (h->comm) g_nccl.CommDestroy(h->comm);
  for (cudaEvent_t e : h->ev_pack_a) cudaEventDestroy(e);
  for (cudaEvent_t e : h->ev_pack_b) cudaEventDestroy(e);
  cudaEvent_t evs[] = {h->ev_cup, h->ev_entry, h->ev_packed, h->ev_exit, h->ev_a[0], h->ev_a[1], h->ev_b[0], h->ev_b[1], h->ev_free[0], h->ev_free[1],
                       h->ev_strip[0], h->ev_strip[1]};
  for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
  cudaStream_t ss[] = {h->s_pack, h->s_a, h->s_b, h->s_out};
  for (cudaStream_t s : ss) if (s) cudaStreamDestroy(s);
  cudaGetLastError();
  delete h;
  return 0;
}

B200_EXPORT uint64_t b200_summa_launches(const b200_summa *h) { return h ? h->launches : 0; }
B200_EXPORT const char *b200_summa_describe(const b200_summa *h) {
  static thread_local char buf[320];
  if (!h) return "summa: no handle";
  snprintf(buf, sizeof buf, "summa %dx%d rank %d (p=%d,q=%d): panels by %s, synchronised by %s", h->P, h->Q, h->rank, h->p, h->q,
           h->transport == TRANSPORT_PULL ? "copy-engine pulls from peer windows (CUDA IPC over NVLink)" : "ncclBroadcast on row/column communicators",
           h->sync == SYNC_MEMOPS ? "stream memory operations on peer control blocks" : "NCCL");
  return buf;
}

/* C := alpha * A * B + beta * C for the GLOBAL m x n x k problem, every matrix distributed 2-D block-cyclically
 * with block nb over the P x Q grid (ScaLAPACK layout, source process 0): this rank holds
 *   a_loc  numroc(m, nb, p, P) x numroc(k, nb, q, Q), leading dimension lda
 *   b_loc  numroc(k, nb, p, P) x numroc(n, nb, q, Q), leading dimension ldb
 *   c_loc  numroc(m, nb, p, P) x numroc(n, nb, q, Q), leading dimension ldc
 * column-major, device or host memory.  dtype: B200_S / D / C / Z (alpha, beta: 1 or 2 values of the type).
 * Collective; enqueued on `stream` (the local products run there) and asynchronous for device operands; with a host
 * c_loc the call returns when C is complete in host memory. */
B200_EXPORT int b200_summa_gemm(b200_summa *h, int dtype, int64_t m, int64_t n, int64_t k, int64_t nb, const void *alpha,
                                const void *a_loc, int64_t lda, const void *b_loc, int64_t ldb, const void *beta, void *c_loc,
                                int64_t ldc, void *stream) {
  if (!h) return fail("b200_summa_gemm", "no handle");
  if (dtype == B200_SB) return fail("b200_summa_gemm", "bf16 operands are not distributed by this driver");
  if (m < 0 || n < 0 || k < 0 || nb <= 0) return fail("b200_summa_gemm", "bad extents");
  cudaStream_t sc = (cudaStream_t)stream;
  const size_t es = b200_in_size(dtype);
  const int P = h->P, Q = h->Q, p = h->p, q = h->q;
  const int64_t m_loc = b200_summa_numroc(m, nb, p, P), n_loc = b200_summa_numroc(n, nb, q, Q);
  const int64_t ka_loc = b200_summa_numroc(k, nb, q, Q), kb_loc = b200_summa_numroc(k, nb, p, P);
  const int64_t m_max = b200_summa_numroc(m, nb, 0, P), n_max = b200_summa_numroc(n, nb, 0, Q);
  const int64_t ka_max = b200_summa_numroc(k, nb, 0, Q), kb_max = b200_summa_numroc(k, nb, 0, P);
  if (lda < (m_loc > 1 ? m_loc : 1) || ldb < (kb_loc > 1 ? kb_loc : 1) || ldc < (m_loc > 1 ? m_loc : 1)) return fail("b200_summa_gemm", "leading dimension smaller than the local piece");
  const int64_t steps = k > 0 ? (k + nb - 1) / nb : 0;
  const bool work = m_loc > 0 && n_loc > 0;
  CU(cudaSetDevice(h->device));
  h->calls++;

  /* window: [A pieces, m_loc x ka_loc dense | B pieces, one dense (w x n_loc) block per local k block]; the offsets are
   * the same on every rank (sized for the largest local piece) */
  const size_t b_off = round_up((size_t)m_max * (size_t)ka_max * es, 1024);
  const size_t need = b_off + round_up((size_t)kb_max * (size_t)n_max * es, 1024);
  if (ensure_window(h, need, h->s_pack)) return 1;
  const size_t a_slot = round_up((size_t)m_max * (size_t)nb * es, 1024), b_slot = round_up((size_t)nb * (size_t)n_max * es, 1024);
  if (ensure_buffer(&h->land, &h->land_bytes, 2 * (a_slot + b_slot))) return 1;
  const bool c_host = work && is_host_pointer(c_loc);
  char *c_dev = (char *)c_loc;
  int64_t ldc_dev = ldc;
  if (c_host) {
    ldc_dev = (int64_t)(round_up((size_t)m_loc * es, 128) / es);
    if (ensure_buffer(&h->cdev, &h->cdev_bytes, (size_t)ldc_dev * (size_t)n_loc * es)) return 1;
    c_dev = h->cdev;
  }
  const double *be_d = (const double *)beta; const float *be_f = (const float *)beta;
  const bool dbl = dtype == B200_D || dtype == B200_Z, cplx = dtype == B200_C || dtype == B200_Z;
  const bool beta_zero = dbl ? (be_d[0] == 0.0 && (!cplx || be_d[1] == 0.0)) : (be_f[0] == 0.f && (!cplx || be_f[1] == 0.f));
  const double one_d[2] = {1.0, 0.0}; const float one_f[2] = {1.f, 0.f};
  const void *one = dbl ? (const void *)one_d : (const void *)one_f;

  /* ---- pack (or upload) my pieces into the window, panel by panel, announcing each panel ------------------- */
  const uint32_t base = h->count;
  const uint32_t end = base + (uint32_t)steps + 1;
  h->count = end;
  CU(cudaEventRecord(h->ev_entry, sc));
  CU(cudaStreamWaitEvent(h->s_pack, h->ev_entry, 0));        /* operands produced on the caller's stream are complete */
  /* the landing buffers are free again only when the local products of the PREVIOUS call (on the caller's stream) have
   * read them: without this a fast owner's flag would let the first pulls of this call overwrite a slot under a slow
   * rank's last product */
  CU(cudaStreamWaitEvent(h->s_a, h->ev_entry, 0));
  CU(cudaStreamWaitEvent(h->s_b, h->ev_entry, 0));
  if (h->last_end) {                                          /* my window is free again when every peer finished pulling */
    for (int qq = 0; qq < Q; qq++) if (wait_for(h, h->s_pack, offsetof(Control, done_a), h->rank_of(p, qq), h->last_end)) return 1;
    for (int pp = 0; pp < P; pp++) if (wait_for(h, h->s_pack, offsetof(Control, done_b), h->rank_of(pp, q), h->last_end)) return 1;
  }
  if (c_host && !beta_zero)
    if (copy2d(c_dev, (size_t)ldc_dev * es, c_loc, (size_t)ldc * es, (size_t)m_loc * es, (size_t)n_loc, h->s_pack)) return 1;
  CU(cudaEventRecord(h->ev_cup, h->s_pack));
  {
    const int64_t na = ka_loc > 0 ? (ka_loc + nb - 1) / nb : 0, nbk = kb_loc > 0 ? (kb_loc + nb - 1) / nb : 0;
    while ((int64_t)h->ev_pack_a.size() < na) { cudaEvent_t e; CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); h->ev_pack_a.push_back(e); }
    while ((int64_t)h->ev_pack_b.size() < nbk) { cudaEvent_t e; CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); h->ev_pack_b.push_back(e); }
    for (int64_t j = 0; j < (na > nbk ? na : nbk); j++) {
      if (j < na) {
        const int64_t c0 = j * nb, w = ka_loc - c0 < nb ? ka_loc - c0 : nb;
        if (copy2d(h->win + (size_t)c0 * (size_t)m_loc * es, (size_t)m_loc * es, (const char *)a_loc + (size_t)c0 * (size_t)lda * es,
                   (size_t)lda * es, (size_t)m_loc * es, (size_t)w, h->s_pack)) return 1;
        if (signal_peers(h, h->s_pack, offsetof(Control, packed_a), true, base + (uint32_t)j + 1)) return 1;
        CU(cudaEventRecord(h->ev_pack_a[(size_t)j], h->s_pack));
      }
      if (j < nbk) {
        const int64_t r0 = j * nb, w = kb_loc - r0 < nb ? kb_loc - r0 : nb;
        if (copy2d(h->win + b_off + (size_t)r0 * (size_t)n_loc * es, (size_t)w * es, (const char *)b_loc + (size_t)r0 * es, (size_t)ldb * es,
                   (size_t)w * es, (size_t)n_loc, h->s_pack)) return 1;
        if (signal_peers(h, h->s_pack, offsetof(Control, packed_b), false, base + (uint32_t)j + 1)) return 1;
        CU(cudaEventRecord(h->ev_pack_b[(size_t)j], h->s_pack));
      }
    }
  }
  CU(cudaEventRecord(h->ev_packed, h->s_pack));
  if (h->sync == SYNC_NCCL && h->transport == TRANSPORT_PULL && h->world > 1) {
    /* barrier: every window is packed (an NCCL kernel: runs before the first product occupies the SMs) */
    if (nccl_barrier(h, h->s_pack)) return 1;
    CU(cudaEventRecord(h->ev_packed, h->s_pack));
  }
  if (h->sync == SYNC_NCCL) {       /* with flags every pull waits for exactly the panel it needs instead */
    CU(cudaStreamWaitEvent(h->s_a, h->ev_packed, 0));
    CU(cudaStreamWaitEvent(h->s_b, h->ev_packed, 0));
  }
  CU(cudaStreamWaitEvent(sc, h->ev_cup, 0));

  /* ---- the sweep ------------------------------------------------------------------------------------------
   * A host C is swept in two column halves (each one through all k panels): the first half goes back to the host
   * while the second one is computed, so only half of the download is exposed at the end.  The A panels are pulled
   * again for the second half (the windows still hold them; a pull costs a few per cent of a local product). */
  /* B200_SUMMA_HOST_HALVES: 0 = one sweep always; n > 1 = smallest local column count that is split (default 8192; tests lower it) */
  static const int64_t halves_from = getenv("B200_SUMMA_HOST_HALVES") ? (atoll(getenv("B200_SUMMA_HOST_HALVES")) == 1 ? 8192 : atoll(getenv("B200_SUMMA_HOST_HALVES"))) : 8192;
  const int nsweeps = (c_host && halves_from > 0 && h->transport == TRANSPORT_PULL && n_loc >= halves_from && steps >= 2) ? 2 : 1;
  auto fetch = [&](int64_t s, int slot, int64_t J0, int64_t J1, const char **a_pan, const char **b_pan) -> int {
    const int64_t w = k - s * nb < nb ? k - s * nb : nb;
    const int qa = (int)(s % Q), pb = (int)(s % P);
    const int64_t ca = (s / Q) * nb, rb = (s / P) * nb;
    char *la = h->land + (size_t)slot * (a_slot + b_slot), *lb = la + a_slot;
    const size_t a_bytes = (size_t)m_loc * (size_t)w * es, b_bytes = (size_t)w * (size_t)(J1 - J0) * es;
    const size_t b_at = b_off + ((size_t)rb * (size_t)n_loc + (size_t)J0 * (size_t)w) * es;      /* columns J0.. of the dense w x n_loc panel */
    if (h->transport == TRANSPORT_PULL) {
      const int ra = h->rank_of(p, qa), rbk = h->rank_of(pb, q);
      if (qa == q) { *a_pan = h->win + (size_t)ca * (size_t)m_loc * es; CU(cudaStreamWaitEvent(h->s_a, h->ev_pack_a[(size_t)(s / Q)], 0)); }
      else {
        if (wait_for(h, h->s_a, offsetof(Control, packed_a), ra, base + (uint32_t)(s / Q) + 1)) return 1;
        if (a_bytes) CU(cudaMemcpyAsync(la, h->peer_win[(size_t)ra] + (size_t)ca * (size_t)m_loc * es, a_bytes, cudaMemcpyDefault, h->s_a));
        *a_pan = la;
      }
      if (pb == p) { *b_pan = h->win + b_at; CU(cudaStreamWaitEvent(h->s_b, h->ev_pack_b[(size_t)(s / P)], 0)); }
      else {
        if (wait_for(h, h->s_b, offsetof(Control, packed_b), rbk, base + (uint32_t)(s / P) + 1)) return 1;
        if (b_bytes) CU(cudaMemcpyAsync(lb, h->peer_win[(size_t)rbk] + b_at, b_bytes, cudaMemcpyDefault, h->s_b));
        *b_pan = lb;
      }
    } else {
      /* both broadcasts on ONE stream in the same order on every rank (two communicators used concurrently from two
       * streams may deadlock when their kernels cannot co-reside) */
      *a_pan = la; *b_pan = lb;
      if (Q > 1) { if (a_bytes) NC(g_nccl.Broadcast(h->win + (size_t)ca * (size_t)m_loc * es, la, a_bytes, ncclInt8, qa, h->row_comm, h->s_a)); }
      else *a_pan = h->win + (size_t)ca * (size_t)m_loc * es;
      if (P > 1) { if (b_bytes) NC(g_nccl.Broadcast(h->win + b_at, lb, b_bytes, ncclInt8, pb, h->col_comm, h->s_a)); }
      else *b_pan = h->win + b_at;
    }
    CU(cudaEventRecord(h->ev_a[slot], h->s_a));
    CU(cudaEventRecord(h->ev_b[slot], h->s_b));
    return 0;
  };

  const bool pulls = work || h->transport == TRANSPORT_NCCL;     /* broadcasts are collective even for a rank with no C */
  for (int v = 0; v < nsweeps; v++) {
    const int64_t J0 = n_loc * v / nsweeps, J1 = n_loc * (v + 1) / nsweeps, nj = J1 - J0;
    const char *a_pan[2] = {nullptr, nullptr}, *b_pan[2] = {nullptr, nullptr};
    if (v > 0) {                      /* both landing slots are free again when the previous sweep's last products have read them */
      for (int sl = 0; sl < 2; sl++) { CU(cudaStreamWaitEvent(h->s_a, h->ev_free[sl], 0)); CU(cudaStreamWaitEvent(h->s_b, h->ev_free[sl], 0)); }
    }
    if (steps > 0 && pulls && fetch(0, 0, J0, J1, &a_pan[0], &b_pan[0])) return 1;
    for (int64_t s = 0; s < steps; s++) {
      const int slot = (int)(s & 1);
      if (s + 1 < steps && pulls) {
        const int ns = (int)((s + 1) & 1);
        if (s >= 1) { CU(cudaStreamWaitEvent(h->s_a, h->ev_free[ns], 0)); CU(cudaStreamWaitEvent(h->s_b, h->ev_free[ns], 0)); }
        if (fetch(s + 1, ns, J0, J1, &a_pan[ns], &b_pan[ns])) return 1;
      }
      if (!work) continue;
      const int64_t w = k - s * nb < nb ? k - s * nb : nb;
      CU(cudaStreamWaitEvent(sc, h->ev_a[slot], 0));
      CU(cudaStreamWaitEvent(sc, h->ev_b[slot], 0));
      const void *bs = s == 0 ? beta : one;
      char *c_sweep = c_dev + (size_t)J0 * (size_t)ldc_dev * es;
      if (c_host && s + 1 == steps) {
        /* last update of these columns of a host C: strips, each downloaded while the next one is computed */
        const int strips = nj >= 4096 ? 4 : 1;
        for (int t = 0; t < strips; t++) {
          const int64_t j0 = nj * t / strips, j1 = nj * (t + 1) / strips;
          if (b200::summa_local_gemm(dtype, m_loc, j1 - j0, w, alpha, a_pan[slot], m_loc, b_pan[slot] + (size_t)j0 * (size_t)w * es, w, bs,
                                     c_sweep + (size_t)j0 * (size_t)ldc_dev * es, ldc_dev, sc)) return 1;
          h->launches++;
          CU(cudaEventRecord(h->ev_strip[t & 1], sc));
          CU(cudaStreamWaitEvent(h->s_out, h->ev_strip[t & 1], 0));
          if (copy2d((char *)c_loc + (size_t)(J0 + j0) * (size_t)ldc * es, (size_t)ldc * es, c_sweep + (size_t)j0 * (size_t)ldc_dev * es,
                     (size_t)ldc_dev * es, (size_t)m_loc * es, (size_t)(j1 - j0), h->s_out)) return 1;
        }
      } else {
        if (b200::summa_local_gemm(dtype, m_loc, nj, w, alpha, a_pan[slot], m_loc, b_pan[slot], w, bs, c_sweep, ldc_dev, sc)) return 1;
        h->launches++;
      }
      CU(cudaEventRecord(h->ev_free[slot], sc));
    }
  }
  if (steps == 0 && work) {           /* k == 0: C := beta * C */
    if (b200::summa_local_gemm(dtype, m_loc, n_loc, 0, alpha, c_dev, ldc_dev, c_dev, ldc_dev, beta, c_dev, ldc_dev, sc)) return 1;
    h->launches++;
    if (c_host) {
      CU(cudaEventRecord(h->ev_strip[0], sc));
      CU(cudaStreamWaitEvent(h->s_out, h->ev_strip[0], 0));
      if (copy2d(c_loc, (size_t)ldc * es, c_dev, (size_t)ldc_dev * es, (size_t)m_loc * es, (size_t)n_loc, h->s_out)) return 1;
    }
  }

  /* ---- tell the owners I am done with their windows; keep the caller's stream ordered after my helpers ------- */
  if (signal_peers(h, h->s_a, offsetof(Control, done_a), true, end)) return 1;
  if (signal_peers(h, h->s_b, offsetof(Control, done_b), false, end)) return 1;
  h->last_end = end;
  if (h->sync == SYNC_NCCL && h->transport == TRANSPORT_PULL && h->world > 1) {
    /* barrier: everyone has finished pulling (after my own last pulls), before anyone repacks */
    CU(cudaEventRecord(h->ev_exit, h->s_b));
    CU(cudaStreamWaitEvent(h->s_a, h->ev_exit, 0));
    if (nccl_barrier(h, h->s_a)) return 1;
  }
  CU(cudaEventRecord(h->ev_exit, h->s_a));
  CU(cudaStreamWaitEvent(sc, h->ev_exit, 0));
  if (h->sync == SYNC_NCCL && h->transport == TRANSPORT_PULL && h->world > 1) {
    /* the next call's pack must not start before the barrier */
    CU(cudaStreamWaitEvent(h->s_pack, h->ev_exit, 0));
  }
  if (c_host) {
    CU(cudaStreamSynchronize(h->s_out));
    CU(cudaStreamSynchronize(sc));
  }
  return 0;
}

}  // extern "C"
