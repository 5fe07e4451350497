/*
 * summa.cu -- the multi-GPU driver INSIDE the library: 2-D block-cyclic SUMMA of C over a P x Q grid of
 * processes (one per GPU), reached through C symbols (b200_summa_*, include/openblas_b200.h).
 *
 * What it replaces:
 *   driver/level3/level3_thread.c:804-862   choice of the nthreads_m x nthreads_n grid
 *   driver/level3/gemm_thread_mn.c:43-61    divide_rule[] (8 workers -> 2 x 4)
 *   level3_thread.c:219-532  inner_thread   every worker packs its share of B ONCE into a buffer the others
 *                                           read (job[].working flags), double buffered (DIVIDE_RATE 2)
 * The same structure, one process per GPU: every rank packs its local pieces of A and B once per call into
 * a WINDOW (dense, cudaMalloc'd, opened in the peers through CUDA IPC at window creation); for each k panel
 * a rank PULLS the slice it needs out of the owner's window with the copy engines over NVLink
 * (cudaMemcpyAsync on peer-mapped memory: no SM is involved, so the transfer really overlaps the persistent
 * 148-CTA DGEMM, which an NCCL broadcast kernel cannot: it has no SM to run on until the GEMM drains) into
 * one of two landing buffers, while the local product of the previous panel runs on the compute stream.
 * Panels owned by the rank itself are used in place.  k is never split across GPUs: no reduction, results
 * are deterministic.  Process-level synchronisation is two tiny NCCL all-reduces per call on the compute
 * stream (windows packed everywhere / everyone done pulling); NCCL is also the fallback transport
 * (ncclBroadcast of panels on row / column communicators, B200_SUMMA_TRANSPORT=nccl) and carries the IPC
 * handles at window creation.  NCCL is dlopen'ed (libnccl.so.2): the BLAS library itself does not link it.
 */
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "gemm_common.cuh"

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

char g_msg[512];
int fail(const char *what, const char *detail) {
  snprintf(g_msg, sizeof g_msg, "b200_summa: %s: %s", what, detail);
  b200::summa_set_error(g_msg);
  return 1;
}
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(#call, cudaGetErrorString(e_)); } while (0)
#define NC(call) do { ncclResult_t r_ = (call); if (r_ != 0) return fail(#call, g_nccl.GetErrorString(r_)); } while (0)

size_t elem_size(int dtype) { return b200_in_size(dtype); }
inline size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

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

struct b200_summa {
  int rank = 0, world = 1, P = 1, Q = 1, p = 0, q = 0, device = 0;
  int transport = 0;                 /* 1 = NCCL broadcast, 2 = copy-engine pull from peer windows */
  ncclComm_t comm = nullptr, row_comm = nullptr, col_comm = nullptr;
  cudaStream_t copy_a = nullptr, copy_b = nullptr;
  cudaEvent_t ev_entry = nullptr, ev_ready[2] = {nullptr, nullptr}, ev_ready_b[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
  int *flag = nullptr;               /* device int for the barrier all-reduces */
  /* windows: my packed A and B pieces (+ the peers' mapped views), landing buffers */
  char *win = nullptr; size_t win_bytes = 0, win_b_off = 0;
  std::vector<char *> peer_win;      /* by rank; [rank] = win */
  std::vector<size_t> peer_b_off;
  char *land = nullptr; size_t land_bytes = 0;
  uint64_t launches = 0;
};

namespace {

int barrier(b200_summa *h, cudaStream_t s) {
  if (h->world == 1) return 0;
  NC(g_nccl.AllReduce(h->flag, h->flag, 1, ncclInt32, ncclSum, h->comm, s));
  return 0;
}

/* (re)create the window for `need_a + need_b` bytes and exchange its IPC handle; collective over all ranks:
 * every rank reaches the same decision because the sizes are all-gathered first */
int ensure_window(b200_summa *h, size_t need_a, size_t need_b, cudaStream_t s) {
  const size_t b_off = round_up(need_a, 1024), need = b_off + round_up(need_b, 1024);
  struct Info { unsigned long long need, have, b_off; cudaIpcMemHandle_t handle; };
  static_assert(sizeof(Info) % 8 == 0, "all-gathered as bytes");
  std::vector<Info> all((size_t)h->world);
  Info mine; memset(&mine, 0, sizeof mine);
  mine.need = need; mine.have = h->win_bytes;
  if (h->world > 1) {
    Info *d = nullptr;
    CU(cudaMalloc((void **)&d, sizeof(Info) * (size_t)(h->world + 1)));
    CU(cudaMemcpyAsync(d + h->world, &mine, sizeof mine, cudaMemcpyHostToDevice, s));
    NC(g_nccl.AllGather(d + h->world, d, sizeof(Info), ncclInt8, h->comm, s));
    CU(cudaMemcpyAsync(all.data(), d, sizeof(Info) * (size_t)h->world, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    CU(cudaFree(d));
  } else {
    all[0] = mine;
  }
  bool grow = false;
  for (const Info &i : all) grow = grow || i.need > i.have;
  if (!grow) { h->win_b_off = b_off; h->peer_b_off.assign((size_t)h->world, 0); for (int r = 0; r < h->world; r++) h->peer_b_off[(size_t)r] = (size_t)all[(size_t)r].b_off_dummy_guard(); return 0; }
  return -1;
}

}  // namespace
