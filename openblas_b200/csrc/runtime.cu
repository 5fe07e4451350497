/*
 * runtime.cu -- device-side runtime of the library: lazy CUDA initialisation, a pool of
 * per-call contexts (stream + device workspace + pinned staging), host<->device staging
 * that honours leading dimensions, and the kernel dispatcher.
 *
 * It takes over the ROLE of the reference's runtime pieces on the GEMM path without
 * resembling them: blas_memory_alloc/free (driver/others/memory.c:1161-1357) handed out
 * pre-mapped packing buffers under a lock -> here a pool of contexts is handed out under a
 * lock; gotoblas_init (memory.c:1508-1565) ran at load time -> here nothing touches CUDA
 * until the first GEMM call, so a process may fork() before its first call; exec_blas
 * (blas_server.c:784-862) fanned work out to threads -> here one stream per concurrent
 * caller.  Calls are synchronous (C is complete in caller-visible memory on return) and
 * re-entrant; results are run-to-run deterministic (no atomics in any reduction).
 */
#include <cuda_runtime.h>
#include <atomic>
#include <mutex>
#include <vector>
#include <thread>
#include <condition_variable>
#include <functional>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>
#include "gemm_common.cuh"

namespace b200 {

/* Size thresholds of the host staging paths.  B200_HOSTSIM (tests/hostsim: this file compiled by the
 * host compiler against a host-memory stand-in for the CUDA runtime and CPU stand-ins for the kernels,
 * to test the staging / scheduling logic without a GPU) shrinks them so that every path is reached
 * with matrices a CPU oracle can multiply; the product build never defines it. */
#ifdef B200_HOSTSIM
#define B200_SMALL_BYTES   ((size_t)16 << 10)
#define B200_SLOT_BYTES    ((size_t)8 << 10)
#define B200_PANEL_MIN     32
#define B200_PANEL_BIG     64
#define B200_PANEL_BIG_AT  256
#define B200_PIPE_BYTES    ((size_t)64 << 10)
#define B200_BATCH_BYTES   ((size_t)256 << 10)
#define B200_POOL_BYTES    ((size_t)4 << 10)
#else
#define B200_SMALL_BYTES   ((size_t)4 << 20)
#define B200_SLOT_BYTES    ((size_t)32 << 20)
#define B200_PANEL_MIN     2048
#define B200_PANEL_BIG     4096
#define B200_PANEL_BIG_AT  8192
#define B200_PIPE_BYTES    ((size_t)128 << 20)
#define B200_BATCH_BYTES   ((size_t)64 << 20)
#define B200_POOL_BYTES    ((size_t)16 << 20)
#endif

/* ---------------------------------------------------------------- process-wide state */
static std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_forced_kernel{B200_K_AUTO};
static std::mutex g_mu;
static std::atomic<int> g_device{-1};   /* published last (release) by ensure_init; g_sm_count and g_init_pid are read after an acquire load of it */
static int g_sm_count = 0;
static pid_t g_init_pid = 0;

static thread_local char t_error[512] = "";
static thread_local const char *t_last_kernel = "";

void count_launch(const char *name) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  t_last_kernel = name;
}
int sm_count() { return g_sm_count; }

static int set_error(cudaError_t e, const char *what) {
  snprintf(t_error, sizeof t_error, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return (int)e;
}
#define CK(call)                                              \
  do {                                                        \
    cudaError_t e_ = (call);                                  \
    if (e_ != cudaSuccess) return set_error(e_, #call);       \
  } while (0)

/* One in-flight call's resources.  Contexts are pooled: at most as many exist as there were
 * concurrent callers. */
struct Context {
  cudaStream_t stream = nullptr;                /* compute (and small-problem copies) */
  cudaStream_t s_in = nullptr, s_out = nullptr; /* H2D / D2H streams of the pipelined host path */
  std::vector<cudaEvent_t> events;
  char *dws = nullptr;  size_t dws_bytes = 0;   /* device workspace */
  char *hws = nullptr;  size_t hws_bytes = 0;   /* pinned host staging */
  /* pinned slot rings for PAGEABLE operands of big problems (see h2d_any / d2h_any) */
  static const int kSlots = 3;
  char *in_slot[kSlots] = {nullptr, nullptr, nullptr};
  char *out_slot[kSlots] = {nullptr, nullptr, nullptr};
  size_t in_slot_bytes = 0, out_slot_bytes = 0;   /* grown on demand up to B200_SLOT_BYTES: a 1024^3 call pins 3 x 8 MB, not 6 x 32 MB */
  cudaEvent_t in_ev[kSlots] = {nullptr, nullptr, nullptr}, out_ev[kSlots] = {nullptr, nullptr, nullptr};
  cudaEvent_t order_ev = nullptr;                 /* orders the compute stream after the caller's legacy-stream work */
  bool in_busy[kSlots] = {false, false, false};
  int in_next = 0, out_next = 0;
  struct PendingOut { bool active = false; char *host; size_t hpitch, width, cols; };
  PendingOut out_pending[kSlots];
};
static std::vector<Context *> g_free_ctx;

static int ensure_init() {
  if (g_device.load(std::memory_order_acquire) >= 0) {
    if (getpid() != g_init_pid) {
      snprintf(t_error, sizeof t_error,
               "the CUDA context was created in parent process %d and cannot be used after fork() "
               "in process %d; call GEMM first in the child, or fork before the first call",
               (int)g_init_pid, (int)getpid());
      return (int)cudaErrorInitializationError;
    }
    return 0;
  }
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_device.load(std::memory_order_acquire) >= 0) return 0;
  int dev = 0, count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    snprintf(t_error, sizeof t_error,
             "no usable CUDA device (%s); this library has no CPU fallback",
             e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    return (int)(e != cudaSuccess ? e : cudaErrorNoDevice);
  }
  CK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    snprintf(t_error, sizeof t_error, "device %d is sm_%d%d; this library is built for sm_100a only",
             dev, prop.major, prop.minor);
    return (int)cudaErrorNoKernelImageForDevice;
  }
  g_sm_count = prop.multiProcessorCount;
  g_init_pid = getpid();
  g_device.store(dev, std::memory_order_release);
  return 0;
}

static thread_local bool t_foreign = false;   /* classify() met device memory this library cannot reach */

/* A lease binds the calling thread to the library's device for the duration of one call (the reference
 * has no such notion; a caller thread's current device may differ from the one the first call bound the
 * library to) and returns the context to the pool afterwards -- scrubbed if the call failed half-way, so
 * a pooled context never carries pending copy-outs into someone else's call. */
struct ContextLease {
  Context *c = nullptr;
  int prev_device = -1;
  bool failed = false;
  ~ContextLease();
};

static int acquire(ContextLease *lease) {
  int err = ensure_init();
  if (err) return err;
  const int dev = g_device.load(std::memory_order_acquire);
  int cur = dev;
  CK(cudaGetDevice(&cur));
  if (cur != dev) { CK(cudaSetDevice(dev)); lease->prev_device = cur; }
  t_foreign = false;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_free_ctx.empty()) { lease->c = g_free_ctx.back(); g_free_ctx.pop_back(); return 0; }
  }
  Context *c = new Context();
  cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete c; return set_error(e, "cudaStreamCreateWithFlags"); }
  lease->c = c;
  return 0;
}
static void scrub(Context *c) {
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->s_in) cudaStreamSynchronize(c->s_in);
  if (c->s_out) cudaStreamSynchronize(c->s_out);
  cudaGetLastError();
  for (int k = 0; k < Context::kSlots; k++) { c->out_pending[k].active = false; c->in_busy[k] = false; }
}
static void release(Context *c) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_free_ctx.push_back(c);
}
ContextLease::~ContextLease() {
  if (c) { if (failed) scrub(c); release(c); }
  if (prev_device >= 0) cudaSetDevice(prev_device);
}

/* The BLAS call is synchronous and works on "the operands as of now".  A device-resident operand may
 * still be being produced by work the caller queued earlier; our streams are non-blocking, so nothing
 * orders us after it implicitly.  An event recorded on the LEGACY default stream completes after
 * everything queued so far on it and on every blocking stream; the compute stream waits for that
 * event on the device.  (Round 1 called cudaDeviceSynchronize() here, which also stalled the host until
 * every other caller's stream had drained.)  Work queued on the caller's own NON-blocking streams is
 * the caller's to synchronise -- or use b200_gemm_async on that stream. */
static int order_after_caller(Context *c) {
  if (!c->order_ev) CK(cudaEventCreateWithFlags(&c->order_ev, cudaEventDisableTiming));
  CK(cudaEventRecord(c->order_ev, cudaStreamLegacy));
  CK(cudaStreamWaitEvent(c->stream, c->order_ev, 0));
  return 0;
}

static int reserve_device(Context *c, size_t bytes) {
  if (bytes <= c->dws_bytes) return 0;
  if (c->dws) { CK(cudaStreamSynchronize(c->stream)); CK(cudaFree(c->dws)); c->dws = nullptr; c->dws_bytes = 0; }
  size_t want = bytes + bytes / 8 + (1u << 20);
  CK(cudaMalloc((void **)&c->dws, want));
  c->dws_bytes = want;
  return 0;
}
static int reserve_pinned(Context *c, size_t bytes) {
  if (bytes <= c->hws_bytes) return 0;
  if (c->hws) { CK(cudaStreamSynchronize(c->stream)); CK(cudaFreeHost(c->hws)); c->hws = nullptr; c->hws_bytes = 0; }
  size_t want = bytes + bytes / 8 + (1u << 20);
  CK(cudaHostAlloc((void **)&c->hws, want, cudaHostAllocDefault));
  c->hws_bytes = want;
  return 0;
}

/* ------------------------------------------------------------------- kernel dispatch */
static void read_scalars(const b200_problem *p, DeviceGemm &g) {
  g.alpha_im = g.beta_im = 0.0;
  switch (p->dtype) {
    case B200_D:
      g.alpha_re = *(const double *)p->alpha; g.beta_re = *(const double *)p->beta; break;
    case B200_Z:
      g.alpha_re = ((const double *)p->alpha)[0]; g.alpha_im = ((const double *)p->alpha)[1];
      g.beta_re = ((const double *)p->beta)[0];   g.beta_im = ((const double *)p->beta)[1]; break;
    case B200_C:
      g.alpha_re = ((const float *)p->alpha)[0]; g.alpha_im = ((const float *)p->alpha)[1];
      g.beta_re = ((const float *)p->beta)[0];   g.beta_im = ((const float *)p->beta)[1]; break;
    default:
      g.alpha_re = *(const float *)p->alpha; g.beta_re = *(const float *)p->beta; break;
  }
}

/* Chooses the kernel family.  AUTO: the roofline kernels take every problem they support
 * above a small-size threshold; everything else goes to the generic kernel. */
static cudaError_t gemm3m_on_device(const DeviceGemm &g, cudaStream_t s);

static cudaError_t dispatch(const DeviceGemm &g, cudaStream_t stream) {
  const bool product = g.k > 0 && !(g.alpha_re == 0.0 && g.alpha_im == 0.0);
  if (!product && g.beta_re == 1.0 && g.beta_im == 0.0) return cudaSuccess; /* C unchanged */
  if (g.algo3m && product && !g.tri) {
    cudaError_t e3 = gemm3m_on_device(g, stream);
    if (e3 != cudaErrorNotSupported) return e3;             /* too small, too big for the workspace, or not complex: 4 multiplies */
  }
  int forced = g_forced_kernel.load(std::memory_order_relaxed);
  bool try_fast = product && forced != B200_K_GENERIC;
  if (forced == B200_K_AUTO) {
    /* below ~64^3 the launch dominates and the 32x32-tile kernel has the lower latency */
    double mnk = (double)g.m * (double)g.n * (double)g.k;
    if (mnk < 64.0 * 64.0 * 64.0 || g.m < 16 || g.n < 16) try_fast = false;
  }
  if (try_fast) {
    cudaError_t e = cudaErrorNotSupported;
    switch (g.dtype) {
      case B200_D:  e = launch_dgemm_dmma(g, stream); break;
      case B200_Z:  e = launch_zgemm_dmma(g, stream); break;
      case B200_S:  e = launch_sgemm_ffma(g, stream); break;
      case B200_C:  e = launch_cgemm_ffma(g, stream); break;
      case B200_SB: e = launch_sbgemm_tcgen05(g, stream); break;
    }
    if (e != cudaErrorNotSupported) return e;
    if (forced == B200_K_FAST) return e; /* caller insisted: report "not supported" */
  }
  if (g.tri) return cudaErrorNotSupported;   /* the generic kernel has no triangle mask: the caller uses block columns */
  return launch_generic(g, stream);
}

/* ------------------------------------------------------------------ pointer handling */
enum PtrKind { PTR_DEVICE, PTR_PINNED, PTR_PAGEABLE };
static PtrKind classify(const void *p) {
  cudaPointerAttributes at;
  cudaError_t e = cudaPointerGetAttributes(&at, p);
  if (e != cudaSuccess) { cudaGetLastError(); return PTR_PAGEABLE; }
  if (at.type == cudaMemoryTypeDevice) {
    const int dev = g_device.load(std::memory_order_acquire);
    if (at.device != dev) {
      /* memory of another GPU: usable in place only over peer access (NVLink) */
      cudaError_t pe = cudaDeviceEnablePeerAccess(at.device, 0);
      if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
        snprintf(t_error, sizeof t_error, "operand %p lives on device %d, the library is bound to device %d and peer access is unavailable (%s)",
                 p, at.device, dev, cudaGetErrorName(pe));
        t_foreign = true;
      } else {
        cudaGetLastError();
      }
    }
    return PTR_DEVICE;
  }
  if (at.type == cudaMemoryTypeManaged) return PTR_DEVICE;
  if (at.type == cudaMemoryTypeHost) return PTR_PINNED;
  return PTR_PAGEABLE;
}

static inline size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

/* ---- GEMM3M: three real products instead of four ------------------------------------------------------------------
 * driver/level3/gemm3m_level3.c computes a complex product from three real ones over specially packed panels
 * (kernel/generic/zgemm3m_*copy*.c add or subtract the real and imaginary parts while packing).  The same identity on
 * whole matrices, with X = op(A), Y = op(B) (conjugation folded into the sign of the imaginary plane):
 *     T1 = Xr Yr,  T2 = Xi Yi,  T3 = (Xr + Xi)(Yr + Yi);   Re(XY) = T1 - T2,  Im(XY) = T3 - T1 - T2.
 * split3 writes the three real planes of each operand once (HBM-bound, O(mk + kn)), the three products are ordinary
 * real GEMMs on the roofline kernels (op() becomes the real kernel's N / T), combine3 applies alpha and beta while it
 * forms C.  6 mnk real flops instead of 8: ZGEMM3M 8192^3 in about 3/4 of ZGEMM's time.  Workspace (3 planes per
 * operand and per product) comes from the stream-ordered pool; when the whole product would need more than
 * B200_3M_WORKSPACE_LIMIT, k is cut into chunks; a product whose C planes alone exceed the limit, or with an extent below
 * B200_3M_MIN, stays on the 4-multiply kernel. */
#ifdef B200_HOSTSIM
#define B200_3M_MIN 8
#else
#define B200_3M_MIN 512
#endif
#define B200_3M_WORKSPACE_LIMIT ((size_t)24 << 30)
#ifdef B200_HOSTSIM
#define B200_3M_CHUNK_MIN 8
#else
#define B200_3M_CHUNK_MIN 1024
#endif
static cudaError_t gemm3m_on_device(const DeviceGemm &g, cudaStream_t s) {
  if (g.dtype != B200_C && g.dtype != B200_Z) return cudaErrorNotSupported;
  static const int64_t min_extent = getenv("B200_3M_MIN") ? atol(getenv("B200_3M_MIN")) : B200_3M_MIN;   /* 0 or less: never */
  if (min_extent <= 0 || g.m < min_extent || g.n < min_extent || g.k < min_extent) return cudaErrorNotSupported;
  const bool dbl = g.dtype == B200_Z;
  const size_t rs = dbl ? 8 : 4;
  /* split3 / combine3 read and write whole (re, im) pairs: operands that are only aligned to their real type take the
   * 4-multiply path, which has its own rules for them */
  if ((((uintptr_t)g.a | (uintptr_t)g.b | (uintptr_t)g.c) & (2 * rs - 1)) != 0) return cudaErrorNotSupported;
  /* k goes through in chunks of kc -- all of it when the workspace allows, else halved until it does (the tall-skinny
   * 65536 x 256 x 65536 shape needs 100 GB of planes in one piece, 13 GB in chunks of 8192): each chunk is split,
   * multiplied and combined into C, the first with the caller's beta, the others with beta = 1 */
  const char *lim_env = getenv("B200_3M_WORKSPACE_BYTES");
  const size_t limit = lim_env ? (size_t)atoll(lim_env) : B200_3M_WORKSPACE_LIMIT;
  auto pitch = [&](int64_t rows) { return (int64_t)(round_up((size_t)rows * rs, 128) / rs); };
  const int64_t lpt = pitch(g.m);
  const size_t plane_t = round_up((size_t)lpt * (size_t)g.n * rs, 256);
  int64_t kc = g.k, lpa = 0, lpb = 0;
  size_t plane_a = 0, plane_b = 0, total = 0;
  for (;;) {
    lpa = pitch((g.transa & 1) ? kc : g.m); lpb = pitch((g.transb & 1) ? g.n : kc);
    plane_a = round_up((size_t)lpa * (size_t)((g.transa & 1) ? g.m : kc) * rs, 256);
    plane_b = round_up((size_t)lpb * (size_t)((g.transb & 1) ? kc : g.n) * rs, 256);
    total = 3 * (plane_a + plane_b + plane_t);
    if (total <= limit) break;
    if (kc <= B200_3M_CHUNK_MIN) return cudaErrorNotSupported;
    kc = ((kc / 2 + B200_3M_CHUNK_MIN - 1) / B200_3M_CHUNK_MIN) * B200_3M_CHUNK_MIN;
  }
#ifndef B200_HOSTSIM
  /* keep up to 8 GiB of freed workspace in the stream-ordered pool between calls (by default the pool hands
   * everything back at the next synchronisation and the next call maps it again); b200_shutdown trims it */
  static std::once_flag pool_once;
  std::call_once(pool_once, [] {
    cudaMemPool_t pool; int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t keep = (uint64_t)8 << 30;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
  });
#endif
  char *ws = nullptr;
  if (cudaMallocAsync((void **)&ws, total, s) != cudaSuccess) { cudaGetLastError(); return cudaErrorNotSupported; }
  char *pa[3], *pb[3], *pt[3];
  for (int i = 0; i < 3; i++) { pa[i] = ws + i * plane_a; pb[i] = ws + 3 * plane_a + i * plane_b; pt[i] = ws + 3 * (plane_a + plane_b) + i * plane_t; }
  const size_t es = 2 * rs;
  cudaError_t e = cudaSuccess;
  DeviceGemm r;
  r.dtype = dbl ? B200_D : B200_S; r.transa = g.transa & 1; r.transb = g.transb & 1;
  r.m = g.m; r.n = g.n; r.lda = lpa; r.ldb = lpb; r.ldc = lpt;
  r.alpha_re = 1.0; r.alpha_im = 0.0; r.beta_re = 0.0; r.beta_im = 0.0;
  for (int64_t k0 = 0; k0 < g.k && e == cudaSuccess; k0 += kc) {
    const int64_t kk = g.k - k0 < kc ? g.k - k0 : kc;
    /* columns k0.. of a stored m x k operand, rows k0.. of a stored k x m one */
    const char *a0 = (const char *)g.a + ((g.transa & 1) ? (size_t)k0 : (size_t)k0 * (size_t)g.lda) * es;
    const char *b0 = (const char *)g.b + ((g.transb & 1) ? (size_t)k0 * (size_t)g.ldb : (size_t)k0) * es;
    e = launch_split3(g.dtype, (g.transa & 1) ? kk : g.m, (g.transa & 1) ? g.m : kk, a0, g.lda, (g.transa & 2) != 0, pa[0], pa[1], pa[2], lpa, s);
    if (e == cudaSuccess) e = launch_split3(g.dtype, (g.transb & 1) ? g.n : kk, (g.transb & 1) ? kk : g.n, b0, g.ldb, (g.transb & 2) != 0, pb[0], pb[1], pb[2], lpb, s);
    r.k = kk;
    for (int i = 0; i < 3 && e == cudaSuccess; i++) {
      r.a = pa[i]; r.b = pb[i]; r.c = pt[i];
      e = dispatch(r, s);
    }
    if (e == cudaSuccess)
      e = launch_combine3(g.dtype, g.m, g.n, pt[0], pt[1], pt[2], lpt, g.alpha_re, g.alpha_im, k0 == 0 ? g.beta_re : 1.0, k0 == 0 ? g.beta_im : 0.0, g.c, g.ldc, s);
  }
  cudaFreeAsync(ws, s);
  return e;
}

/* staged copy descriptor for one operand */
struct Operand {
  const char *host = nullptr;    /* user pointer (host side) or nullptr when already device */
  char *dev = nullptr;           /* device address the kernel will use */
  int64_t rows = 0, cols = 0;    /* stored extent, column-major */
  int64_t ld_user = 0, ld_dev = 0;
  size_t es = 0;
  PtrKind kind = PTR_DEVICE;
  size_t bytes_dev() const { return (size_t)ld_dev * (size_t)cols * es; }
};

/* Small problems (the 27 783 x 2 ctest calls per precision are all <= 35^3): pack every
 * host operand into ONE pinned block, one H2D, one kernel, one D2H, then copy only the
 * m x n window of C back so padding rows of the caller's C stay bit-identical
 * (ctest LDERES, c_dblat3.f:2352-2412). */
static const size_t kSmallBytes = B200_SMALL_BYTES;

static void pack_to(char *dst, const Operand &o) {
  size_t row_bytes = (size_t)o.rows * o.es;
  for (int64_t j = 0; j < o.cols; j++)
    memcpy(dst + (size_t)j * (size_t)o.ld_dev * o.es, o.host + (size_t)j * (size_t)o.ld_user * o.es, row_bytes);
}

/* ---------------------------------------------------------------- pageable host operands
 * BLAS callers hand us malloc'd memory.  A cudaMemcpy from pageable memory is staged by the driver
 * at a few GB/s on one thread; instead the columns are copied into pinned slots by a small pool of
 * host threads (the reference uses all cores for the GEMM itself; we only borrow a few for memcpy)
 * while the previous slot is in flight on the DMA engine. */
static const size_t kSlotBytes = B200_SLOT_BYTES;

class HostPool {
 public:
  static HostPool &get() {                     /* never destroyed; stop() joins the workers */
    static std::once_flag once;
    std::call_once(once, [] { instance() = new HostPool(); });
    return *instance();
  }
  static HostPool *&instance() { static HostPool *p = nullptr; return p; }
  /* run fn(i) for i in [0, n) on the pool plus the calling thread */
  void parallel_for(int64_t n, const std::function<void(int64_t)> &fn) {
    if (n <= 0) return;
    if (n == 1 || workers_.empty()) { for (int64_t i = 0; i < n; i++) fn(i); return; }
    std::unique_lock<std::mutex> lk(mu_);
    while (busy_) done_cv_.wait(lk);           /* one job at a time; concurrent callers queue up */
    busy_ = true; fn_ = &fn; n_ = n; next_ = 0; left_ = n;
    gen_.fetch_add(1, std::memory_order_release);
    const bool wake = sleepers_ > 0;
    lk.unlock();
    if (wake) work_cv_.notify_all();
    run_some();
    lk.lock();
    while (left_ > 0) done_cv_.wait(lk);
    busy_ = false; fn_ = nullptr;
    lk.unlock();
    done_cv_.notify_all();
  }
  /* workers still spinning after a recent job pick the next one up in microseconds; sleeping ones take up
   * to milliseconds to wake on a virtualised host */
  bool hot() { std::lock_guard<std::mutex> lk(mu_); return !workers_.empty() && sleepers_ == 0 && gen_.load() > 0; }
  /* wake sleeping workers without giving them work: they spin for the next millisecond, so the copies that
   * follow (the next chunk of this call, the next call of a loop) find them hot */
  void poke() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      if (workers_.empty() || busy_ || sleepers_ == 0) return;
      gen_.fetch_add(1, std::memory_order_release);
    }
    work_cv_.notify_all();
  }
  void stop() {
    { std::lock_guard<std::mutex> lk(mu_); stop_.store(true); }
    work_cv_.notify_all();
    for (std::thread &t : workers_) if (t.joinable()) t.join();
    workers_.clear();
  }
 private:
  HostPool() {
    unsigned hc = std::thread::hardware_concurrency();
    int n = (int)(hc / 2); if (n > 7) n = 7; if (n < 1) n = 0;
    for (int i = 0; i < n; i++) workers_.emplace_back([this] { loop(); });
  }
  void run_some() {
    for (;;) {
      int64_t i;
      { std::lock_guard<std::mutex> lk(mu_); if (!fn_ || next_ >= n_) return; i = next_++; }
      (*fn_)(i);
      { std::lock_guard<std::mutex> lk(mu_); if (--left_ == 0) done_cv_.notify_all(); }
    }
  }
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      /* spin for a millisecond after the last job (a benchmark loop or a solver issues the next call
       * right away), then sleep */
      const auto t0 = std::chrono::steady_clock::now();
      bool got = false;
      while (std::chrono::steady_clock::now() - t0 < std::chrono::microseconds(1000)) {
        if (gen_.load(std::memory_order_acquire) != seen || stop_.load(std::memory_order_relaxed)) { got = true; break; }
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
      }
      if (!got) {
        std::unique_lock<std::mutex> lk(mu_);
        sleepers_++;
        while (gen_.load(std::memory_order_acquire) == seen && !stop_.load()) work_cv_.wait(lk);
        sleepers_--;
      }
      if (stop_.load()) return;
      seen = gen_.load(std::memory_order_acquire);
      run_some();
    }
  }
  std::mutex mu_; std::condition_variable work_cv_, done_cv_;
  std::vector<std::thread> workers_;
  const std::function<void(int64_t)> *fn_ = nullptr;
  int64_t n_ = 0, next_ = 0, left_ = 0; bool busy_ = false;
  std::atomic<uint64_t> gen_{0};
  std::atomic<bool> stop_{false};
  int sleepers_ = 0;
};

/* columns [0, cols) of `width` bytes each, pitches in bytes */
static void host_copy_cols(char *dst, size_t dpitch, const char *src, size_t spitch, size_t width, size_t cols) {
  const size_t total = width * cols;
  /* waking sleeping pool threads costs up to milliseconds on a virtualised host (measured: a 1024^3
   * DGEMM call went from 2 ms to 16 ms when its 8 MB operands were split over the pool), so only
   * copies of at least 16 MB are shared out */
  /* spinning workers (a call finished less than a millisecond ago) take shares of 512 KB; otherwise 2 MB shares
   * from 16 MB up */
  const bool hot = HostPool::get().hot();
  if (!hot && total >= B200_POOL_BYTES / 16) HostPool::get().poke();
  const size_t share_at = hot ? B200_POOL_BYTES / 16 : B200_POOL_BYTES, share = hot ? B200_POOL_BYTES / 32 : B200_POOL_BYTES / 8;
  int64_t parts = total >= share_at ? (int64_t)(total / share) : 1;
  if (parts > 16) parts = 16;
  if ((size_t)parts > cols) parts = (int64_t)cols;
  if (parts < 1) parts = 1;
  HostPool::get().parallel_for(parts, [&](int64_t p) {
    size_t c0 = cols * (size_t)p / (size_t)parts, c1 = cols * (size_t)(p + 1) / (size_t)parts;
    if (dpitch == width && spitch == width) { memcpy(dst + c0 * width, src + c0 * width, (c1 - c0) * width); return; }
    for (size_t c = c0; c < c1; c++) memcpy(dst + c * dpitch, src + c * spitch, width);
  });
}

static int drain_all_out(Context *ctx);

/* in / out slot rings, grown to the size this transfer needs (at most B200_SLOT_BYTES each) */
static int ensure_slots(Context *ctx, bool out, size_t transfer_bytes) {
  size_t want = round_up(transfer_bytes < (size_t)1 ? 1 : transfer_bytes, (size_t)1 << 20);
  if (want > kSlotBytes) want = kSlotBytes;
#ifdef B200_HOSTSIM
  want = kSlotBytes;
#endif
  size_t &have = out ? ctx->out_slot_bytes : ctx->in_slot_bytes;
  if (have >= want) return 0;
  char **slot = out ? ctx->out_slot : ctx->in_slot;
  cudaEvent_t *ev = out ? ctx->out_ev : ctx->in_ev;
  if (have) {                                    /* growing: nothing may still be in flight through the old slots */
    if (ctx->stream) CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->s_in) CK(cudaStreamSynchronize(ctx->s_in));
    if (ctx->s_out) CK(cudaStreamSynchronize(ctx->s_out));
    if (out) { int e = drain_all_out(ctx); if (e) return e; }
    for (int i = 0; i < Context::kSlots; i++) { CK(cudaFreeHost(slot[i])); slot[i] = nullptr; if (!out) ctx->in_busy[i] = false; }
    have = 0;
  }
  for (int i = 0; i < Context::kSlots; i++) {
    CK(cudaHostAlloc((void **)&slot[i], want, cudaHostAllocDefault));
    if (!ev[i]) CK(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
  }
  have = want;
  return 0;
}

static int drain_out(Context *ctx, int k) {
  Context::PendingOut &po = ctx->out_pending[k];
  if (!po.active) return 0;
  CK(cudaEventSynchronize(ctx->out_ev[k]));
  host_copy_cols(po.host, po.hpitch, ctx->out_slot[k], po.width, po.width, po.cols);
  po.active = false;
  return 0;
}
static int drain_all_out(Context *ctx) {
  for (int k = 0; k < Context::kSlots; k++) { int e = drain_out(ctx, k); if (e) return e; }
  return 0;
}

/* host (pinned or pageable) -> device, `cols` columns of `width` bytes, enqueued on s */
static int h2d_any(Context *ctx, cudaStream_t s, PtrKind kind, char *dev, size_t dpitch, const char *host, size_t hpitch,
                   size_t width, size_t cols) {
  if (kind != PTR_PAGEABLE || width == 0 || cols == 0) {
    CK(cudaMemcpy2DAsync(dev, dpitch, host, hpitch, width, cols, cudaMemcpyHostToDevice, s));
    return 0;
  }
  if (kSlotBytes / width == 0) {                    /* a single column larger than a slot: let the driver stage it */
    CK(cudaMemcpy2DAsync(dev, dpitch, host, hpitch, width, cols, cudaMemcpyHostToDevice, s));
    return 0;
  }
  /* slots of a third of the transfer: the pool fills slot i + 1 while the DMA engine drains slot i */
  int err = ensure_slots(ctx, false, width * cols / 3 > width ? width * cols / 3 : width);
  if (err) return err;
  const size_t cpc = ctx->in_slot_bytes / width;    /* columns per chunk */
  for (size_t c0 = 0; c0 < cols; c0 += cpc) {
    const size_t nc = cols - c0 < cpc ? cols - c0 : cpc;
    const int k = ctx->in_next; ctx->in_next = (k + 1) % Context::kSlots;
    if (ctx->in_busy[k]) CK(cudaEventSynchronize(ctx->in_ev[k]));
    host_copy_cols(ctx->in_slot[k], width, host + c0 * hpitch, hpitch, width, nc);
    CK(cudaMemcpy2DAsync(dev + c0 * dpitch, dpitch, ctx->in_slot[k], width, width, nc, cudaMemcpyHostToDevice, s));
    CK(cudaEventRecord(ctx->in_ev[k], s));
    ctx->in_busy[k] = true;
  }
  return 0;
}

/* device -> host; for pageable memory the copy-out of a slot is deferred (drain_out) */
static int d2h_any(Context *ctx, cudaStream_t s, PtrKind kind, char *host, size_t hpitch, const char *dev, size_t dpitch,
                   size_t width, size_t cols) {
  if (kind != PTR_PAGEABLE || width == 0 || cols == 0 || kSlotBytes / width == 0) {
    CK(cudaMemcpy2DAsync(host, hpitch, dev, dpitch, width, cols, cudaMemcpyDeviceToHost, s));
    return 0;
  }
  int err = ensure_slots(ctx, true, width * cols / 3 > width ? width * cols / 3 : width);
  if (err) return err;
  const size_t cpc = ctx->out_slot_bytes / width;
  for (size_t c0 = 0; c0 < cols; c0 += cpc) {
    const size_t nc = cols - c0 < cpc ? cols - c0 : cpc;
    const int k = ctx->out_next; ctx->out_next = (k + 1) % Context::kSlots;
    if ((err = drain_out(ctx, k))) return err;
    CK(cudaMemcpy2DAsync(ctx->out_slot[k], width, dev + c0 * dpitch, dpitch, width, nc, cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(ctx->out_ev[k], s));
    Context::PendingOut &po = ctx->out_pending[k];
    po.active = true; po.host = host + c0 * hpitch; po.hpitch = hpitch; po.width = width; po.cols = nc;
  }
  return 0;
}

/* Pipelined host path for big problems: C is cut into kPanel x kPanel blocks; row panels of
 * op(A) and column panels of op(B) are uploaded alternately on the H2D stream, every block whose two
 * panels have landed is multiplied on the compute stream, and finished blocks are downloaded on the
 * D2H stream -- so after the first pair of panels the PCIe transfers in both directions hide
 * behind the GEMMs (the reference, being a CPU library, has no such phase; this is what makes the
 * host-pointer BLAS call approach the device-resident rate). */
static const int64_t kPanelMin = B200_PANEL_MIN;   /* smallest block edge; see panel_edge() */

/* Block edge of the pipeline: 2048 gives the earliest start (first A and B panels are small) but a
 * 2048 x 2048 block is only 256 C tiles = 1.7 waves of 148 CTAs (86 % efficient); 4096 gives 1024 tiles
 * = 6.9 waves.  Use the larger block once the problem has enough of them to keep the pipe busy. */
static int64_t panel_edge(int64_t m, int64_t n) { return (m >= B200_PANEL_BIG_AT && n >= B200_PANEL_BIG_AT) ? B200_PANEL_BIG : B200_PANEL_MIN; }

static int event_at(Context *ctx, size_t i, cudaEvent_t *out) {
  while (ctx->events.size() <= i) {
    cudaEvent_t e;
    CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->events.push_back(e);
  }
  *out = ctx->events[i];
  return 0;
}

static int run_pipelined(Context *ctx, const b200_problem *p, const DeviceGemm &g, const Operand &A,
                         const Operand &B, const Operand &C, bool use_beta) {
  if (!ctx->s_in) {
    CK(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
  }
  const int64_t kPanel = panel_edge(p->m, p->n);
  const int64_t pm = (p->m + kPanel - 1) / kPanel, pn = (p->n + kPanel - 1) / kPanel;
  const size_t ies = A.es, oes = C.es;
  const bool ta = p->transa & 1, tb = p->transb & 1;
  cudaStream_t sc = ctx->stream, si = ctx->s_in, so = ctx->s_out;
  size_t ev = 0;
  std::vector<cudaEvent_t> ev_a(pm, nullptr), ev_b(pn, nullptr);
  int err;

  auto upload_a = [&](int64_t i) -> int {
    if (A.kind == PTR_DEVICE) return 0;
    const int64_t i0 = i * kPanel, mb = (p->m - i0 < kPanel) ? p->m - i0 : kPanel;
    if (!ta) err = h2d_any(ctx, si, A.kind, A.dev + i0 * ies, A.ld_dev * ies, A.host + i0 * ies, A.ld_user * ies, mb * ies, p->k);
    else     err = h2d_any(ctx, si, A.kind, A.dev + i0 * A.ld_dev * ies, A.ld_dev * ies, A.host + i0 * A.ld_user * ies, A.ld_user * ies, p->k * ies, mb);
    if (err) return err;
    if ((err = event_at(ctx, ev++, &ev_a[i]))) return err;
    CK(cudaEventRecord(ev_a[i], si));
    return 0;
  };
  auto upload_b = [&](int64_t j) -> int {
    if (B.kind == PTR_DEVICE) return 0;
    const int64_t j0 = j * kPanel, nb = (p->n - j0 < kPanel) ? p->n - j0 : kPanel;
    if (!tb) err = h2d_any(ctx, si, B.kind, B.dev + j0 * B.ld_dev * ies, B.ld_dev * ies, B.host + j0 * B.ld_user * ies, B.ld_user * ies, p->k * ies, nb);
    else     err = h2d_any(ctx, si, B.kind, B.dev + j0 * ies, B.ld_dev * ies, B.host + j0 * ies, B.ld_user * ies, nb * ies, p->k);
    if (err) return err;
    if ((err = event_at(ctx, ev++, &ev_b[j]))) return err;
    CK(cudaEventRecord(ev_b[j], si));
    return 0;
  };
  auto block = [&](int64_t i, int64_t j) -> int {
    const int64_t i0 = i * kPanel, j0 = j * kPanel;
    const int64_t mb = (p->m - i0 < kPanel) ? p->m - i0 : kPanel, nb = (p->n - j0 < kPanel) ? p->n - j0 : kPanel;
    DeviceGemm b = g;
    b.m = mb; b.n = nb;
    b.a = A.dev + (ta ? i0 * A.ld_dev : i0) * ies;
    b.b = B.dev + (tb ? j0 : j0 * B.ld_dev) * ies;
    b.c = C.dev + (i0 + j0 * C.ld_dev) * oes;
    if (C.kind != PTR_DEVICE && use_beta) {
      cudaEvent_t e;
      if ((err = h2d_any(ctx, si, C.kind, (char *)b.c, C.ld_dev * oes, C.host + (i0 + j0 * C.ld_user) * oes, C.ld_user * oes, mb * oes, nb))) return err;
      if ((err = event_at(ctx, ev++, &e))) return err;
      CK(cudaEventRecord(e, si));
      CK(cudaStreamWaitEvent(sc, e, 0));
    }
    if (ev_a[i]) CK(cudaStreamWaitEvent(sc, ev_a[i], 0));
    if (ev_b[j]) CK(cudaStreamWaitEvent(sc, ev_b[j], 0));
    CK(dispatch(b, sc));
    if (C.kind != PTR_DEVICE) {
      cudaEvent_t e;
      if ((err = event_at(ctx, ev++, &e))) return err;
      CK(cudaEventRecord(e, sc));
      CK(cudaStreamWaitEvent(so, e, 0));
      if ((err = d2h_any(ctx, so, C.kind, (char *)p->c + (i0 + j0 * C.ld_user) * oes, C.ld_user * oes, (const char *)b.c, C.ld_dev * oes, mb * oes, nb))) return err;
    }
    return 0;
  };

  const int64_t steps = pm > pn ? pm : pn;
  for (int64_t t = 0; t < steps; t++) {
    if (t < pm && (err = upload_a(t))) return err;
    if (t < pn && (err = upload_b(t))) return err;
    if (t < pm) for (int64_t j = 0; j <= t && j < pn; j++) if ((err = block(t, j))) return err;   /* new row    */
    if (t < pn) for (int64_t i = 0; i < t && i < pm; i++) if ((err = block(i, t))) return err;    /* new column */
  }
  CK(cudaStreamSynchronize(si));
  CK(cudaStreamSynchronize(sc));
  if ((err = drain_all_out(ctx))) return err;
  CK(cudaStreamSynchronize(so));
  return 0;
}

static int run_on_context(Context *ctx, const b200_problem *p) {
  DeviceGemm g;
  g.dtype = p->dtype; g.transa = p->transa; g.transb = p->transb;
  g.m = p->m; g.n = p->n; g.k = p->k;
  g.algo3m = p->algo3m;
  read_scalars(p, g);
  const bool product = g.k > 0 && !(g.alpha_re == 0.0 && g.alpha_im == 0.0);
  const bool use_beta = !(g.beta_re == 0.0 && g.beta_im == 0.0);
  if (!product && g.beta_re == 1.0 && g.beta_im == 0.0) return 0;

  Operand A, B, C;
  A.es = B.es = b200_in_size(p->dtype); C.es = b200_out_size(p->dtype);
  A.rows = (p->transa & 1) ? p->k : p->m; A.cols = (p->transa & 1) ? p->m : p->k;
  B.rows = (p->transb & 1) ? p->n : p->k; B.cols = (p->transb & 1) ? p->k : p->n;
  C.rows = p->m; C.cols = p->n;
  A.ld_user = p->lda; B.ld_user = p->ldb; C.ld_user = p->ldc;
  A.kind = product ? classify(p->a) : PTR_DEVICE;   /* A, B are never touched without a product */
  B.kind = product ? classify(p->b) : PTR_DEVICE;
  C.kind = classify(p->c);

  if (t_foreign) return (int)cudaErrorInvalidDevice;
  if ((product && (A.kind == PTR_DEVICE || B.kind == PTR_DEVICE)) || C.kind == PTR_DEVICE) {
    int e = order_after_caller(ctx);
    if (e) return e;
  }

  Operand *ops[3] = {&A, &B, &C};
  const void *user[3] = {p->a, p->b, p->c};
  size_t need = 0;
  for (int i = 0; i < 3; i++) {
    Operand &o = *ops[i];
    if (o.kind == PTR_DEVICE) { o.dev = (char *)user[i]; o.ld_dev = o.ld_user; continue; }
    o.host = (const char *)user[i];
    /* device copies get a leading dimension rounded to 128 bytes: keeps every column
     * 16-byte aligned for cp.async / TMA whatever the caller's ld was */
    o.ld_dev = (int64_t)(round_up((size_t)(o.rows > 0 ? o.rows : 1) * o.es, 128) / o.es);
    need += round_up(o.bytes_dev(), 256);
  }

  cudaStream_t s = ctx->stream;
  if (need == 0) {
    g.a = A.dev; g.b = B.dev; g.c = C.dev; g.lda = A.ld_dev; g.ldb = B.ld_dev; g.ldc = C.ld_dev;
    CK(dispatch(g, s));
    CK(cudaStreamSynchronize(s));
    return 0;
  }

  int err = reserve_device(ctx, need);
  if (err) return err;
  size_t off = 0;
  for (int i = 0; i < 3; i++) {
    Operand &o = *ops[i];
    if (o.kind == PTR_DEVICE) continue;
    o.dev = ctx->dws + off;
    off += round_up(o.bytes_dev(), 256);
  }
  g.a = A.dev; g.b = B.dev; g.c = C.dev; g.lda = A.ld_dev; g.ldb = B.ld_dev; g.ldc = C.ld_dev;

  if (need <= kSmallBytes) {
    err = reserve_pinned(ctx, need);
    if (err) return err;
    /* mirror the device layout in the pinned block so one copy moves everything; C is
     * uploaded only when beta will read it */
    size_t up_begin = (size_t)-1, up_end = 0;
    for (int i = 0; i < 3; i++) {
      Operand &o = *ops[i];
      if (o.kind == PTR_DEVICE) continue;
      if (i == 2 && !use_beta) continue;
      if (i < 2 && !product) continue;
      size_t o_off = (size_t)(o.dev - ctx->dws);
      pack_to(ctx->hws + o_off, o);
      if (o_off < up_begin) up_begin = o_off;
      if (o_off + o.bytes_dev() > up_end) up_end = o_off + o.bytes_dev();
    }
    if (up_end > up_begin)
      CK(cudaMemcpyAsync(ctx->dws + up_begin, ctx->hws + up_begin, up_end - up_begin, cudaMemcpyHostToDevice, s));
    CK(dispatch(g, s));
    if (C.kind != PTR_DEVICE) {
      size_t c_off = (size_t)(C.dev - ctx->dws);
      CK(cudaMemcpyAsync(ctx->hws + c_off, C.dev, C.bytes_dev(), cudaMemcpyDeviceToHost, s));
      CK(cudaStreamSynchronize(s));
      size_t row_bytes = (size_t)C.rows * C.es;
      char *uc = (char *)p->c;
      for (int64_t j = 0; j < C.cols; j++)
        memcpy(uc + (size_t)j * (size_t)C.ld_user * C.es, ctx->hws + c_off + (size_t)j * (size_t)C.ld_dev * C.es, row_bytes);
    } else {
      CK(cudaStreamSynchronize(s));
    }
    return 0;
  }

  if (product && p->m >= 2 * kPanelMin && p->n >= 2 * kPanelMin && need >= B200_PIPE_BYTES) return run_pipelined(ctx, p, g, A, B, C, use_beta);

  /* large host operands: strided DMA straight from / to the caller's memory (full PCIe
   * rate when it is pinned; staged by the driver when it is pageable) */
  static const bool trace = getenv("B200_TRACE") != nullptr;      /* B200_TRACE=1: per-phase host timestamps of this path on stderr */
  const auto t0 = std::chrono::steady_clock::now();
  auto since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
  double t_up = 0, t_launch = 0, t_down = 0, t_drain = 0;
  for (int i = 0; i < 3; i++) {
    Operand &o = *ops[i];
    if (o.kind == PTR_DEVICE) continue;
    if (i == 2 && !use_beta) continue;
    if (i < 2 && !product) continue;
    if ((err = h2d_any(ctx, s, o.kind, o.dev, (size_t)o.ld_dev * o.es, o.host, (size_t)o.ld_user * o.es,
                       (size_t)o.rows * o.es, (size_t)o.cols))) return err;
  }
  t_up = since();
  CK(dispatch(g, s));
  t_launch = since();
  if (C.kind != PTR_DEVICE) {
    if ((err = d2h_any(ctx, s, C.kind, (char *)p->c, (size_t)C.ld_user * C.es, C.dev, (size_t)C.ld_dev * C.es,
                       (size_t)C.rows * C.es, (size_t)C.cols))) return err;
    t_down = since();
    if ((err = drain_all_out(ctx))) return err;
    t_drain = since();
  }
  CK(cudaStreamSynchronize(s));
  if (trace)
    fprintf(stderr, "b200 trace %lldx%lldx%lld: uploads enqueued %.3f ms, kernel enqueued %.3f, download enqueued %.3f, drained %.3f, done %.3f\n",
            (long long)p->m, (long long)p->n, (long long)p->k, t_up, t_launch, t_down, t_drain, since());
  return 0;
}

#include "runtime_level3.inl"

/* summa.cu: one local product of the distributed driver, column-major NN on device pointers, enqueued on `stream` */
int summa_local_gemm(int dtype, int64_t m, int64_t n, int64_t k, const void *alpha, const void *a, int64_t lda, const void *b,
                     int64_t ldb, const void *beta, void *c, int64_t ldc, cudaStream_t stream) {
  if (m <= 0 || n <= 0) return 0;
  b200_problem p = {dtype, B200_N, B200_N, m, n, k, lda, ldb, ldc, alpha, beta, a, b, c};
  DeviceGemm g;
  g.dtype = dtype; g.transa = B200_N; g.transb = B200_N; g.m = m; g.n = n; g.k = k;
  g.lda = lda; g.ldb = ldb; g.ldc = ldc; g.a = a; g.b = b; g.c = c;
  read_scalars(&p, g);
  CK(dispatch(g, stream));
  return 0;
}
void summa_set_error(const char *msg) { snprintf(t_error, sizeof t_error, "%s", msg); }

}  // namespace b200

/* ------------------------------------------------------------------------ C ABI ---- */
using namespace b200;

extern "C" {

B200_HIDDEN int b200_run_problem(const b200_problem *p) {
  {  /* C := 1 * C + 0: nothing to do, return before CUDA is touched (the reference's drivers leave the
        same way: level3.c:229-259 scales only when beta != 1 and returns when alpha == 0 or k == 0) */
    DeviceGemm g;
    read_scalars(p, g);
    const bool product = p->k > 0 && !(g.alpha_re == 0.0 && g.alpha_im == 0.0);
    if (!product && g.beta_re == 1.0 && g.beta_im == 0.0) return 0;
  }
  ContextLease lease;
  int err = acquire(&lease);
  if (err) return err;
  t_error[0] = 0;
  err = run_on_context(lease.c, p);
  lease.failed = err != 0;
  return err;
}

B200_HIDDEN int b200_run_level3(const b200_l3_problem *p) {
  {  /* nothing to do (alpha == 0 or k == 0 with beta == 1): return before CUDA is touched */
    const bool symm = p->routine == B200_SYMM || p->routine == B200_HEMM;
    const bool trxm = p->routine == B200_TRMM || p->routine == B200_TRSM;
    const bool product = !(p->alpha[0] == 0.0 && p->alpha[1] == 0.0) && (symm || trxm || p->k > 0);
    if (!trxm && !product && p->beta[0] == 1.0 && p->beta[1] == 0.0) return 0;
  }
  ContextLease lease;
  int err = acquire(&lease);
  if (err) return err;
  t_error[0] = 0;
  err = run_level3_on_context(lease.c, p);
  lease.failed = err != 0;
  return err;
}

/* gemm_batch (interface/gemm_batch.c:322-366 hands one queue entry per matrix to the thread pool):
 * when every operand lives in host memory and the whole batch packs into 64 MB, all A and B
 * blocks go up in ONE copy, every matrix gets its own kernel launch on one stream, and all C
 * blocks come back in ONE copy -- one synchronisation per batch instead of one per matrix.  A batch
 * whose operands are all device memory is launched back to back and synchronised once.  Anything
 * else (mixed host / device operands, huge matrices) runs matrix by matrix. */
/* A batch of SMALL matrices (every m, n <= 128, one precision) is ONE launch of the grouped kernel
 * (gemm_generic.cu): the problem list and its tile prefix sums are built in pinned memory at `hdesc`,
 * copied to `ddesc`, and a 1-D grid walks all tiles of all matrices.  1000 x 64^3 DGEMMs: 50 ms as 1000
 * launches of the roofline kernel (one or four CTAs each), well under a millisecond grouped. */
static bool batch_is_groupable(const b200_problem *p, int64_t count) {
  static const bool enabled = !(getenv("B200_BATCH_GROUPED") && atoi(getenv("B200_BATCH_GROUPED")) == 0);
  if (!enabled || count < 2 || g_forced_kernel.load(std::memory_order_relaxed) == B200_K_FAST) return false;
  for (int64_t i = 0; i < count; i++)
    if (p[i].dtype != p[0].dtype || p[i].m > 128 || p[i].n > 128) return false;
  return true;
}
static size_t grouped_desc_bytes(size_t count) { return round_up(count * sizeof(DeviceGemm) + (count + 1) * sizeof(int64_t), 256); }
/* fills the descriptor block at hdesc (the caller uploads it to ddesc, before the launch, on stream s) */
static int64_t fill_batch_desc(const std::vector<DeviceGemm> &gs, char *hdesc) {
  const size_t count = gs.size();
  DeviceGemm *hp = (DeviceGemm *)hdesc;
  int64_t *ht = (int64_t *)(hdesc + count * sizeof(DeviceGemm));
  int64_t tiles = 0;
  for (size_t i = 0; i < count; i++) {
    const DeviceGemm &g = gs[i];
    memcpy(&hp[i], &g, sizeof g);
    ht[i] = tiles;
    const bool product = g.k > 0 && !(g.alpha_re == 0.0 && g.alpha_im == 0.0);
    if (product || !(g.beta_re == 1.0 && g.beta_im == 0.0)) tiles += generic_tile_count(g.m, g.n);   /* C := 1 * C: no tiles, C keeps its bits */
  }
  ht[count] = tiles;
  return tiles;
}
static cudaError_t launch_batch_grouped(int dtype, size_t count, int64_t tiles, const char *ddesc, cudaStream_t s) {
  return launch_grouped(dtype, (const DeviceGemm *)ddesc, (const int64_t *)(ddesc + count * sizeof(DeviceGemm)), (int)count, tiles, s);
}

static int run_batch_packed(Context *ctx, const b200_problem *p, int64_t count, bool *handled) {
  *handled = false;
  static const bool enabled = !(getenv("B200_BATCH_PACKED") && atoi(getenv("B200_BATCH_PACKED")) == 0);
  if (!enabled) return 0;
  struct Item { Operand A, B, C; DeviceGemm g; bool product, use_beta; };
  std::vector<Item> items((size_t)count);
  size_t ab_bytes = 0, c_bytes = 0;
  bool any_beta = false;

  /* every operand already on the device: no staging at all -- order after whatever the caller has in
   * flight, launch on one stream (one grouped launch when the matrices are small), synchronise once */
  {
    bool all_device = true;
    for (int64_t i = 0; i < count && all_device; i++) {
      DeviceGemm g;
      read_scalars(&p[i], g);
      const bool product = p[i].k > 0 && !(g.alpha_re == 0.0 && g.alpha_im == 0.0);
      all_device = classify(p[i].c) == PTR_DEVICE && (!product || (classify(p[i].a) == PTR_DEVICE && classify(p[i].b) == PTR_DEVICE));
    }
    if (t_foreign) return (int)cudaErrorInvalidDevice;
    if (all_device) {
      int oe = order_after_caller(ctx);
      if (oe) return oe;
      std::vector<DeviceGemm> gs((size_t)count);
      for (int64_t i = 0; i < count; i++) {
        const b200_problem &q = p[i];
        DeviceGemm &g = gs[(size_t)i];
        g.dtype = q.dtype; g.transa = q.transa; g.transb = q.transb; g.m = q.m; g.n = q.n; g.k = q.k;
        g.a = q.a; g.b = q.b; g.c = q.c; g.lda = q.lda; g.ldb = q.ldb; g.ldc = q.ldc;
        read_scalars(&q, g);
      }
      if (batch_is_groupable(p, count)) {
        const size_t db = grouped_desc_bytes((size_t)count);
        int err = reserve_device(ctx, db);
        if (err) return err;
        if ((err = reserve_pinned(ctx, db))) return err;
        const int64_t tiles = fill_batch_desc(gs, ctx->hws);
        CK(cudaMemcpyAsync(ctx->dws, ctx->hws, db, cudaMemcpyHostToDevice, ctx->stream));
        CK(launch_batch_grouped(gs[0].dtype, gs.size(), tiles, ctx->dws, ctx->stream));
      } else {
        for (const DeviceGemm &g : gs) CK(dispatch(g, ctx->stream));
      }
      CK(cudaStreamSynchronize(ctx->stream));
      *handled = true;
      return 0;
    }
  }

  for (int64_t i = 0; i < count; i++) {
    Item &it = items[(size_t)i];
    const b200_problem &q = p[i];
    DeviceGemm &g = it.g;
    g.dtype = q.dtype; g.transa = q.transa; g.transb = q.transb; g.m = q.m; g.n = q.n; g.k = q.k;
    read_scalars(&q, g);
    it.product = g.k > 0 && !(g.alpha_re == 0.0 && g.alpha_im == 0.0);
    it.use_beta = !(g.beta_re == 0.0 && g.beta_im == 0.0);
    any_beta |= it.use_beta;
    Operand &A = it.A, &B = it.B, &C = it.C;
    A.es = B.es = b200_in_size(q.dtype); C.es = b200_out_size(q.dtype);
    A.rows = (q.transa & 1) ? q.k : q.m; A.cols = (q.transa & 1) ? q.m : q.k;
    B.rows = (q.transb & 1) ? q.n : q.k; B.cols = (q.transb & 1) ? q.k : q.n;
    C.rows = q.m; C.cols = q.n;
    A.ld_user = q.lda; B.ld_user = q.ldb; C.ld_user = q.ldc;
    A.host = (const char *)q.a; B.host = (const char *)q.b; C.host = (const char *)q.c;
    if ((it.product && (classify(q.a) == PTR_DEVICE || classify(q.b) == PTR_DEVICE)) || classify(q.c) == PTR_DEVICE) return 0;
    Operand *ops[3] = {&A, &B, &C};
    for (int o = 0; o < 3; o++) {
      Operand &x = *ops[o];
      x.ld_dev = (int64_t)(round_up((size_t)(x.rows > 0 ? x.rows : 1) * x.es, 128) / x.es);
      if (o < 2) { if (it.product) ab_bytes += round_up(x.bytes_dev(), 256); }
      else c_bytes += round_up(x.bytes_dev(), 256);
    }
  }
  /* packed block: [problem list (grouped launch only) | all A and B | all C]; the list rides in the same upload */
  const bool grouped = batch_is_groupable(p, count);
  const size_t desc_bytes = grouped ? grouped_desc_bytes((size_t)count) : 0;
  ab_bytes += desc_bytes;
  const size_t total = ab_bytes + c_bytes;
  if (total > B200_BATCH_BYTES) return 0;
  int err = reserve_device(ctx, total);
  if (err) return err;
  if ((err = reserve_pinned(ctx, total))) return err;
  size_t off_ab = desc_bytes, off_c = ab_bytes;
  for (Item &it : items) {
    if (it.product) {
      it.A.dev = ctx->dws + off_ab; pack_to(ctx->hws + off_ab, it.A); off_ab += round_up(it.A.bytes_dev(), 256);
      it.B.dev = ctx->dws + off_ab; pack_to(ctx->hws + off_ab, it.B); off_ab += round_up(it.B.bytes_dev(), 256);
    } else {
      it.A.dev = it.B.dev = ctx->dws;
    }
    it.C.dev = ctx->dws + off_c;
    if (it.use_beta) pack_to(ctx->hws + off_c, it.C);
    off_c += round_up(it.C.bytes_dev(), 256);
  }
  cudaStream_t s = ctx->stream;
  std::vector<DeviceGemm> gs;
  gs.reserve(items.size());
  for (Item &it : items) {
    DeviceGemm &g = it.g;
    g.a = it.A.dev; g.b = it.B.dev; g.c = it.C.dev; g.lda = it.A.ld_dev; g.ldb = it.B.ld_dev; g.ldc = it.C.ld_dev;
    gs.push_back(g);
  }
  const int64_t tiles = grouped ? fill_batch_desc(gs, ctx->hws) : 0;
  const size_t up = any_beta ? total : ab_bytes;
  if (up) CK(cudaMemcpyAsync(ctx->dws, ctx->hws, up, cudaMemcpyHostToDevice, s));
  if (grouped) {
    CK(launch_batch_grouped(gs[0].dtype, gs.size(), tiles, ctx->dws, s));
  } else {
    for (const DeviceGemm &g : gs) CK(dispatch(g, s));
  }
  if (c_bytes) CK(cudaMemcpyAsync(ctx->hws + ab_bytes, ctx->dws + ab_bytes, c_bytes, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  for (size_t i = 0; i < items.size(); i++) {
    Item &it = items[i];
    if (!it.product && it.g.beta_re == 1.0 && it.g.beta_im == 0.0) continue;     /* C untouched */
    const Operand &C = it.C;
    const size_t c_off = (size_t)(C.dev - ctx->dws), row_bytes = (size_t)C.rows * C.es;
    char *uc = (char *)p[i].c;
    for (int64_t j = 0; j < C.cols; j++)
      memcpy(uc + (size_t)j * (size_t)C.ld_user * C.es, ctx->hws + c_off + (size_t)j * (size_t)C.ld_dev * C.es, row_bytes);
  }
  *handled = true;
  return 0;
}

B200_HIDDEN int b200_run_batch(const b200_problem *p, int64_t count) {
  {  /* a batch made only of C := 1 * C + 0 members has nothing to do: return before CUDA is touched */
    bool any = false;
    for (int64_t i = 0; i < count && !any; i++) {
      DeviceGemm g;
      read_scalars(&p[i], g);
      const bool product = p[i].k > 0 && !(g.alpha_re == 0.0 && g.alpha_im == 0.0);
      any = product || !(g.beta_re == 1.0 && g.beta_im == 0.0);
    }
    if (!any) return 0;
  }
  ContextLease lease;
  int err = acquire(&lease);
  if (err) return err;
  t_error[0] = 0;
  bool handled = false;
  if ((err = run_batch_packed(lease.c, p, count, &handled))) { lease.failed = true; return err; }
  if (handled) return 0;
  for (int64_t i = 0; i < count; i++) {
    err = run_on_context(lease.c, &p[i]);
    if (err) { lease.failed = true; return err; }
  }
  return 0;
}

#define B200_CONVERT_INPLACE_BYTES ((size_t)4096)
static int run_convert_on_context(Context *ctx, int dir, int64_t n, const void *in, int64_t inc_in, void *out, int64_t inc_out) {
  static const size_t in_sz[4] = {4, 8, 2, 2}, out_sz[4] = {2, 2, 4, 8};
  if (n <= 0) return 0;
  /* increment 0 (the reference's loops then touch element 0 only): every output takes in[0]; every input lands
   * in out[0], the last one winning */
  if (inc_out == 0 && n > 1) { in = (const char *)in + (n - 1) * inc_in * (int64_t)in_sz[dir]; n = 1; }
  cudaStream_t s = ctx->stream;
  const int64_t ainc_in = inc_in < 0 ? -inc_in : inc_in, ainc_out = inc_out < 0 ? -inc_out : inc_out;
  bool in_dev = classify(in) == PTR_DEVICE, out_dev = classify(out) == PTR_DEVICE;
  if (t_foreign) return (int)cudaErrorInvalidDevice;
  if (in_dev || out_dev) { int e = order_after_caller(ctx); if (e) return e; }
  /* spans in elements of the strided arrays; the interface already moved negative-increment
   * pointers to the lowest address and the kernel walks with the signed increment */
  size_t in_bytes = ((size_t)(n - 1) * ainc_in + 1) * in_sz[dir];
  size_t out_bytes = ((size_t)(n - 1) * ainc_out + 1) * out_sz[dir];
  const char *in_lo = (const char *)in + (inc_in < 0 ? (n - 1) * inc_in * (int64_t)in_sz[dir] : 0);
  char *out_lo = (char *)out + (inc_out < 0 ? (n - 1) * inc_out * (int64_t)out_sz[dir] : 0);
  size_t in_off = 0, out_off = round_up(in_bytes, 256);
  size_t need = (in_dev ? 0 : out_off) + (out_dev ? 0 : round_up(out_bytes, 256));
  int err;
  /* A handful of elements in host memory (test/compare_sgemm_sbgemm.c converts its matrices ONE element per call,
   * 12 million calls): the kernel reads and writes the pinned staging block in place over PCIe (pinned memory is
   * device-addressable under unified addressing) -- one launch and one synchronisation, no copy-engine round trips. */
  if (!in_dev && !out_dev && in_bytes + out_bytes <= B200_CONVERT_INPLACE_BYTES) {
    if ((err = reserve_pinned(ctx, out_off + round_up(out_bytes, 256)))) return err;
    char *h_in = ctx->hws, *h_out = ctx->hws + out_off;
    memcpy(h_in, in_lo, in_bytes);
    if (ainc_out > 1) memcpy(h_out, out_lo, out_bytes);        /* the caller's bytes between the elements come along */
    CK(launch_convert(dir, n, h_in + ((const char *)in - in_lo), inc_in, h_out + ((char *)out - out_lo), inc_out, s));
    CK(cudaStreamSynchronize(s));
    memcpy(out_lo, h_out, out_bytes);
    return 0;
  }
  if (need) { err = reserve_device(ctx, out_off + round_up(out_bytes, 256)); if (err) return err; }
  const char *d_in_lo = in_dev ? in_lo : ctx->dws + in_off;
  char *d_out_lo = out_dev ? out_lo : ctx->dws + out_off;
  if (!in_dev) CK(cudaMemcpyAsync((void *)d_in_lo, in_lo, in_bytes, cudaMemcpyHostToDevice, s));
  /* a strided output keeps the caller's bytes between elements: bring them along */
  if (!out_dev && ainc_out > 1) CK(cudaMemcpyAsync(d_out_lo, out_lo, out_bytes, cudaMemcpyHostToDevice, s));
  const char *d_in = d_in_lo + ((const char *)in - in_lo);
  char *d_out = d_out_lo + ((char *)out - out_lo);
  CK(launch_convert(dir, n, d_in, inc_in, d_out, inc_out, s));
  if (!out_dev) CK(cudaMemcpyAsync(out_lo, d_out_lo, out_bytes, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return 0;
}

B200_HIDDEN int b200_run_convert(int dir, int64_t n, const void *in, int64_t inc_in, void *out,
                                 int64_t inc_out) {
  ContextLease lease;
  int err = acquire(&lease);
  if (err) return err;
  err = run_convert_on_context(lease.c, dir, n, in, inc_in, out, inc_out);
  lease.failed = err != 0;
  return err;
}

#include "runtime_bf16.inl"

B200_HIDDEN void b200_fatal(const char *where, int err) {
  fprintf(stderr, "openblas_b200: %s failed: %s [cuda error %d]. There is no CPU fallback.\n", where,
          t_error[0] ? t_error : "unknown error", err);
  abort();
}

B200_EXPORT int b200_gemm(int dtype, int transa, int transb, int64_t m, int64_t n, int64_t k,
                          const void *alpha, const void *A, int64_t lda, const void *B, int64_t ldb,
                          const void *beta, void *C, int64_t ldc) {
  if (m <= 0 || n <= 0) return 0;
  b200_problem p = {dtype, transa, transb, m, n, k, lda, ldb, ldc, alpha, beta, A, B, C};
  return b200_run_problem(&p);
}

B200_EXPORT int b200_gemm_async(int dtype, int transa, int transb, int64_t m, int64_t n, int64_t k,
                                const void *alpha, const void *A, int64_t lda, const void *B,
                                int64_t ldb, const void *beta, void *C, int64_t ldc, void *stream) {
  if (m <= 0 || n <= 0) return 0;
  int err = ensure_init();
  if (err) return err;
  b200_problem p = {dtype, transa, transb, m, n, k, lda, ldb, ldc, alpha, beta, A, B, C};
  DeviceGemm g;
  g.dtype = dtype; g.transa = transa; g.transb = transb; g.m = m; g.n = n; g.k = k;
  g.lda = lda; g.ldb = ldb; g.ldc = ldc; g.a = A; g.b = B; g.c = C;
  read_scalars(&p, g);
  t_error[0] = 0;
  /* CUDA convention: a NULL stream is the legacy default stream.  Nothing is synchronised here;
   * ordering against the caller's other work is the caller's stream semantics. */
  CK(dispatch(g, (cudaStream_t)stream));
  return 0;
}

B200_EXPORT void b200_set_kernel(int kernel) { g_forced_kernel.store(kernel); }
B200_EXPORT int b200_get_kernel(void) { return g_forced_kernel.load(); }
B200_EXPORT uint64_t b200_launch_count(void) { return g_launches.load(); }
B200_EXPORT const char *b200_last_kernel(void) { return t_last_kernel; }
B200_EXPORT const char *b200_last_error(void) { return t_error; }
B200_EXPORT const char *b200_version(void) { return "openblas_b200 0.1 (sm_100a; ABI of OpenBLAS 0.3.28.dev)"; }

B200_EXPORT int b200_init(int device) {
  if (g_device.load(std::memory_order_acquire) < 0) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return set_error(e, "cudaSetDevice");
  }
  return ensure_init();
}

/* The counterpart of gotoblas_quit (driver/others/memory.c:1566-1600, the reference's library destructor):
 * frees every pooled context (streams, events, device workspace, pinned staging and slot rings) and joins the
 * host copy pool.  Safe to call with no call in flight; the next call initialises again. */
static void shutdown_impl(void) {
  if (g_device.load(std::memory_order_acquire) < 0 || getpid() != g_init_pid) return;
  if (HostPool::instance()) HostPool::instance()->stop();
  std::vector<Context *> ctxs;
  { std::lock_guard<std::mutex> lk(g_mu); ctxs.swap(g_free_ctx); }
  int cur = -1;
  const int dev = g_device.load(std::memory_order_acquire);
  if (cudaGetDevice(&cur) == cudaSuccess && cur != dev) cudaSetDevice(dev); else cur = -1;
#ifndef B200_HOSTSIM
  { cudaMemPool_t pool; if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) cudaMemPoolTrimTo(pool, 0); cudaGetLastError(); }
#endif
  for (Context *c : ctxs) {
    scrub(c);
    if (c->dws) cudaFree(c->dws);
    if (c->hws) cudaFreeHost(c->hws);
    for (int i = 0; i < Context::kSlots; i++) {
      if (c->in_slot[i]) cudaFreeHost(c->in_slot[i]);
      if (c->out_slot[i]) cudaFreeHost(c->out_slot[i]);
      if (c->in_ev[i]) cudaEventDestroy(c->in_ev[i]);
      if (c->out_ev[i]) cudaEventDestroy(c->out_ev[i]);
    }
    for (cudaEvent_t e : c->events) cudaEventDestroy(e);
    if (c->order_ev) cudaEventDestroy(c->order_ev);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->s_in) cudaStreamDestroy(c->s_in);
    if (c->s_out) cudaStreamDestroy(c->s_out);
    delete c;
  }
  cudaGetLastError();
  if (cur >= 0) cudaSetDevice(cur);
}
B200_EXPORT void b200_shutdown(void) { shutdown_impl(); }
/* At unload / process exit only the host copy pool is joined.  No CUDA call is made from a library destructor:
 * at exit the CUDA runtime's own teardown may already have run (in any order relative to ours), and touching a
 * dead runtime crashes the process after main() has returned; the operating system reclaims the rest.  A program
 * that dlclose()s the library mid-run calls b200_shutdown() first. */
__attribute__((destructor)) static void b200_library_destructor(void) {
  if (g_device.load(std::memory_order_acquire) < 0 || getpid() != g_init_pid) return;
  if (HostPool::instance()) HostPool::instance()->stop();
}

B200_EXPORT void *b200_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (ensure_init()) return nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
B200_EXPORT void b200_host_free(void *p) { if (p) cudaFreeHost(p); }

}  // extern "C"
