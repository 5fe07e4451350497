/*
 * cuda_shim.cpp -- TEST INFRASTRUCTURE.  A host-memory stand-in for the ~20 CUDA runtime calls
 * openblas_b200/csrc/runtime.cu makes, so that the library's HOST logic (pointer classification,
 * packing, slot rings, panel pipeline, batch staging, the level-3 recursions) can run and be checked
 * on a machine without a GPU (tests/test_hostsim.py).  Never linked into the product.
 *
 *   - "device" memory is malloc'd, registered, and POISONED with 0xFF bytes (NaN in every precision),
 *     so a result that depends on device memory nobody wrote shows up as NaN;
 *   - "pinned" memory is malloc'd and registered; anything else classifies as pageable host memory;
 *   - copies execute immediately, in program order; streams and events are inert handles.  That is a
 *     legal serialisation of a correctly synchronised program: it checks WHAT is copied and computed,
 *     not whether an event wait is missing (the GPU tests and compute-sanitizer cover that).
 */
#include <cuda_runtime.h>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <map>
#include <mutex>

namespace {
/* leaked on purpose: the library's destructor (b200_shutdown) frees blocks at exit, after static destructors ran */
std::mutex &g_mu = *new std::mutex;
std::map<const char *, std::pair<size_t, int>> &g_blocks = *new std::map<const char *, std::pair<size_t, int>>;   /* base -> (bytes, 1 device / 2 pinned) */
size_t g_device_bytes = 0;
std::atomic<size_t> g_copies{0};

void *alloc_block(size_t n, int kind) {
  if (n == 0) n = 1;
  void *p = nullptr;
  if (posix_memalign(&p, 256, (n + 255) / 256 * 256)) return nullptr;
  memset(p, kind == 1 ? 0xFF : 0xA5, n);
  std::lock_guard<std::mutex> lk(g_mu);
  g_blocks[(const char *)p] = {n, kind};
  if (kind == 1) g_device_bytes += n;
  return p;
}
int kind_of(const void *q) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_blocks.upper_bound((const char *)q);
  if (it == g_blocks.begin()) return 0;
  --it;
  return ((const char *)q < it->first + it->second.first) ? it->second.second : 0;
}
void free_block(void *p) {
  if (!p) return;
  { std::lock_guard<std::mutex> lk(g_mu); g_blocks.erase((const char *)p); }
  free(p);
}
}  // namespace

extern "C" {

cudaError_t cudaGetDeviceCount(int *count) { *count = 1; return cudaSuccess; }
cudaError_t cudaGetDevice(int *device) { *device = 0; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *prop, int) {
  memset(prop, 0, sizeof *prop);
  strcpy(prop->name, "HOSTSIM (no GPU)");
  prop->major = 10; prop->minor = 0; prop->multiProcessorCount = 148;
  return cudaSuccess;
}
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char *cudaGetErrorName(cudaError_t e) { return e == cudaSuccess ? "cudaSuccess" : "cudaErrorHostSim"; }
const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "host simulation error"; }

cudaError_t cudaMalloc(void **p, size_t n) { *p = alloc_block(n, 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFree(void *p) { free_block(p); return cudaSuccess; }
cudaError_t cudaHostAlloc(void **p, size_t n, unsigned int) { *p = alloc_block(n, 2); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFreeHost(void *p) { free_block(p); return cudaSuccess; }

cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *at, const void *ptr) {
  memset(at, 0, sizeof *at);
  const int k = kind_of(ptr);
  at->type = k == 1 ? cudaMemoryTypeDevice : k == 2 ? cudaMemoryTypeHost : cudaMemoryTypeUnregistered;
  return cudaSuccess;
}

cudaError_t cudaMemcpy(void *dst, const void *src, size_t n, cudaMemcpyKind) { g_copies++; memmove(dst, src, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind, cudaStream_t) { g_copies++; memmove(dst, src, n); return cudaSuccess; }
cudaError_t cudaMemcpy2DAsync(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height, cudaMemcpyKind,
                              cudaStream_t) {
  g_copies++;
  for (size_t r = 0; r < height; r++) memmove((char *)dst + r * dpitch, (const char *)src + r * spitch, width);
  return cudaSuccess;
}

cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned int) { *s = (cudaStream_t)malloc(8); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamQuery(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
cudaError_t cudaDeviceEnablePeerAccess(int, unsigned int) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned int) { return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned int) { *e = (cudaEvent_t)malloc(8); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }

/* what csrc/summa.cu needs beyond the above; a 1 x 1 grid never talks to a peer, so the IPC and driver entry points only
 * have to exist (they answer "not supported": the driver then takes its NCCL-synchronised branch, unused at world == 1) */
cudaError_t cudaMallocAsync(void **p, size_t n, cudaStream_t) { return cudaMalloc(p, n); }
cudaError_t cudaFreeAsync(void *p, cudaStream_t) { return cudaFree(p); }
cudaError_t cudaMemset(void *p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *, void *) { return cudaErrorNotSupported; }
cudaError_t cudaIpcOpenMemHandle(void **, cudaIpcMemHandle_t, unsigned int) { return cudaErrorNotSupported; }
cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }
cudaError_t cudaGetDriverEntryPoint(const char *, void **fn, unsigned long long, cudaDriverEntryPointQueryResult *q) {
  *fn = nullptr;
  if (q) *q = cudaDriverEntryPointSymbolNotFound;
  return cudaSuccess;
}

/* hooks for the tests */
__attribute__((visibility("default"))) void *hostsim_device_alloc(size_t n) { return alloc_block(n, 1); }
__attribute__((visibility("default"))) void *hostsim_pinned_alloc(size_t n) { return alloc_block(n, 2); }
__attribute__((visibility("default"))) void hostsim_free(void *p) { free_block(p); }
__attribute__((visibility("default"))) size_t hostsim_copy_count(void) { return g_copies.load(); }

}  // extern "C"
