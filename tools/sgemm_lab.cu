/*
 * sgemm_lab.cu -- MEASUREMENT TOOL (not part of the library): runs one configuration of the
 * warp-specialised SGEMM (openblas_b200/csrc/sgemm_ws.cuh) or the round-1 kernel on n^3 problems for
 * the four op combinations, checks sampled entries against a double-precision dot product and prints
 * TFLOP/s.  One configuration per process so a trap in one does not poison the others.
 *
 *   sgemm_lab <cfg> [n=8192] [reps=5] [m n k]
 *   cfg 0 = round-1 kernel (launch_sgemm_ffma), 1.. = Cfg<> instances below
 * build: see tools/sgemm_lab.sh
 */
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../openblas_b200/csrc/sgemm_ws.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

namespace b200 {
static int g_sms = 148;
int sm_count() { return g_sms; }
static const char *g_last = "";
void count_launch(const char *name) { g_last = name; }
}  // namespace b200
using namespace b200;

__global__ void fill(float *p, size_t n, unsigned seed) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    unsigned x = (unsigned)(i * 2654435761u) ^ seed;
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    p[i] = (float)(x >> 8) * (1.0f / 16777216.0f) - 0.5f;
  }
}

/* sampled check: err / (k eps gauge) of entry (i, j) */
__global__ void check(int ta, int tb, int64_t m, int64_t n, int64_t k, const float *A, int64_t lda, const float *B, int64_t ldb,
                      const float *C, int64_t ldc, const float *C0, float alpha, float beta, int samples, double *worst) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= samples) return;
  unsigned x = s * 2654435761u + 12345u;
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15;
  int64_t i = (s < 64) ? (s & 1 ? m - 1 - (s >> 1) % m : (s >> 1) % m) : (int64_t)(x % (unsigned)m);
  x *= 0x846ca68bu; x ^= x >> 16;
  int64_t j = (s < 64) ? (s & 2 ? n - 1 - (s >> 2) % n : (s >> 2) % n) : (int64_t)(x % (unsigned)n);
  double acc = 0, gauge = 0;
  for (int64_t l = 0; l < k; l++) {
    double a = ta ? A[l + i * lda] : A[i + l * lda];
    double b = tb ? B[j + l * ldb] : B[l + j * ldb];
    acc += a * b; gauge += fabs(a * b);
  }
  double want = alpha * acc + (beta != 0.f ? (double)beta * C0[i + j * ldc] : 0.0);
  gauge = fabs(alpha) * gauge + (beta != 0.f ? fabs((double)beta * C0[i + j * ldc]) : 0.0);
  double r = fabs((double)C[i + j * ldc] - want) / ((double)k * 1.1920929e-7 * gauge + 1e-300);
  atomicMax((unsigned long long *)worst, __double_as_longlong(r));   /* non-negative doubles order like integers */
}

typedef cudaError_t (*LaunchFn)(const DeviceGemm &, cudaStream_t);
struct Variant { const char *name; LaunchFn fn; };
static Variant variants[] = {
    {"round1 128x128x16 (launch_sgemm_ffma)", launch_sgemm_ffma},
    {"ws TM16 2x4 256x128 S6 rc224 2feed", sws::launch<sws::Cfg<16, 2, 4, 6, 1, 2, 4, 224, 56>>},
    {"ws TM16 2x4 256x128 S6 rc224 xpf", sws::launch<sws::Cfg<16, 2, 4, 6, 1, 2, 4, 224, 56, true>>},
    {"ws TM16 2x4 256x128 S6 rc224 xpf snake", sws::launch<sws::Cfg<16, 2, 4, 6, 1, 2, 4, 224, 56, true, true>>},
    {"ws TM16 2x4 256x128 S6 rc224 snake", sws::launch<sws::Cfg<16, 2, 4, 6, 1, 2, 4, 224, 56, false, true>>},
    {"ws TM16 2x4 256x128 S5 STG3 rc224 xpf", sws::launch<sws::Cfg<16, 2, 4, 5, 1, 3, 4, 224, 56, true>>},
    {"ws TM16 2x4 256x128 S4 rc224 xpf", sws::launch<sws::Cfg<16, 2, 4, 4, 1, 2, 4, 224, 56, true>>},
    {"ws TM16 2x4 256x128 S6 rc232 xpf", sws::launch<sws::Cfg<16, 2, 4, 6, 1, 2, 4, 232, 40, true>>},
    {"ws TM8  4x4 256x128 S5 rc112 xpf", sws::launch<sws::Cfg<8, 4, 4, 5, 1, 2, 4, 112, 32, true>>},
    {"ws TM8  3x4 192x128 S5 pw1 xpf", sws::launch<sws::Cfg<8, 3, 4, 5, 1, 2, 1, 0, 0, true>>},
    {"ws TM16 2x4 256x128 S6 pw1 (168 regs)", sws::launch<sws::Cfg<16, 2, 4, 6, 1>>},
    {"ws TM16 2x4 S5 rc224 xpf (library r2a)", sws::launch<sws::Cfg<16, 2, 4, 5, 1, 2, 4, 224, 56, true>>},
    {"ws TM16 2x4 S5 rc224 xpf HELP", sws::launch<sws::Cfg<16, 2, 4, 5, 1, 2, 4, 224, 56, true, false, true>>},
    {"ws TM16 2x4 S5 STG3 rc224 xpf HELP", sws::launch<sws::Cfg<16, 2, 4, 5, 1, 3, 4, 224, 56, true, false, true>>},
    {"ws TM16 2x4 S4 STG3 rc224 xpf HELP", sws::launch<sws::Cfg<16, 2, 4, 4, 1, 3, 4, 224, 56, true, false, true>>},
};

int main(int argc, char **argv) {
  const int cfg = argc > 1 ? atoi(argv[1]) : 1;
  const int64_t nn = argc > 2 ? atoll(argv[2]) : 8192;
  const int reps = argc > 3 ? atoi(argv[3]) : 5;
  int64_t m = nn, n = nn, k = nn;
  if (argc > 6) { m = atoll(argv[4]); n = atoll(argv[5]); k = atoll(argv[6]); }
  if (cfg < 0 || cfg >= (int)(sizeof variants / sizeof variants[0])) { fprintf(stderr, "no such cfg\n"); return 2; }
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  g_sms = prop.multiProcessorCount;
  const int64_t big = m > k ? m : k, big2 = n > k ? n : k;
  const int64_t lda = big + 4 * (argc > 7 ? atoi(argv[7]) : 0), ldb = big2 + 4 * (argc > 7 ? atoi(argv[7]) : 0), ldc = m;
  float *A, *B, *C, *C0;
  double *worst;
  CK(cudaMalloc(&A, (size_t)lda * big * 4)); CK(cudaMalloc(&B, (size_t)ldb * big2 * 4));
  CK(cudaMalloc(&C, (size_t)ldc * n * 4)); CK(cudaMalloc(&C0, (size_t)ldc * n * 4));
  CK(cudaMalloc(&worst, 8));
  fill<<<1024, 256>>>(A, (size_t)lda * big, 1u); fill<<<1024, 256>>>(B, (size_t)ldb * big2, 2u); fill<<<1024, 256>>>(C0, (size_t)ldc * n, 3u);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int op = 0; op < 4; op++) {
    const int ta = op & 1, tb = op >> 1;
    for (int pass = 0; pass < 2; pass++) {          /* pass 0: alpha = 1, beta = 0, timed; pass 1: alpha, beta != 0, checked only */
      DeviceGemm g;
      g.dtype = B200_S; g.transa = ta; g.transb = tb; g.m = m; g.n = n; g.k = k; g.lda = lda; g.ldb = ldb; g.ldc = ldc;
      g.a = A; g.b = B; g.c = C; g.alpha_re = pass ? 0.7 : 1.0; g.alpha_im = 0; g.beta_re = pass ? 1.3 : 0.0; g.beta_im = 0; g.tri = 0;
      CK(cudaMemcpy(C, C0, (size_t)ldc * n * 4, cudaMemcpyDeviceToDevice));
      cudaError_t e = variants[cfg].fn(g, 0);
      if (e != cudaSuccess) { printf("cfg %d %s op %c%c: launch -> %s\n", cfg, variants[cfg].name, "NT"[ta], "NT"[tb], cudaGetErrorString(e)); return 1; }
      e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("cfg %d %s op %c%c: run -> %s\n", cfg, variants[cfg].name, "NT"[ta], "NT"[tb], cudaGetErrorString(e)); return 1; }
      CK(cudaMemset(worst, 0, 8));
      check<<<16, 256>>>(ta, tb, m, n, k, A, lda, B, ldb, C, ldc, C0, (float)g.alpha_re, (float)g.beta_re, 4096, worst);
      double w;
      CK(cudaMemcpy(&w, worst, 8, cudaMemcpyDeviceToHost));
      float ms = 0;
      if (pass == 0) {
        for (int i = 0; i < 2; i++) variants[cfg].fn(g, 0);
        CK(cudaEventRecord(e0));
        for (int i = 0; i < reps; i++) variants[cfg].fn(g, 0);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= reps;
        printf("cfg %d %-36s %c%c %lldx%lldx%lld: %7.3f ms %6.2f TFLOP/s  worst err/(k eps gauge) = %.4f %s\n", cfg, variants[cfg].name, "NT"[ta],
               "NT"[tb], (long long)m, (long long)n, (long long)k, ms, 2.0 * m * n * k / (ms * 1e-3) / 1e12, w, w <= 2.0 ? "ok" : "WRONG");
      } else {
        printf("cfg %d %-36s %c%c alpha=0.7 beta=1.3: worst = %.4f %s\n", cfg, variants[cfg].name, "NT"[ta], "NT"[tb], w, w <= 2.0 ? "ok" : "WRONG");
      }
      fflush(stdout);
    }
  }
  return 0;
}
