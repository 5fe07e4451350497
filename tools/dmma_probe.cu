/* dmma_probe.cu -- MEASUREMENT TOOL: what limits DMMA.8x8x4 throughput in a GEMM-like instruction mix?
 * Variants (all: 148*2 CTAs... grid = SMs, 256 threads, 64 accumulator doubles per lane like the DGEMM warp tile):
 *   0  same A/B registers for every DMMA (the tools/peaks.cu number)
 *   1  8 A x 4 B operand registers, i-outer/j-inner order (DGEMM kernel order), no loads
 *   2  as 1 but operands re-loaded from shared memory every k4 step (12 LDS.64 per 32 DMMA)
 *   3  as 2 plus a __syncthreads every 4 k4 steps
 *   4  as 1 with j-outer/i-inner order
 *   5  PTX m16n8k8 shape (ptxas expands to 4 DMMA.8x8x4 with its own operand order), operands from smem
 */
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include <algorithm>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
constexpr int ITERS = 1024;

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int V>
__global__ void __launch_bounds__(256, 1) probe(double *out, const double *in) {
  __shared__ double sm[6144];
  for (int i = threadIdx.x; i < 6144; i += 256) sm[i] = in[i & 4095];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
  double af[8], bf[4];
#pragma unroll
  for (int i = 0; i < 8; i++) af[i] = sm[lane + 32 * i];
#pragma unroll
  for (int j = 0; j < 4; j++) bf[j] = sm[1024 + lane + 32 * j];
  const double *pa = sm + (lane >> 2) + (lane & 3) * 132 + warp * 8;
  if (V == 5) {
    /* m16n8k8: A 16x8 (4 regs), B 8x8 (2 regs), C 16x8 (4 regs) per lane */
    double c[16][4];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0;
    for (int it = 0; it < ITERS; it++) {
      /* 16 MMAs of 16x8x8 = 64 DMMA.8x8x4 = two "k4 steps" of the 64x32 warp tile; operands: 4 A frags (4 regs) + 4 B frags (2 regs) */
      double a[4][4], b[4][2];
      const double *p = pa + (it & 7) * 264;
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int r = 0; r < 4; r++) a[i][r] = p[i * 8 + r * 528];
#pragma unroll
      for (int j = 0; j < 4; j++) { b[j][0] = p[2048 + j * 8]; b[j][1] = p[2048 + j * 8 + 528]; }
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
          asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                       : "+d"(c[i * 4 + j][0]), "+d"(c[i * 4 + j][1]), "+d"(c[i * 4 + j][2]), "+d"(c[i * 4 + j][3])
                       : "d"(a[i][0]), "d"(a[i][1]), "d"(a[i][2]), "d"(a[i][3]), "d"(b[j][0]), "d"(b[j][1]));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * 256 + threadIdx.x] = s;
    return;
  }
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int k4 = 0; k4 < 2; k4++) {
      if (V == 2 || V == 3) {
        const double *p = pa + ((it * 2 + k4) & 3) * 528;
#pragma unroll
        for (int i = 0; i < 8; i++) af[i] = p[i * 8];
#pragma unroll
        for (int j = 0; j < 4; j++) bf[j] = p[2048 + j * 8];
      }
      if (V == 0) {
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) dmma(acc[i][j][0], acc[i][j][1], bf[0], af[0]);
      } else if (V == 4) {
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
          for (int i = 0; i < 8; i++) dmma(acc[i][j][0], acc[i][j][1], bf[j], af[i]);
      } else {
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) dmma(acc[i][j][0], acc[i][j][1], bf[j], af[i]);
      }
    }
    if (V == 3 && (it & 1)) __syncthreads();
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) s += acc[i][j][0] + acc[i][j][1];
  out[blockIdx.x * 256 + threadIdx.x] = s;
}

template <int V> double run(double *out, const double *in, int sms) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  probe<V><<<sms, 256>>>(out, in); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; r++) {
    CK(cudaEventRecord(e0)); probe<V><<<sms, 256>>>(out, in); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = std::min(best, ms);
  }
  double dmmas = (double)sms * 8 * ITERS * 64;   /* warp-level DMMA.8x8x4 */
  return dmmas * 512.0 / (best * 1e-3) / 1e12;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  double *out, *in; CK(cudaMalloc(&out, sms * 256 * 8)); CK(cudaMalloc(&in, 4096 * 8));
  std::vector<double> h(4096); for (auto &x : h) x = (double)rand() / RAND_MAX - 0.5;
  CK(cudaMemcpy(in, h.data(), 4096 * 8, cudaMemcpyHostToDevice));
  printf("{\"v0_same_operands\": %.2f", run<0>(out, in, sms));
  printf(", \"v1_8x4_operands_i_outer\": %.2f", run<1>(out, in, sms));
  printf(", \"v2_plus_lds\": %.2f", run<2>(out, in, sms));
  printf(", \"v3_plus_lds_barrier\": %.2f", run<3>(out, in, sms));
  printf(", \"v4_8x4_operands_j_outer\": %.2f", run<4>(out, in, sms));
  printf(", \"v5_m16n8k8_lds\": %.2f}\n", run<5>(out, in, sms));
  return 0;
}
