/*
 * peaks.cu -- MEASUREMENT TOOL (not part of the library): measures on the box the roofline
 * denominators MEASURED_PEAKS.json does not carry -- the FP64 and FP32 (non-TF32) dense peaks
 * -- two ways:
 *   (1) instruction-issue microbenchmarks: DMMA.8x8x4 (mma.sync m8n8k4 f64), DFMA, FFMA and
 *       FFMA2 (fma.rn.f32x2) from registers, all SMs, 4..16 warps per SM
 *   (2) cuBLAS cublasDgemm / cublasSgemm (pedantic math: no TF32) / bf16 GemmEx at n = 8192
 *       as the practical "library ceiling" (cuBLAS is used ONLY here, never in the product)
 * Prints one JSON object; tools/run_peaks.sh stores it as gpurun_out/peaks.json.
 */
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int ITERS = 2048;

__global__ void k_dmma(double *out, double seed) {
  double a = seed + threadIdx.x, b = seed * 0.5;
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; i++) { c[i][0] = i; c[i][1] = -i; }
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dfma(double *out, double seed) {
  double a = seed + threadIdx.x, b = seed * 0.5;
  double c[32];
#pragma unroll
  for (int i = 0; i < 32; i++) c[i] = i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 32; i++) c[i] = fma(a, b, c[i]);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 32; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma(float *out, float seed) {
  float a = seed + threadIdx.x, b = seed * 0.5f;
  float c[32];
#pragma unroll
  for (int i = 0; i < 32; i++) c[i] = i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 32; i++) c[i] = fmaf(a, b, c[i]);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 32; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma2(float *out, float seed) {
  unsigned long long a, b, c[32];
  float a0 = seed + threadIdx.x, b0 = seed * 0.5f;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a0 + 1.f));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b0));
#pragma unroll
  for (int i = 0; i < 32; i++) { float x = i; asm("mov.b64 %0, {%1, %2};" : "=l"(c[i]) : "f"(x), "f"(-x)); }
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 32; i++) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c[i]) : "l"(a), "l"(b));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 32; i++) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(c[i])); s += x + y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static double time_ms(F launch, int reps = 5) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch();
  CK(cudaDeviceSynchronize());
  std::vector<float> t;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); t.push_back(ms);
  }
  return *std::min_element(t.begin(), t.end());
}

/* peaks [quick] [device]: "quick" = issue rates at 8 and 16 warps/SM + cuBLAS D/S at 8192 only (a few seconds:
 * bench.py runs it inside the benchmark lease so that the FP64 / FP32 roofline denominators are measured on the
 * same box, same clocks, as the numbers they divide) */
int main(int argc, char **argv) {
  bool quick = false; int device = 0;
  for (int i = 1; i < argc; i++) { if (!strcmp(argv[i], "quick")) quick = true; else device = atoi(argv[i]); }
  CK(cudaSetDevice(device));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, device));
  int sms = p.multiProcessorCount;
  void *buf; CK(cudaMalloc(&buf, (size_t)sms * 8 * 1024 * 8));
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d", p.name, sms, p.clockRate);
  for (int warps : {4, 8, 16}) {
    if (quick && warps == 4) continue;
    int threads = warps * 32, blocks = sms * 2;
    double tot_threads = (double)blocks * threads;
    double ms;
    ms = time_ms([&] { k_dmma<<<blocks, threads>>>((double *)buf, 1.0); });
    printf(", \"dmma_tflops_w%d\": %.2f", warps, (tot_threads / 32) * ITERS * 16 * 512.0 / (ms * 1e-3) / 1e12);
    ms = time_ms([&] { k_dfma<<<blocks, threads>>>((double *)buf, 1.0); });
    printf(", \"dfma_tflops_w%d\": %.2f", warps, tot_threads * ITERS * 32 * 2.0 / (ms * 1e-3) / 1e12);
    ms = time_ms([&] { k_ffma<<<blocks, threads>>>((float *)buf, 1.0f); });
    printf(", \"ffma_tflops_w%d\": %.2f", warps, tot_threads * ITERS * 32 * 2.0 / (ms * 1e-3) / 1e12);
    ms = time_ms([&] { k_ffma2<<<blocks, threads>>>((float *)buf, 1.0f); });
    printf(", \"ffma2_tflops_w%d\": %.2f", warps, tot_threads * ITERS * 32 * 4.0 / (ms * 1e-3) / 1e12);
  }
  CK(cudaGetLastError());

  cublasHandle_t h; cublasCreate(&h);
  for (int n : {4096, 8192, 16384}) {
    if (quick && n != 8192) continue;
    size_t bytes = (size_t)n * n * 8;
    double *a, *b, *c;
    if (cudaMalloc(&a, bytes) != cudaSuccess || cudaMalloc(&b, bytes) != cudaSuccess || cudaMalloc(&c, bytes) != cudaSuccess) break;
    CK(cudaMemset(a, 0, bytes)); CK(cudaMemset(b, 0, bytes));
    /* non-trivial data (zeros can under-report power-limited throughput) */
    {
      std::vector<double> hbuf((size_t)n * 64);
      for (auto &x : hbuf) x = (double)rand() / RAND_MAX - 0.5;
      for (size_t off = 0; off < (size_t)n * n; off += hbuf.size()) {
        size_t cnt = std::min(hbuf.size(), (size_t)n * n - off);
        CK(cudaMemcpy(a + off, hbuf.data(), cnt * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(b + off, hbuf.data(), cnt * 8, cudaMemcpyHostToDevice));
      }
    }
    double one = 1.0, zero = 0.0;
    cublasSetMathMode(h, CUBLAS_DEFAULT_MATH);
    double ms = time_ms([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, a, n, b, n, &zero, c, n); }, n >= 16384 ? 3 : 5);
    printf(", \"cublas_dgemm_tflops_%d\": %.2f", n, 2.0 * n * n * n / (ms * 1e-3) / 1e12);
    float fone = 1.f, fzero = 0.f;
    /* reuse the buffers as fp32 random data */
    cublasSetMathMode(h, CUBLAS_PEDANTIC_MATH);
    ms = time_ms([&] { cublasSgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &fone, (float *)a, n, (float *)b, n, &fzero, (float *)c, n); });
    printf(", \"cublas_sgemm_pedantic_tflops_%d\": %.2f", n, 2.0 * n * n * n / (ms * 1e-3) / 1e12);
    cublasSetMathMode(h, CUBLAS_DEFAULT_MATH);
    if (n == 8192 && !quick) {
      /* bf16 values: fill with small numbers to avoid NaN patterns */
      std::vector<__nv_bfloat16> hb((size_t)n * 64);
      for (auto &x : hb) x = __float2bfloat16((float)rand() / RAND_MAX - 0.5f);
      for (size_t off = 0; off < (size_t)n * n; off += hb.size()) {
        CK(cudaMemcpy((__nv_bfloat16 *)a + off, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
        CK(cudaMemcpy((__nv_bfloat16 *)b + off, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
      }
      ms = time_ms([&] { cublasGemmEx(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &fone, a, CUDA_R_16BF, n, b, CUDA_R_16BF, n, &fzero, c, CUDA_R_32F, n, CUBLAS_COMPUTE_32F, CUBLAS_GEMM_DEFAULT); }, 10);
      printf(", \"cublas_bf16_f32out_tflops_%d\": %.2f", n, 2.0 * n * n * n / (ms * 1e-3) / 1e12);
      ms = time_ms([&] { cublasGemmEx(h, CUBLAS_OP_T, CUBLAS_OP_N, n, n, n, &fone, a, CUDA_R_16BF, n, b, CUDA_R_16BF, n, &fzero, c, CUDA_R_32F, n, CUBLAS_COMPUTE_32F, CUBLAS_GEMM_DEFAULT); }, 10);
      printf(", \"cublas_bf16_f32out_TN_tflops_%d\": %.2f", n, 2.0 * n * n * n / (ms * 1e-3) / 1e12);
    }
    cudaFree(a); cudaFree(b); cudaFree(c);
  }
  printf("}\n");
  return 0;
}
