/*
 * ffma_probe.cu -- MEASUREMENT TOOL for round 2 (not part of the library): what the FP32 pipe of one SM
 * sustains for the INNER LOOP SHAPES an SGEMM kernel can have, before any kernel is written around
 * them.  Every variant is the k loop of a register-blocked tile, operands either kept in registers
 * (pure issue rate) or re-loaded from shared memory every k with LDS.128 in the conflict-free 2 x 2
 * quad pattern of profiles/r01_lds_wavefronts.txt:
 *
 *     tile   rows x cols per thread, rows held as f32x2 pairs (FFMA2) or scalars (FFMA)
 *     lds    fragments from registers only / from shared memory each k
 *     warps  resident warps per SM (CTA = 128 or 256 threads, 1..4 CTAs per SM)
 *
 * Round-1 findings it should explain: the 8 x 8 FFMA2 kernel reaches 57 TFLOP/s with loads and barriers,
 * 60 with neither, while the register-only FFMA2 peak is 67.6 and cuBLAS's SGEMM runs at 66.8.
 * usage: ffma_probe            (prints one line per variant; run under ncu for stall reasons)
 */
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
typedef unsigned long long u64;
__device__ __forceinline__ void ffma2(u64 &c, u64 a, float b) {
  u64 bb;
  asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(bb));
}

constexpr int KSTEPS = 16, LDS_STRIDE = 132;     /* one k tile of a 128-wide operand, as in sgemm_ffma.cu */

/* ROWS x COLS outputs per thread; PACKED: rows as f32x2 pairs */
template <int ROWS, int COLS, bool PACKED, bool FROM_SMEM>
__global__ void probe(float *out, int iters) {
  __shared__ __align__(16) float sa[KSTEPS * LDS_STRIDE], sb[KSTEPS * LDS_STRIDE];
  for (int i = threadIdx.x; i < KSTEPS * LDS_STRIDE; i += blockDim.x) { sa[i] = 1.0f + i * 1e-6f; sb[i] = 0.5f - i * 1e-6f; }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int pm = ((lane >> 2) & 3) * 2 + (lane & 1), pn = (lane >> 4) * 2 + ((lane >> 1) & 1);
  const float *pa = sa + pm * 4, *pb = sb + pn * 4;
  float accs[PACKED ? 1 : ROWS][PACKED ? 1 : COLS];
  u64 accp[PACKED ? ROWS / 2 : 1][PACKED ? COLS : 1];
  if (PACKED) { for (int i = 0; i < ROWS / 2; i++) for (int j = 0; j < COLS; j++) accp[i][j] = 0ull; }
  else { for (int i = 0; i < ROWS; i++) for (int j = 0; j < COLS; j++) accs[i][j] = 0.f; }
  float av[ROWS], bv[COLS];
#pragma unroll
  for (int i = 0; i < ROWS; i++) av[i] = 1.0f + i + lane;
#pragma unroll
  for (int j = 0; j < COLS; j++) bv[j] = 0.25f * (j + 1);
  for (int it = 0; it < iters; it++) {
#pragma unroll 4
    for (int k = 0; k < KSTEPS; k++) {
      if (FROM_SMEM) {
#pragma unroll
        for (int i = 0; i < ROWS; i += 4) {      /* rows i..i+3 from one LDS.128, next group 32 floats further */
          const float4 v = *reinterpret_cast<const float4 *>(pa + k * LDS_STRIDE + (i / 4) * 32);
          av[i] = v.x; av[i + 1] = v.y; av[i + 2] = v.z; av[i + 3] = v.w;
        }
#pragma unroll
        for (int j = 0; j < COLS; j += 4) {
          const float4 v = *reinterpret_cast<const float4 *>(pb + k * LDS_STRIDE + (j / 4) * 16);
          bv[j] = v.x; bv[j + 1] = v.y; bv[j + 2] = v.z; bv[j + 3] = v.w;
        }
      }
      if (PACKED) {
#pragma unroll
        for (int i = 0; i < ROWS / 2; i++) {
          u64 ap;
          asm("mov.b64 %0, {%1, %2};" : "=l"(ap) : "f"(av[2 * i]), "f"(av[2 * i + 1]));
#pragma unroll
          for (int j = 0; j < COLS; j++) ffma2(accp[i][j], ap, bv[j]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < ROWS; i++)
#pragma unroll
          for (int j = 0; j < COLS; j++) accs[i][j] = fmaf(av[i], bv[j], accs[i][j]);
      }
    }
  }
  float s = 0.f;
  if (PACKED) {
    for (int i = 0; i < ROWS / 2; i++) for (int j = 0; j < COLS; j++) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(accp[i][j])); s += lo + hi; }
  } else {
    for (int i = 0; i < ROWS; i++) for (int j = 0; j < COLS; j++) s += accs[i][j];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ROWS, int COLS, bool PACKED, bool FROM_SMEM>
static void run(const char *name, int threads, int ctas_per_sm, int sms, float *buf) {
  const int iters = 4096 / (ROWS * COLS / 64);
  auto kern = probe<ROWS, COLS, PACKED, FROM_SMEM>;
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, 0));
  if (occ < ctas_per_sm) { printf("%-34s %3d thr x %d CTA/SM: only %d CTAs fit (registers)\n", name, threads, ctas_per_sm, occ); return; }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  kern<<<sms * ctas_per_sm, threads>>>(buf, iters);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  kern<<<sms * ctas_per_sm, threads>>>(buf, iters);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double flops = 2.0 * ROWS * COLS * KSTEPS * (double)iters * threads * ctas_per_sm * sms;
  cudaFuncAttributes fa;
  CK(cudaFuncGetAttributes(&fa, kern));
  printf("%-34s %3d thr x %d CTA/SM (%2d warps/SM, %3d regs): %6.1f TFLOP/s\n", name, threads, ctas_per_sm, threads * ctas_per_sm / 32,
         fa.numRegs, flops / (ms * 1e-3) / 1e12);
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  float *buf;
  CK(cudaMalloc(&buf, (size_t)sms * 4 * 512 * sizeof(float)));
  printf("# %s, %d SMs\n", prop.name, sms);
  for (int ctas = 1; ctas <= 4; ctas++) {
    run<8, 8, true, false>("8x8 FFMA2 registers", 128, ctas, sms, buf);
    run<8, 8, true, true>("8x8 FFMA2 + 4 LDS.128 / k", 128, ctas, sms, buf);
    run<8, 8, false, false>("8x8 FFMA registers", 128, ctas, sms, buf);
    run<8, 8, false, true>("8x8 FFMA + 4 LDS.128 / k", 128, ctas, sms, buf);
  }
  for (int ctas = 1; ctas <= 2; ctas++) {
    run<16, 8, true, false>("16x8 FFMA2 registers", 128, ctas, sms, buf);
    run<16, 8, true, true>("16x8 FFMA2 + 6 LDS.128 / k", 128, ctas, sms, buf);
    run<8, 16, true, true>("8x16 FFMA2 + 6 LDS.128 / k", 128, ctas, sms, buf);
    run<16, 8, true, true>("16x8 FFMA2 + 6 LDS.128 / k", 256, ctas, sms, buf);
    run<8, 8, true, true>("8x8 FFMA2 + 4 LDS.128 / k", 256, ctas, sms, buf);
  }
  return 0;
}
