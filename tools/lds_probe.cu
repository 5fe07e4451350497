// Shared-memory wavefront probe: how many LSU wavefronts one warp-wide LDS.128 / LDS.64 costs for a
// given lane -> 16-byte (8-byte) chunk pattern.  Run under
//   ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,smsp__inst_executed_op_shared_ld.sum
// and divide (launch order = the order printed).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct Pattern { int chunk[32]; };

template <int BYTES>
__global__ void probe(Pattern p, int iters, int row_bytes, float *out, long long *cycles) {
  extern __shared__ __align__(16) char smem[];
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) ((float *)smem)[i] = (float)i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem) + p.chunk[lane] * BYTES;
  float acc = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < 16; k++) {
      uint32_t a = base + (uint32_t)(k * row_bytes);
      if (BYTES == 16) {
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
        acc += v.x + v.y + v.z + v.w;
      } else {
        float2 v;
        asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
        acc += v.x + v.y;
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
  struct { const char *name; int bytes; int (*f)(int); } pats[] = {
      {"v4 lane&7        (A today)", 16, [](int l) { return l & 7; }},
      {"v4 lane>>3       (B today)", 16, [](int l) { return l >> 3; }},
      {"v4 lane&3", 16, [](int l) { return l & 3; }},
      {"v4 (lane>>2)&7", 16, [](int l) { return (l >> 2) & 7; }},
      {"v4 lane          (all distinct)", 16, [](int l) { return l; }},
      {"v4 0             (all same)", 16, [](int l) { return 0; }},
      {"v4 lane&15", 16, [](int l) { return l & 15; }},
      {"v4 lane>>1", 16, [](int l) { return l >> 1; }},
      {"v4 lane>>2", 16, [](int l) { return l >> 2; }},
      {"v4 (lane>>3)*8   (B, one 128B row each)", 16, [](int l) { return (l >> 3) * 8; }},
      {"v4 (lane&1)|((lane>>3)&2)... zigzag 2x", 16, [](int l) { return (l & 1) | ((l >> 3) & 2); }},
      {"v4 (lane>>1)&3", 16, [](int l) { return (l >> 1) & 3; }},
      {"v4 (lane>>4)", 16, [](int l) { return l >> 4; }},
      {"v4 ((lane>>4)<<1)|(lane&1)", 16, [](int l) { return ((l >> 4) << 1) | (l & 1); }},
      {"v4 ((lane>>2)&3)*2+(lane&1)   (A, 2x2 quads)", 16, [](int l) { return ((l >> 2) & 3) * 2 + (l & 1); }},
      {"v4 (lane>>4)*2+((lane>>1)&1)  (B, 2x2 quads)", 16, [](int l) { return (l >> 4) * 2 + ((l >> 1) & 1); }},
      {"v2 lane&15", 8, [](int l) { return l & 15; }},
      {"v2 lane>>1", 8, [](int l) { return l >> 1; }},
      {"v2 lane>>3", 8, [](int l) { return l >> 3; }},
      {"v2 lane&7", 8, [](int l) { return l & 7; }},
      {"v2 lane", 8, [](int l) { return l; }},
      {"v2 lane>>2", 8, [](int l) { return l >> 2; }},
  };
  float *out; long long *cyc;
  const int blocks = 148, threads = 256, iters = 256;
  cudaMalloc(&out, blocks * threads * 4); cudaMalloc(&cyc, blocks * 8);
  long long h[148];
  for (auto &pt : pats) {
    Pattern p;
    for (int l = 0; l < 32; l++) p.chunk[l] = pt.f(l);
    for (int row_bytes : {528, 512}) {
      if (pt.bytes == 16) probe<16><<<blocks, threads, 48 * 1024>>>(p, iters, row_bytes, out, cyc);
      else probe<8><<<blocks, threads, 48 * 1024>>>(p, iters, row_bytes, out, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
      // 8 warps x iters x 16 loads per block, one block per SM
      printf("%-44s row %3d B  (%s)\n", pt.name, row_bytes, cudaGetErrorString(e));   /* launch order = ncu ID order */
    }
  }
  return 0;
}
