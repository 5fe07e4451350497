/* async_copy.cuh -- cp.async (LDGSTS) helpers shared by the cp.async-fed kernels. */
#pragma once
#include <cuda_runtime.h>

namespace b200 {

/* 16-byte copy global -> shared; bytes beyond src_bytes (0, 4, 8, 12 or 16) are zero-filled. */
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

/* DMMA.8x8x4: D(8x8) += A(8x4, row) * B(4x8, col); lane l holds A[l/4][l%4], B[l%4][l/4] and
 * D[l/4][2*(l%4) + {0,1}]. */
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

/* tile index -> (m-block, n-block): walk C in bands of BAND m-blocks so the tiles in flight
 * share few A row-panels and B column-panels in L2 */
template <int BAND>
__device__ __forceinline__ void banded_tile_coords(int64_t t, int64_t tiles_m, int64_t tiles_n, int64_t &bm,
                                                   int64_t &bn) {
  int64_t per_band = (int64_t)BAND * tiles_n;
  int64_t band = t / per_band, r = t % per_band;
  int64_t band_rows = tiles_m - band * BAND;
  if (band_rows > BAND) band_rows = BAND;
  bm = band * BAND + r % band_rows;
  bn = r / band_rows;
}

}  // namespace b200
