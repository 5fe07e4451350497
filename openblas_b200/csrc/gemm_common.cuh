/*
 * gemm_common.cuh -- types shared by the device-side dispatcher and all GEMM kernels.
 */
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <math.h>
#include "shim.h"

namespace b200 {

/* A column-major problem with DEVICE pointers and scalars already read from the host.
 * alpha/beta are carried as double pairs: every fp32 value is exactly representable, so the
 * fp32 kernels narrow them back without rounding. */
struct DeviceGemm {
  int dtype;
  int transa, transb;     /* 0 N, 1 T, 2 conj, 3 conj-trans */
  int64_t m, n, k;
  int64_t lda, ldb, ldc;
  const void *a;
  const void *b;
  void *c;
  double alpha_re, alpha_im;
  double beta_re, beta_im;
  /* 0: all of C.  1 / 2: only the lower / upper triangle of the (square) m x n window is computed
   * and written (SYRK family, runtime_level3.inl): kernels that support it skip the tiles outside
   * the triangle and mask the stores of the tiles the diagonal crosses; the others answer
   * cudaErrorNotSupported and the caller falls back to block columns. */
  int tri = 0;
  /* GEMM3M entry points (complex types): the dispatcher may compute the product from three real GEMMs
   * (runtime.cu: gemm3m_on_device); 0 everywhere else */
  int algo3m = 0;
};

/* triangle tests on a tile [m0, m0 + bm) x [n0, n0 + bn) and on an element */
__host__ __device__ __forceinline__ bool tri_outside(int tri, int64_t m0, int bm, int64_t n0, int bn) {
  return tri == 1 ? (m0 + bm - 1 < n0) : tri == 2 ? (m0 > n0 + bn - 1) : false;
}
__host__ __device__ __forceinline__ bool tri_partial(int tri, int64_t m0, int bm, int64_t n0, int bn) {
  return tri == 1 ? (m0 < n0 + bn - 1) : tri == 2 ? (m0 + bm - 1 > n0) : false;
}
/* Kernels with SQUARE tiles enumerate only the nt*(nt+1)/2 tiles of the triangle (row by row of the
 * lower triangle, mirrored for the upper one), so the static round-robin over persistent CTAs stays
 * balanced; kernels with rectangular tiles walk all tiles and skip with tri_outside(). */
__host__ __device__ __forceinline__ int64_t tri_tile_count(int64_t nt) { return nt * (nt + 1) / 2; }
__host__ __device__ __forceinline__ void tri_tile_coords(int64_t u, int tri, int64_t &bm, int64_t &bn) {
  int64_t r = (int64_t)((sqrt(8.0 * (double)u + 1.0) - 1.0) * 0.5);
  while (r * (r + 1) / 2 > u) r--;
  while ((r + 1) * (r + 2) / 2 <= u) r++;
  const int64_t c = u - r * (r + 1) / 2;
  if (tri == 1) { bm = r; bn = c; } else { bm = c; bn = r; }
}
/* Tiles twice as tall as wide (the warp-specialised S/CGEMM kernel: 256 x 128, complex 128 x 64), tm x tn of them over
 * a square window (tn = 2 tm or 2 tm - 1): tile (bm, bn) has an element in the lower triangle iff bn <= 2 bm + 1, in
 * the upper one iff bn >= 2 bm.  The lower triangle is enumerated row by row (row r holds min(tn, 2 r + 2) tiles,
 * r (r + 1) before it), the upper one column by column (column c holds c / 2 + 1 tiles; q (q + 1) before column 2 q,
 * (q + 1)^2 before column 2 q + 1).  Walking ALL tiles and skipping, as round 1 did for rectangular tiles, leaves the
 * static round-robin over persistent CTAs unbalanced by a tile or two per CTA out of seven (SSYRK 8192: 47 TFLOP/s
 * against 61 for the GEMM). */
__host__ __device__ __forceinline__ int64_t tri21_tile_count(int tri, int64_t tm, int64_t tn) {
  if (tri == 1) { const int64_t full = tm * (tm + 1), over = 2 * tm - tn; return full - (over > 0 ? over : 0); }
  const int64_t q = tn / 2;
  return (tn & 1) ? (q + 1) * (q + 1) : q * (q + 1);
}
__host__ __device__ __forceinline__ void tri21_tile_coords(int64_t u, int tri, int64_t &bm, int64_t &bn) {
  int64_t q = (int64_t)((sqrt(4.0 * (double)u + 1.0) - 1.0) * 0.5);
  while (q * (q + 1) > u) q--;
  while ((q + 1) * (q + 2) <= u) q++;
  if (tri == 1) { bm = q; bn = u - q * (q + 1); return; }
  if (u >= (q + 1) * (q + 1)) { bn = 2 * q + 1; bm = u - (q + 1) * (q + 1); } else { bn = 2 * q; bm = u - q * (q + 1); }
}
/* tiles twice as WIDE as tall (the ZGEMM kernel: 64 x 128): the transposed picture of the above */
__host__ __device__ __forceinline__ int64_t tri12_tile_count(int tri, int64_t tm, int64_t tn) { return tri21_tile_count(3 - tri, tn, tm); }
__host__ __device__ __forceinline__ void tri12_tile_coords(int64_t u, int tri, int64_t &bm, int64_t &bn) { tri21_tile_coords(u, 3 - tri, bn, bm); }
__host__ __device__ __forceinline__ bool tri_keep(int tri, int64_t m, int64_t n) {
  return tri == 1 ? m >= n : tri == 2 ? m <= n : true;
}

/* per-precision element traits */
template <int DT> struct Traits;
template <> struct Traits<B200_S> {
  using In = float; using Out = float; using Real = float;
  static constexpr bool kComplex = false;
};
template <> struct Traits<B200_D> {
  using In = double; using Out = double; using Real = double;
  static constexpr bool kComplex = false;
};
template <> struct Traits<B200_C> {
  using In = float2; using Out = float2; using Real = float;
  static constexpr bool kComplex = true;
};
template <> struct Traits<B200_Z> {
  using In = double2; using Out = double2; using Real = double;
  static constexpr bool kComplex = true;
};
template <> struct Traits<B200_SB> {
  using In = uint16_t; using Out = float; using Real = float;
  static constexpr bool kComplex = false;
};

/* every kernel launch of the library goes through this counter (b200_launch_count) */
void count_launch(const char *kernel_name);

/* kernel family launchers: return cudaSuccess, or cudaErrorNotSupported when the family
 * cannot take this problem (shape / alignment) and the dispatcher should fall through. */
cudaError_t launch_generic(const DeviceGemm &g, cudaStream_t stream);
cudaError_t launch_dgemm_dmma(const DeviceGemm &g, cudaStream_t stream);
cudaError_t launch_zgemm_dmma(const DeviceGemm &g, cudaStream_t stream);
cudaError_t launch_sgemm_ffma(const DeviceGemm &g, cudaStream_t stream);
cudaError_t launch_cgemm_ffma(const DeviceGemm &g, cudaStream_t stream);
cudaError_t launch_sbgemm_tcgen05(const DeviceGemm &g, cudaStream_t stream);
/* gemm_batch: one launch over every tile of every problem (device arrays; first_tile has count + 1 entries) */
int64_t generic_tile_count(int64_t m, int64_t n);
cudaError_t launch_grouped(int dtype, const DeviceGemm *problems_dev, const int64_t *first_tile_dev, int count, int64_t tiles,
                           cudaStream_t stream);
cudaError_t launch_convert(int dir, int64_t n, const void *in, int64_t inc_in, void *out,
                           int64_t inc_out, cudaStream_t stream);

/* bf16_level12.cu: SBGEMV / SBDOT on device pointers (x, y at their logical first element) */
size_t sbgemv_workspace_bytes(int trans, int64_t m, int64_t n);
cudaError_t launch_sbgemv(int trans, int64_t m, int64_t n, float alpha, const void *a, int64_t lda, const void *x, int64_t incx,
                          float beta, void *y, int64_t incy, void *workspace, cudaStream_t stream);
size_t sbdot_workspace_bytes();
cudaError_t launch_sbdot(int64_t n, const void *x, int64_t incx, const void *y, int64_t incy, void *workspace, float *result,
                         cudaStream_t stream);

/* level3_aux.cu: helpers of the symmetric level-3 family */
/* the region rows [i0, i0 + nr) x columns [j0, j0 + nc) of the full matrix (nr / nc < 0: everything) */
cudaError_t launch_expand_symmetric(int dtype, int uplo, int herm, int64_t n, const void *a, int64_t lda, void *out,
                                    int64_t ldo, cudaStream_t stream, int64_t i0 = 0, int64_t nr = -1, int64_t j0 = 0, int64_t nc = -1);
cudaError_t launch_tri_merge(int dtype, int uplo, int herm, int64_t n, const void *t, int64_t ldt, double beta_re,
                             double beta_im, void *c, int64_t ldc, cudaStream_t stream);

cudaError_t launch_real_diagonal(int dtype, int64_t n, void *c, int64_t ldc, cudaStream_t stream);
/* GEMM3M (level3_aux.cu): the stored rows x cols complex matrix x -> three real planes of pitch ldp: re, sign * im and
 * their sum (conj: sign = -1); and C := alpha * ((t1 - t2) + i (t3 - t1 - t2)) + beta * C from the three real products */
cudaError_t launch_split3(int dtype, int64_t rows, int64_t cols, const void *x, int64_t ldx, int conj, void *re, void *im, void *sum,
                          int64_t ldp, cudaStream_t stream);
cudaError_t launch_combine3(int dtype, int64_t m, int64_t n, const void *t1, const void *t2, const void *t3, int64_t ldt, double alpha_re,
                            double alpha_im, double beta_re, double beta_im, void *c, int64_t ldc, cudaStream_t stream);
/* base case of the recursive TRMM / TRSM: one nb x nb (nb <= tri_block_max(dtype)) triangular block against nrhs vectors */
int tri_block_max(int dtype);
cudaError_t launch_tri_block(int dtype, int solve, int nb, int64_t nrhs, int eff_lower, int unit, int cj, const void *f, int64_t fs_i,
                             int64_t fs_k, double ar, double ai, void *b, int64_t rs, int64_t cs, cudaStream_t stream);

int sm_count();

}  // namespace b200
