/*
 * level3_aux.cu -- the two HBM-bound helper kernels behind SYMM/HEMM and the SYRK family
 * (runtime_level3.inl).  Both are plain coalesced passes over an n x n matrix: 32 x 8 thread
 * blocks walk columns (the contiguous direction of column-major storage) so every warp reads and
 * writes 128 / 256 / 512 contiguous bytes; the grid is sized to the matrix, capped at a multiple
 * of the SM count with a grid-stride loop.
 *
 *   expand_symmetric: out = the full matrix of a symmetric / Hermitian operand of which only
 *     one triangle is referenced (what the reference's symm_?copy / hemm_?copy packers do
 *     panel by panel: kernel/generic/symm_ucopy_*.c, zhemm_utcopy_*.c -- the Hermitian ones
 *     also force the diagonal's imaginary part to zero).
 *   tri_merge: C(tri) = T + beta * C(tri) on one triangle only, beta == 0 never reads C
 *     (driver/level3/syrk_kernel.c writes only the triangle part of a diagonal block through
 *     a scratch tile; syrk_k.c's syrk_beta scales only the triangle); Hermitian flavour zeroes
 *     the diagonal's imaginary part (zherk_kernel.c, zherk_beta.c).  T == nullptr means "0".
 */
#include "gemm_common.cuh"

namespace b200 {
namespace {

template <class T> struct Cx { static constexpr bool value = false; };
template <> struct Cx<float2> { static constexpr bool value = true; };
template <> struct Cx<double2> { static constexpr bool value = true; };

template <class T> __device__ __forceinline__ T conj_of(T v) { return v; }
template <> __device__ __forceinline__ float2 conj_of(float2 v) { return make_float2(v.x, -v.y); }
template <> __device__ __forceinline__ double2 conj_of(double2 v) { return make_double2(v.x, -v.y); }
template <class T> __device__ __forceinline__ T real_only(T v) { return v; }
template <> __device__ __forceinline__ float2 real_only(float2 v) { return make_float2(v.x, 0.f); }
template <> __device__ __forceinline__ double2 real_only(double2 v) { return make_double2(v.x, 0.0); }

template <class T>
__global__ void __launch_bounds__(256) expand_symmetric_kernel(int uplo, int herm, int64_t n, const T *__restrict__ a,
                                                               int64_t lda, T *__restrict__ out, int64_t ldo) {
  const int64_t tiles_i = (n + 31) / 32, tiles_j = (n + 7) / 8;
  for (int64_t t = blockIdx.x; t < tiles_i * tiles_j; t += gridDim.x) {
    const int64_t i = (t % tiles_i) * 32 + threadIdx.x, j = (t / tiles_i) * 8 + threadIdx.y;
    if (i >= n || j >= n) continue;
    const bool stored = uplo ? (i >= j) : (i <= j);       /* lower: on or below the diagonal */
    T v = stored ? a[i + j * lda] : a[j + i * lda];
    if (herm) v = (i == j) ? real_only(v) : (stored ? v : conj_of(v));
    out[i + j * ldo] = v;
  }
}

template <class T, class R>
__device__ __forceinline__ T axpby(T t, R br, R bi, T c) {
  if constexpr (Cx<T>::value) {
    T r;
    r.x = t.x + (br * c.x - bi * c.y);
    r.y = t.y + (br * c.y + bi * c.x);
    return r;
  } else {
    return t + br * c;
  }
}

template <class T, class R>
__global__ void __launch_bounds__(256) tri_merge_kernel(int uplo, int herm, int64_t n, const T *__restrict__ t_, int64_t ldt,
                                                        R br, R bi, T *__restrict__ c, int64_t ldc) {
  const int64_t tiles_i = (n + 31) / 32, tiles_j = (n + 7) / 8;
  const bool use_beta = !(br == R(0) && bi == R(0));
  for (int64_t t = blockIdx.x; t < tiles_i * tiles_j; t += gridDim.x) {
    const int64_t i = (t % tiles_i) * 32 + threadIdx.x, j = (t / tiles_i) * 8 + threadIdx.y;
    if (i >= n || j >= n) continue;
    if (uplo ? (i < j) : (i > j)) continue;               /* outside the named triangle: untouched */
    T v;
    memset(&v, 0, sizeof v);
    if (t_) v = t_[i + j * ldt];
    if (use_beta) v = axpby<T, R>(v, br, bi, c[i + j * ldc]);
    if (herm && i == j) v = real_only(v);
    c[i + j * ldc] = v;
  }
}

template <class T>
cudaError_t expand_t(int uplo, int herm, int64_t n, const void *a, int64_t lda, void *out, int64_t ldo, cudaStream_t s) {
  const int64_t tiles = ((n + 31) / 32) * ((n + 7) / 8);
  const int64_t cap = (int64_t)sm_count() * 16;
  expand_symmetric_kernel<T><<<(unsigned)(tiles < cap ? tiles : cap), dim3(32, 8), 0, s>>>(uplo, herm, n, (const T *)a, lda, (T *)out, ldo);
  return cudaGetLastError();
}
template <class T, class R>
cudaError_t merge_t(int uplo, int herm, int64_t n, const void *t, int64_t ldt, double br, double bi, void *c, int64_t ldc,
                    cudaStream_t s) {
  const int64_t tiles = ((n + 31) / 32) * ((n + 7) / 8);
  const int64_t cap = (int64_t)sm_count() * 16;
  tri_merge_kernel<T, R><<<(unsigned)(tiles < cap ? tiles : cap), dim3(32, 8), 0, s>>>(uplo, herm, n, (const T *)t, ldt, (R)br, (R)bi, (T *)c, ldc);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_expand_symmetric(int dtype, int uplo, int herm, int64_t n, const void *a, int64_t lda, void *out,
                                    int64_t ldo, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  cudaError_t e;
  switch (dtype) {
    case B200_S: e = expand_t<float>(uplo, 0, n, a, lda, out, ldo, stream); break;
    case B200_D: e = expand_t<double>(uplo, 0, n, a, lda, out, ldo, stream); break;
    case B200_C: e = expand_t<float2>(uplo, herm, n, a, lda, out, ldo, stream); break;
    case B200_Z: e = expand_t<double2>(uplo, herm, n, a, lda, out, ldo, stream); break;
    default: return cudaErrorNotSupported;
  }
  if (e == cudaSuccess) count_launch("expand_symmetric");
  return e;
}

cudaError_t launch_tri_merge(int dtype, int uplo, int herm, int64_t n, const void *t, int64_t ldt, double beta_re,
                             double beta_im, void *c, int64_t ldc, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  cudaError_t e;
  switch (dtype) {
    case B200_S: e = merge_t<float, float>(uplo, 0, n, t, ldt, beta_re, 0.0, c, ldc, stream); break;
    case B200_D: e = merge_t<double, double>(uplo, 0, n, t, ldt, beta_re, 0.0, c, ldc, stream); break;
    case B200_C: e = merge_t<float2, float>(uplo, herm, n, t, ldt, beta_re, beta_im, c, ldc, stream); break;
    case B200_Z: e = merge_t<double2, double>(uplo, herm, n, t, ldt, beta_re, beta_im, c, ldc, stream); break;
    default: return cudaErrorNotSupported;
  }
  if (e == cudaSuccess) count_launch("tri_merge");
  return e;
}

}  // namespace b200
