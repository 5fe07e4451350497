/*
 * sim_kernels.cpp -- TEST INFRASTRUCTURE.  CPU stand-ins for the kernel launchers declared in
 * openblas_b200/csrc/gemm_common.cuh, used only by the host-simulation build of the library
 * (tests/hostsim, tests/test_hostsim.py): every GEMM "launch" is the CPU oracle (oracle/gemm_oracle.c,
 * bit-identical to the reference's GENERIC build), the helper kernels of the level-3 family are plain
 * loops with the semantics documented in level3_aux.cu.  The launchers keep the ELIGIBILITY rules of
 * the real ones (alignment, tri support, size limits) so that the dispatcher and the fallbacks of
 * runtime.cu / runtime_level3.inl take the same decisions as on the GPU.
 */
#include <complex>
#include <cstring>
#include <vector>
#include "gemm_common.cuh"

extern "C" {
int oracle_gemm(int dtype, int transa, int transb, long m, long n, long k, const void *alpha, const void *a, long lda, const void *b,
                long ldb, const void *beta, void *c, long ldc, long q, long unroll_m);
void oracle_tobf16(long n, const float *in, long inc_in, uint16_t *out, long inc_out);
void oracle_bf16to(long n, const uint16_t *in, long inc_in, float *out, long inc_out);
uint16_t oracle_f32_to_bf16(float f);
int oracle_sbgemv(int trans, long m, long n, float alpha, const unsigned short *a, long lda, const unsigned short *x, long incx,
                  float beta, float *y, long incy, double *gauge);
double oracle_sbdot(long n, const unsigned short *x, long incx, const unsigned short *y, long incy, double *gauge);
float oracle_bf16_to_f32(uint16_t h);
}

namespace b200 {
namespace {

size_t out_size(int dt) { return dt == B200_S ? 4 : dt == B200_D ? 8 : dt == B200_C ? 8 : dt == B200_Z ? 16 : 4; }

void scalars(const DeviceGemm &g, double a64[2], float a32[2], double b64[2], float b32[2], const void **al, const void **be) {
  a64[0] = g.alpha_re; a64[1] = g.alpha_im; b64[0] = g.beta_re; b64[1] = g.beta_im;
  a32[0] = (float)g.alpha_re; a32[1] = (float)g.alpha_im; b32[0] = (float)g.beta_re; b32[1] = (float)g.beta_im;
  const bool dbl = g.dtype == B200_D || g.dtype == B200_Z;
  *al = dbl ? (const void *)a64 : (const void *)a32;
  *be = dbl ? (const void *)b64 : (const void *)b32;
}

cudaError_t sim_gemm(const DeviceGemm &g, const char *name) {
  double a64[2], b64[2]; float a32[2], b32[2]; const void *al, *be;
  scalars(g, a64, a32, b64, b32, &al, &be);
  if (!g.tri) {
    oracle_gemm(g.dtype, g.transa, g.transb, g.m, g.n, g.k, al, g.a, g.lda, g.b, g.ldb, be, g.c, g.ldc, 0, 0);
  } else {   /* full product into a scratch copy, then only the triangle goes back */
    const size_t es = out_size(g.dtype);
    std::vector<char> t((size_t)g.m * (size_t)g.n * es);
    for (int64_t j = 0; j < g.n; j++) memcpy(&t[(size_t)j * g.m * es], (const char *)g.c + (size_t)j * g.ldc * es, (size_t)g.m * es);
    oracle_gemm(g.dtype, g.transa, g.transb, g.m, g.n, g.k, al, g.a, g.lda, g.b, g.ldb, be, t.data(), g.m, 0, 0);
    for (int64_t j = 0; j < g.n; j++)
      for (int64_t i = 0; i < g.m; i++)
        if (tri_keep(g.tri, i, j)) memcpy((char *)g.c + ((size_t)i + (size_t)j * g.ldc) * es, &t[((size_t)i + (size_t)j * g.m) * es], es);
  }
  if (name) count_launch(name);
  return cudaSuccess;
}

template <class T> T conj_if(T v, bool) { return v; }
template <class R> std::complex<R> conj_if(std::complex<R> v, bool on) { return on ? std::conj(v) : v; }
template <class T> T real_part_only(T v) { return v; }
template <class R> std::complex<R> real_part_only(std::complex<R> v) { return std::complex<R>(v.real(), 0); }

template <class T> void expand(int uplo, int herm, const T *a, int64_t lda, T *out, int64_t ldo, int64_t i0, int64_t nr, int64_t j0, int64_t nc) {
  for (int64_t j = j0; j < j0 + nc; j++)
    for (int64_t i = i0; i < i0 + nr; i++) {
      const bool stored = uplo ? (i >= j) : (i <= j);
      T v = stored ? a[i + j * lda] : a[j + i * lda];
      if (herm) v = (i == j) ? real_part_only(v) : conj_if(v, !stored);
      out[(i - i0) + (j - j0) * ldo] = v;
    }
}
template <class T> void merge(int uplo, int herm, int64_t n, const T *t, int64_t ldt, T beta, T *c, int64_t ldc) {
  for (int64_t j = 0; j < n; j++)
    for (int64_t i = 0; i < n; i++) {
      if (uplo ? (i < j) : (i > j)) continue;
      T v = t ? t[i + j * ldt] : T(0);
      if (beta != T(0)) v = v + beta * c[i + j * ldc];
      if (herm && i == j) v = real_part_only(v);
      c[i + j * ldc] = v;
    }
}
template <class T> void tri_block(int solve, int nb, int64_t nrhs, int eff_lower, int unit, int cj, const T *f, int64_t fs_i, int64_t fs_k,
                                  T alpha, T *b, int64_t rs, int64_t cs) {
  std::vector<T> E((size_t)nb * nb, T(0));
  for (int i = 0; i < nb; i++)
    for (int k = 0; k < nb; k++) {
      T v = T(0);
      if (i == k) v = unit ? T(1) : f[i * fs_i + k * fs_k];
      else if (eff_lower ? (k < i) : (k > i)) v = f[i * fs_i + k * fs_k];
      E[(size_t)i * nb + k] = conj_if(v, cj != 0);
    }
  std::vector<T> x(nb), y(nb);
  for (int64_t c = 0; c < nrhs; c++) {
    for (int r = 0; r < nb; r++) x[r] = b[r * rs + c * cs];
    if (!solve) {
      for (int i = 0; i < nb; i++) {
        T acc = T(0);
        for (int k = eff_lower ? 0 : i; k < (eff_lower ? i + 1 : nb); k++) acc += E[(size_t)i * nb + k] * x[k];
        y[i] = alpha * acc;
      }
      x = y;
    } else {
      for (int ii = 0; ii < nb; ii++) {
        const int i = eff_lower ? ii : nb - 1 - ii;
        T acc = alpha * x[i];
        for (int k = eff_lower ? 0 : i + 1; k < (eff_lower ? i : nb); k++) acc -= E[(size_t)i * nb + k] * x[k];
        x[i] = unit ? acc : acc / E[(size_t)i * nb + i];
      }
    }
    for (int r = 0; r < nb; r++) b[r * rs + c * cs] = x[r];
  }
}

}  // namespace

cudaError_t launch_generic(const DeviceGemm &g, cudaStream_t) { return g.tri ? cudaErrorNotSupported : sim_gemm(g, "sim_generic"); }
int64_t generic_tile_count(int64_t m, int64_t n) { return ((m + 31) / 32) * ((n + 31) / 32); }
/* the grouped kernel: the problem list and tile prefix sums are read from "device" memory, as the real kernel does */
cudaError_t launch_grouped(int dtype, const DeviceGemm *problems, const int64_t *first_tile, int count, int64_t tiles, cudaStream_t) {
  if (first_tile[count] != tiles) return cudaErrorInvalidValue;
  for (int i = 0; i < count; i++) {
    if (problems[i].dtype != dtype) return cudaErrorInvalidValue;
    if (first_tile[i + 1] == first_tile[i]) continue;            /* no tiles: C untouched */
    if (first_tile[i + 1] - first_tile[i] != generic_tile_count(problems[i].m, problems[i].n)) return cudaErrorInvalidValue;
    sim_gemm(problems[i], nullptr);
  }
  count_launch("sim_grouped");
  return cudaSuccess;
}
cudaError_t launch_dgemm_dmma(const DeviceGemm &g, cudaStream_t) {
  if (g.dtype != B200_D || (((uintptr_t)g.a | (uintptr_t)g.b | (uintptr_t)g.c) & 7)) return cudaErrorNotSupported;
  const bool bulk_ok = g.lda % 2 == 0 && g.ldb % 2 == 0 && ((((uintptr_t)g.a | (uintptr_t)g.b) & 15) == 0);
  if (g.tri && (g.m != g.n || !bulk_ok)) return cudaErrorNotSupported;
  return sim_gemm(g, bulk_ok ? "sim_dgemm_producer_warp" : "sim_dgemm_cp_async");
}
cudaError_t launch_zgemm_dmma(const DeviceGemm &g, cudaStream_t) {
  if (g.dtype != B200_Z || (((uintptr_t)g.a | (uintptr_t)g.b | (uintptr_t)g.c) & 15)) return cudaErrorNotSupported;
  return sim_gemm(g, "sim_zgemm");
}
cudaError_t launch_sgemm_ffma(const DeviceGemm &g, cudaStream_t) {
  if (g.dtype != B200_S || (((uintptr_t)g.a | (uintptr_t)g.b | (uintptr_t)g.c) & 3)) return cudaErrorNotSupported;
  if (g.tri && g.m != g.n) return cudaErrorNotSupported;
  return sim_gemm(g, "sim_sgemm");
}
cudaError_t launch_cgemm_ffma(const DeviceGemm &g, cudaStream_t) {
  if (g.dtype != B200_C || (((uintptr_t)g.a | (uintptr_t)g.b | (uintptr_t)g.c) & 7)) return cudaErrorNotSupported;
  if (g.tri && g.m != g.n) return cudaErrorNotSupported;
  return sim_gemm(g, "sim_cgemm");
}
cudaError_t launch_sbgemm_tcgen05(const DeviceGemm &g, cudaStream_t) {
  /* the real launcher takes every shape and alignment (misaligned operands are repacked on the device first) */
  if (g.dtype != B200_SB || (g.tri && g.m != g.n) || ((uintptr_t)g.c & 3) || g.m < 1 || g.n < 1 || g.k < 1) return cudaErrorNotSupported;
  return sim_gemm(g, "sim_sbgemm");
}

cudaError_t launch_convert(int dir, int64_t n, const void *in, int64_t inc_in, void *out, int64_t inc_out, cudaStream_t) {
  for (int64_t i = 0; i < n; i++) {
    switch (dir) {
      case 0: ((uint16_t *)out)[i * inc_out] = oracle_f32_to_bf16(((const float *)in)[i * inc_in]); break;
      case 1: ((uint16_t *)out)[i * inc_out] = oracle_f32_to_bf16((float)((const double *)in)[i * inc_in]); break;
      case 2: ((float *)out)[i * inc_out] = oracle_bf16_to_f32(((const uint16_t *)in)[i * inc_in]); break;
      default: ((double *)out)[i * inc_out] = (double)oracle_bf16_to_f32(((const uint16_t *)in)[i * inc_in]); break;
    }
  }
  count_launch("sim_convert");
  return cudaSuccess;
}

cudaError_t launch_expand_symmetric(int dtype, int uplo, int herm, int64_t n, const void *a, int64_t lda, void *out, int64_t ldo, cudaStream_t,
                                    int64_t i0, int64_t nr, int64_t j0, int64_t nc) {
  if (nr < 0) { i0 = 0; nr = n; }
  if (nc < 0) { j0 = 0; nc = n; }
  if (nr <= 0 || nc <= 0) return cudaSuccess;
  switch (dtype) {
    case B200_S: expand<float>(uplo, 0, (const float *)a, lda, (float *)out, ldo, i0, nr, j0, nc); break;
    case B200_D: expand<double>(uplo, 0, (const double *)a, lda, (double *)out, ldo, i0, nr, j0, nc); break;
    case B200_C: expand<std::complex<float>>(uplo, herm, (const std::complex<float> *)a, lda, (std::complex<float> *)out, ldo, i0, nr, j0, nc); break;
    case B200_Z: expand<std::complex<double>>(uplo, herm, (const std::complex<double> *)a, lda, (std::complex<double> *)out, ldo, i0, nr, j0, nc); break;
    default: return cudaErrorNotSupported;
  }
  count_launch("sim_expand_symmetric");
  return cudaSuccess;
}
cudaError_t launch_tri_merge(int dtype, int uplo, int herm, int64_t n, const void *t, int64_t ldt, double br, double bi, void *c, int64_t ldc,
                             cudaStream_t) {
  if (n <= 0) return cudaSuccess;
  switch (dtype) {
    case B200_S: merge<float>(uplo, 0, n, (const float *)t, ldt, (float)br, (float *)c, ldc); break;
    case B200_D: merge<double>(uplo, 0, n, (const double *)t, ldt, br, (double *)c, ldc); break;
    case B200_C: merge<std::complex<float>>(uplo, herm, n, (const std::complex<float> *)t, ldt, std::complex<float>((float)br, (float)bi), (std::complex<float> *)c, ldc); break;
    case B200_Z: merge<std::complex<double>>(uplo, herm, n, (const std::complex<double> *)t, ldt, std::complex<double>(br, bi), (std::complex<double> *)c, ldc); break;
    default: return cudaErrorNotSupported;
  }
  count_launch("sim_tri_merge");
  return cudaSuccess;
}
/* GEMM3M helpers: the same element-wise arithmetic as level3_aux.cu's split3 / combine3 kernels */
template <class R> void split3(int64_t rows, int64_t cols, const R *x, int64_t ldx, R sign, R *re, R *im, R *sum, int64_t ldp) {
  for (int64_t c = 0; c < cols; c++)
    for (int64_t r = 0; r < rows; r++) {
      const R vr = x[2 * (r + c * ldx)], vi = sign * x[2 * (r + c * ldx) + 1];
      re[r + c * ldp] = vr; im[r + c * ldp] = vi; sum[r + c * ldp] = vr + vi;
    }
}
template <class R> void combine3(int64_t m, int64_t n, const R *t1, const R *t2, const R *t3, int64_t ldt, R ar, R ai, R br, R bi, bool use_beta,
                                 R *c, int64_t ldc) {
  for (int64_t j = 0; j < n; j++)
    for (int64_t r = 0; r < m; r++) {
      const R a1 = t1[r + j * ldt], a2 = t2[r + j * ldt], a3 = t3[r + j * ldt];
      const R pr = a1 - a2, pi = (a3 - a1) - a2;
      R ox = ar * pr - ai * pi, oy = ar * pi + ai * pr;
      if (use_beta) {
        const R cx = c[2 * (r + j * ldc)], cy = c[2 * (r + j * ldc) + 1];
        ox += br * cx - bi * cy; oy += br * cy + bi * cx;
      }
      c[2 * (r + j * ldc)] = ox; c[2 * (r + j * ldc) + 1] = oy;
    }
}
cudaError_t launch_split3(int dtype, int64_t rows, int64_t cols, const void *x, int64_t ldx, int conj, void *re, void *im, void *sum, int64_t ldp,
                          cudaStream_t) {
  if (rows <= 0 || cols <= 0) return cudaSuccess;
  if (dtype == B200_C) split3<float>(rows, cols, (const float *)x, ldx, conj ? -1.f : 1.f, (float *)re, (float *)im, (float *)sum, ldp);
  else if (dtype == B200_Z) split3<double>(rows, cols, (const double *)x, ldx, conj ? -1.0 : 1.0, (double *)re, (double *)im, (double *)sum, ldp);
  else return cudaErrorNotSupported;
  count_launch("sim_split3");
  return cudaSuccess;
}
cudaError_t launch_combine3(int dtype, int64_t m, int64_t n, const void *t1, const void *t2, const void *t3, int64_t ldt, double ar, double ai,
                            double br, double bi, void *c, int64_t ldc, cudaStream_t) {
  if (m <= 0 || n <= 0) return cudaSuccess;
  const bool use_beta = !(br == 0.0 && bi == 0.0);
  if (dtype == B200_C) combine3<float>(m, n, (const float *)t1, (const float *)t2, (const float *)t3, ldt, (float)ar, (float)ai, (float)br, (float)bi, use_beta, (float *)c, ldc);
  else if (dtype == B200_Z) combine3<double>(m, n, (const double *)t1, (const double *)t2, (const double *)t3, ldt, ar, ai, br, bi, use_beta, (double *)c, ldc);
  else return cudaErrorNotSupported;
  count_launch("sim_combine3");
  return cudaSuccess;
}
cudaError_t launch_real_diagonal(int dtype, int64_t n, void *c, int64_t ldc, cudaStream_t) {
  if (dtype == B200_C) for (int64_t i = 0; i < n; i++) ((float *)c)[2 * (i + i * ldc) + 1] = 0.f;
  if (dtype == B200_Z) for (int64_t i = 0; i < n; i++) ((double *)c)[2 * (i + i * ldc) + 1] = 0.0;
  count_launch("sim_real_diagonal");
  return cudaSuccess;
}
int tri_block_max(int dtype) { return (dtype == B200_S || dtype == B200_D) ? 128 : 64; }
cudaError_t launch_tri_block(int dtype, int solve, int nb, int64_t nrhs, int eff_lower, int unit, int cj, const void *f, int64_t fs_i,
                             int64_t fs_k, double ar, double ai, void *b, int64_t rs, int64_t cs, cudaStream_t) {
  if (nb <= 0 || nrhs <= 0) return cudaSuccess;
  if (nb > tri_block_max(dtype)) return cudaErrorInvalidValue;
  switch (dtype) {
    case B200_S: tri_block<float>(solve, nb, nrhs, eff_lower, unit, 0, (const float *)f, fs_i, fs_k, (float)ar, (float *)b, rs, cs); break;
    case B200_D: tri_block<double>(solve, nb, nrhs, eff_lower, unit, 0, (const double *)f, fs_i, fs_k, ar, (double *)b, rs, cs); break;
    case B200_C: tri_block<std::complex<float>>(solve, nb, nrhs, eff_lower, unit, cj, (const std::complex<float> *)f, fs_i, fs_k, std::complex<float>((float)ar, (float)ai), (std::complex<float> *)b, rs, cs); break;
    case B200_Z: tri_block<std::complex<double>>(solve, nb, nrhs, eff_lower, unit, cj, (const std::complex<double> *)f, fs_i, fs_k, std::complex<double>(ar, ai), (std::complex<double> *)b, rs, cs); break;
    default: return cudaErrorNotSupported;
  }
  count_launch(solve ? "sim_tri_block_solve" : "sim_tri_block_multiply");
  return cudaSuccess;
}

/* SBGEMV / SBDOT: the oracle on the staged operands.  The launchers get x and y at their LOGICAL first element;
 * the oracle, like the interface, takes the lowest address and moves it itself -- undo the move for it. */
size_t sbgemv_workspace_bytes(int, int64_t, int64_t) { return 256; }
cudaError_t launch_sbgemv(int trans, int64_t m, int64_t n, float alpha, const void *a, int64_t lda, const void *x, int64_t incx, float beta,
                          void *y, int64_t incy, void *, cudaStream_t) {
  const int64_t lenx = trans ? m : n, leny = trans ? n : m;
  const unsigned short *x0 = (const unsigned short *)x + (incx < 0 ? (lenx - 1) * incx : 0);
  float *y0 = (float *)y + (incy < 0 ? (leny - 1) * incy : 0);
  oracle_sbgemv(trans, m, n, alpha, (const unsigned short *)a, lda, x0, incx, beta, y0, incy, nullptr);
  count_launch("sim_sbgemv");
  return cudaSuccess;
}
size_t sbdot_workspace_bytes() { return 256; }
cudaError_t launch_sbdot(int64_t n, const void *x, int64_t incx, const void *y, int64_t incy, void *, float *result, cudaStream_t) {
  const unsigned short *x0 = (const unsigned short *)x + (incx < 0 ? (n - 1) * incx : 0);
  const unsigned short *y0 = (const unsigned short *)y + (incy < 0 ? (n - 1) * incy : 0);
  *result = (float)oracle_sbdot(n, x0, incx, y0, incy, nullptr);
  count_launch("sim_sbdot");
  return cudaSuccess;
}

}  // namespace b200
