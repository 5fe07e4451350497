/*
 * level3_aux.cu -- the helper kernels behind the rest of level 3 (runtime_level3.inl) and GEMM3M: HBM-bound
 * passes (expand_symmetric, tri_merge, real_diagonal, split3, combine3) and the triangular block kernels of
 * TRMM / TRSM (tri_block_kernel, tri_block_reg_kernel, described where they are defined).  The first two are plain coalesced
 * passes over an n x n matrix: 32 x 8 thread
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
 *   split3 / combine3: the element-wise passes around the three real products of GEMM3M (runtime.cu:
 *     gemm3m_on_device; what kernel/generic/zgemm3m_ncopy_*.c / zgemm3m_tcopy_*.c do while packing in the reference).
 */
#include <cstdlib>
#include <cstring>
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

/* out(i - i0, j - j0) = full(i, j) for the region rows [i0, i0 + nr) x columns [j0, j0 + nc) of the n x n matrix */
template <class T>
__global__ void __launch_bounds__(256) expand_symmetric_kernel(int uplo, int herm, const T *__restrict__ a, int64_t lda,
                                                               T *__restrict__ out, int64_t ldo, int64_t i0, int64_t nr, int64_t j0, int64_t nc) {
  const int64_t tiles_i = (nr + 31) / 32, tiles_j = (nc + 7) / 8;
  for (int64_t t = blockIdx.x; t < tiles_i * tiles_j; t += gridDim.x) {
    const int64_t li = (t % tiles_i) * 32 + threadIdx.x, lj = (t / tiles_i) * 8 + threadIdx.y;
    if (li >= nr || lj >= nc) continue;
    const int64_t i = i0 + li, j = j0 + lj;
    const bool stored = uplo ? (i >= j) : (i <= j);       /* lower: on or below the diagonal */
    T v = stored ? a[i + j * lda] : a[j + i * lda];
    if (herm) v = (i == j) ? real_only(v) : (stored ? v : conj_of(v));
    out[li + lj * ldo] = v;
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

/* ---------------------------------------------------------------------------------------------
 * tri_block_kernel: the base case of the recursive TRMM / TRSM (runtime_level3.inl).  One CTA takes
 * one nb x nb (nb <= 64) diagonal block E and up to 32 right-hand sides, both staged in shared memory:
 *     SOLVE = false   x := alpha * E x         (rows bottom-up for a lower E, so it works in place)
 *     SOLVE = true    x := E^-1 (alpha * x)    by forward / backward substitution -- the reference's
 *                     trsm kernels substitute too (kernel/generic/trsm_kernel_LN.c `solve`), no
 *                     explicit inverse of the block is formed
 * E(i,k) = cj(F[i*fs_i + k*fs_k]) lets the host describe op(F) (left side) or op(F)^T (right side:
 * X E = alpha B  <=>  E^T X^T = alpha B^T); element r of right-hand side c lives at B[r*rs + c*cs].
 * Loads and stores walk whichever index is contiguous in memory.  E(i,k) reads are warp broadcasts,
 * X(k, thread) reads hit consecutive banks. */
template <class T> __device__ __forceinline__ T t_zero() { T v; memset(&v, 0, sizeof v); return v; }
template <class T> __device__ __forceinline__ T t_one() { T v; memset(&v, 0, sizeof v); *reinterpret_cast<decltype(v.x) *>(&v) = 1; return v; }
template <> __device__ __forceinline__ float t_one<float>() { return 1.f; }
template <> __device__ __forceinline__ double t_one<double>() { return 1.0; }
__device__ __forceinline__ float t_mul(float a, float b) { return a * b; }
__device__ __forceinline__ double t_mul(double a, double b) { return a * b; }
__device__ __forceinline__ float2 t_mul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 t_mul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float t_add(float a, float b) { return a + b; }
__device__ __forceinline__ double t_add(double a, double b) { return a + b; }
__device__ __forceinline__ float2 t_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 t_add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float t_sub(float a, float b) { return a - b; }
__device__ __forceinline__ double t_sub(double a, double b) { return a - b; }
__device__ __forceinline__ float2 t_sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 t_sub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float t_div(float a, float b) { return a / b; }
__device__ __forceinline__ double t_div(double a, double b) { return a / b; }
__device__ __forceinline__ float2 t_div(float2 a, float2 b) {
  float d = b.x * b.x + b.y * b.y;
  return make_float2((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d);
}
__device__ __forceinline__ double2 t_div(double2 a, double2 b) {
  double d = b.x * b.x + b.y * b.y;
  return make_double2((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d);
}
template <class T, class R> __device__ __forceinline__ T t_scalar(R re, R im) {
  if constexpr (Cx<T>::value) { T v; v.x = re; v.y = im; return v; } else { return (T)re; }
}

constexpr int TRI_NB = 64;      /* largest diagonal block */
constexpr int TRI_LANES = 8;    /* lanes that share one right-hand side (the dot products are split over them) */
constexpr int TRI_RHS = 32;     /* right-hand sides per CTA: 256 threads */
constexpr int TRI_LDE = TRI_NB + 1;
constexpr int TRI_LDX = TRI_NB + 8;   /* X[rhs][k], stride = 8 (mod 32) elements: lane (rhs, q) reads word rhs*8 + q of a bank row */

__device__ __forceinline__ float t_shfl_xor(float v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ double t_shfl_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ float2 t_shfl_xor(float2 v, int m) { return make_float2(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m)); }
__device__ __forceinline__ double2 t_shfl_xor(double2 v, int m) { return make_double2(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m)); }

/* 256 threads: right-hand side c = thread / 8; the eight lanes of a group take the k = q, q+8, ... terms
 * of every dot product and combine them with three shuffles.  The substitution is a chain of nb
 * dependent steps, so what counts is the length of one step (ncu: one thread per right-hand side
 * spent 66 us per 64 x 64 block, IPC 0.2, stalled on shared memory -- a third of a whole DTRSM):
 * all loads of a row are issued together (fixed trip count, predicated), X is laid out so that
 * they are conflict free, and the diagonal is inverted once while E is loaded, as the reference's
 * packers do (kernel/generic/trsm_ltcopy_4.c stores INV(a)), so no division sits in the chain. */
template <class T, class R, bool SOLVE>
__global__ void __launch_bounds__(TRI_RHS * TRI_LANES) tri_block_kernel(int nb, int64_t nrhs, int eff_lower, int unit, int cj,
                                                                        const T *__restrict__ f, int64_t fs_i, int64_t fs_k, R ar, R ai,
                                                                        T *__restrict__ b, int64_t rs, int64_t cs) {
  extern __shared__ __align__(16) unsigned char tri_smem[];
  constexpr int NT = TRI_RHS * TRI_LANES;
  T *E = reinterpret_cast<T *>(tri_smem);
  T *X = E + TRI_NB * TRI_LDE;
  const int t = threadIdx.x, c = t / TRI_LANES, q = t % TRI_LANES;
  const int64_t c0 = (int64_t)blockIdx.x * TRI_RHS;
  if (nb < TRI_NB)      /* rows / columns beyond nb are read (and discarded) by the unrolled loop: keep them finite */
    for (int idx = t; idx < TRI_NB * TRI_LDE + TRI_RHS * TRI_LDX; idx += NT) E[idx] = t_zero<T>();
  __syncthreads();
  for (int idx = t; idx < nb * nb; idx += NT) {
    const int i = fs_i == 1 ? idx % nb : idx / nb, k = fs_i == 1 ? idx / nb : idx % nb;
    T v = t_zero<T>();
    if (i == k) {
      v = unit ? t_one<T>() : f[i * fs_i + k * fs_k];
      if (SOLVE && !unit) v = t_div(t_one<T>(), v);         /* 1 / diagonal; conj(1/z) = 1/conj(z) */
    } else if (eff_lower ? (k < i) : (k > i)) {
      v = f[i * fs_i + k * fs_k];
    }
    if (cj) v = conj_of(v);
    E[i * TRI_LDE + k] = v;
  }
  for (int idx = t; idx < nb * TRI_RHS; idx += NT) {
    const int r = rs == 1 ? idx % nb : idx / TRI_RHS, cc = rs == 1 ? idx / nb : idx % TRI_RHS;
    X[cc * TRI_LDX + r] = (c0 + cc < nrhs) ? b[r * rs + (c0 + cc) * cs] : t_zero<T>();
  }
  __syncthreads();
  {   /* every lane takes part (shuffles); right-hand sides beyond nrhs hold zeros */
    const T alpha = t_scalar<T, R>(ar, ai);
    T *xc = X + c * TRI_LDX;
    for (int ii = 0; ii < nb; ii++) {
      /* row order: a product overwrites x_i after its last use, a solve needs the x_k it depends on */
      const int i = (SOLVE == (eff_lower != 0)) ? ii : nb - 1 - ii;
      const int k_lo = eff_lower ? 0 : (SOLVE ? i + 1 : i), k_hi = eff_lower ? (SOLVE ? i : i + 1) : nb;
      const T *ei = E + i * TRI_LDE;
      T acc = t_zero<T>(), acc2 = t_zero<T>();
#pragma unroll
      for (int j = 0; j < TRI_NB / TRI_LANES; j += 2) {
        const int k0 = q + j * TRI_LANES, k1 = k0 + TRI_LANES;
        const T e0 = ei[k0], x0 = xc[k0];                 /* k < 64 always: inside the arrays */
        const T e1 = ei[k1], x1 = xc[k1];
        if (k0 >= k_lo && k0 < k_hi) acc = t_add(acc, t_mul(e0, x0));
        if (k1 >= k_lo && k1 < k_hi) acc2 = t_add(acc2, t_mul(e1, x1));
      }
      acc = t_add(acc, acc2);
      acc = t_add(acc, t_shfl_xor(acc, 1));
      acc = t_add(acc, t_shfl_xor(acc, 2));
      acc = t_add(acc, t_shfl_xor(acc, 4));
      if (q == 0) {
        if (SOLVE) {
          const T v = t_sub(t_mul(alpha, xc[i]), acc);
          xc[i] = unit ? v : t_mul(v, ei[i]);
        } else {
          xc[i] = t_mul(alpha, acc);
        }
      }
      __syncwarp();
    }
  }
  __syncthreads();
  for (int idx = t; idx < nb * TRI_RHS; idx += NT) {
    const int r = rs == 1 ? idx % nb : idx / TRI_RHS, cc = rs == 1 ? idx / nb : idx % TRI_RHS;
    if (c0 + cc < nrhs) b[r * rs + (c0 + cc) * cs] = X[cc * TRI_LDX + r];
  }
}

/* Real types: the right-hand sides live in REGISTERS.  Round 1's kernel above keeps them in shared memory and spends
 * ~40 us per 64 x 64 block on a chain of 64 shuffle-reduced dot products (128 serial launches: a quarter of a DTRSM).
 * Here a GROUP of four adjacent lanes owns one right-hand side, 16 of its 64 entries each (row pairs 8q + 2h, 8q + 2h + 1
 * for lane h of the group), and the algorithm is column oriented: for pivot k the owner finishes x_k (times the
 * inverted diagonal), one shuffle hands it to the partner, and both apply the updates b_i -= e_ik x_k to their rows
 * below -- independent FMAs fed by LDS.128 loads of column k of E that every pair reads at the same addresses
 * (broadcast, conflict free) and that are all issued before the first FMA (32 registers of x leave room; a first
 * version with 64 entries per thread had none left and serialised on every load: slower than round 1).  Only the
 * 64 pivots form a dependent chain.  TRMM accumulates y = E x the same way into a second register set.  Blocks
 * smaller than 64 are padded with the identity.  Right-hand sides that are rows of B (right side, cs == 1) are read
 * and written straight from global memory, columns of B (rs == 1) through a shared-memory tile. */
/* two instantiations: 64 x 64 blocks with 4 lanes x 16 rows per right-hand side (16 right-hand sides per 64-thread
 * CTA), and 128 x 128 blocks with 8 lanes x 16 rows (64 per 512-thread CTA, so that 8192 right-hand sides are ONE wave of
 * 128 CTAs: the block is 128 KB of shared memory, one CTA per SM): the larger base case removes the deepest
 * level of the recursion -- 64 GEMMs with k = 64 that ran at 1.4 TFLOP/s (47 us each, profiles/r02_dtrsm8192_launches_*) */
template <class T> struct __align__(16) Pair2 { T x, y; };          /* float2 / double2 pairs: 16 / 32 bytes */
template <> struct __align__(16) Pair2<double> { double x, y; };
template <> struct __align__(8) Pair2<float> { float x, y; };
/* x - e * xk and y + e * xk in the working precision (complex: four real FMAs) */
__device__ __forceinline__ float t_fnma(float e, float xk, float x) { return fmaf(-e, xk, x); }
__device__ __forceinline__ double t_fnma(double e, double xk, double x) { return fma(-e, xk, x); }
__device__ __forceinline__ float2 t_fnma(float2 e, float2 xk, float2 x) {
  x.x = fmaf(-e.x, xk.x, x.x); x.x = fmaf(e.y, xk.y, x.x); x.y = fmaf(-e.x, xk.y, x.y); x.y = fmaf(-e.y, xk.x, x.y); return x;
}
__device__ __forceinline__ double2 t_fnma(double2 e, double2 xk, double2 x) {
  x.x = fma(-e.x, xk.x, x.x); x.x = fma(e.y, xk.y, x.x); x.y = fma(-e.x, xk.y, x.y); x.y = fma(-e.y, xk.x, x.y); return x;
}
__device__ __forceinline__ float t_fma(float e, float xk, float y) { return fmaf(e, xk, y); }
__device__ __forceinline__ double t_fma(double e, double xk, double y) { return fma(e, xk, y); }
__device__ __forceinline__ float2 t_fma(float2 e, float2 xk, float2 y) {
  y.x = fmaf(e.x, xk.x, y.x); y.x = fmaf(-e.y, xk.y, y.x); y.y = fmaf(e.x, xk.y, y.y); y.y = fmaf(e.y, xk.x, y.y); return y;
}
__device__ __forceinline__ double2 t_fma(double2 e, double2 xk, double2 y) {
  y.x = fma(e.x, xk.x, y.x); y.x = fma(-e.y, xk.y, y.x); y.y = fma(e.x, xk.y, y.y); y.y = fma(e.y, xk.x, y.y); return y;
}
__device__ __forceinline__ float t_shfl_idx(float v, unsigned l) { return __shfl_sync(0xffffffffu, v, l); }
__device__ __forceinline__ double t_shfl_idx(double v, unsigned l) { return __shfl_sync(0xffffffffu, v, l); }
__device__ __forceinline__ float2 t_shfl_idx(float2 v, unsigned l) { return make_float2(__shfl_sync(0xffffffffu, v.x, l), __shfl_sync(0xffffffffu, v.y, l)); }
__device__ __forceinline__ double2 t_shfl_idx(double2 v, unsigned l) { return make_double2(__shfl_sync(0xffffffffu, v.x, l), __shfl_sync(0xffffffffu, v.y, l)); }

template <class T, int N, int LP, int TRR_RHS, bool SOLVE, bool LOWER>
__global__ void __launch_bounds__(LP * TRR_RHS, (LP * TRR_RHS == 256 ? 2 : 1)) tri_block_reg_kernel(int nb, int64_t nrhs, int unit, int cj, const T *__restrict__ f, int64_t fs_i,
                                                                    int64_t fs_k, T alpha, T *__restrict__ b, int64_t rs, int64_t cs) {
  constexpr int TRR_THREADS = LP * TRR_RHS, CH = 2 * LP, NQ = N / CH, NL = N / LP;   /* chunk of rows, chunks, rows per lane */
  extern __shared__ __align__(16) unsigned char tri_smem[];
  T *E = reinterpret_cast<T *>(tri_smem);          /* E[k][i]: column k contiguous; SOLVE keeps 1 / e_kk on the diagonal */
  T *Xs = E + N * N;                               /* [rhs][N + 1] staging tile when the right-hand sides are columns of B */
  const int t = threadIdx.x, h = t % LP, rl = t / LP;
  const int64_t c0 = (int64_t)blockIdx.x * TRR_RHS, c = c0 + rl;
  /* the block: N * N / TRR_THREADS elements per thread, fetched in batches of 16 independent loads (one load per
   * loop trip would pay the global-memory latency 64 times over: that, not the arithmetic, was most of the kernel) */
  {
    constexpr int BATCH = 16;
    const bool col_major = fs_i == 1;               /* consecutive threads walk the contiguous direction of F */
#pragma unroll 1
    for (int base = 0; base < N * N; base += BATCH * TRR_THREADS) {
      T v[BATCH];
#pragma unroll
      for (int j = 0; j < BATCH; j++) {
        const int idx = base + j * TRR_THREADS + t;
        const int i = col_major ? idx % N : idx / N, k = col_major ? idx / N : idx % N;
        const bool diag = i == k, inside = i < nb && k < nb && (LOWER ? (k < i) : (k > i));
        v[j] = ((diag && !unit && i < nb) || inside) ? f[i * fs_i + k * fs_k] : (diag ? t_one<T>() : t_zero<T>());
        if (cj) v[j] = conj_of(v[j]);
      }
#pragma unroll
      for (int j = 0; j < BATCH; j++) {
        const int idx = base + j * TRR_THREADS + t;
        const int i = col_major ? idx % N : idx / N, k = col_major ? idx / N : idx % N;
        E[k * N + i] = v[j];
      }
    }
    if (SOLVE) {                                    /* the N divisions of the whole solve, off the pivot chain */
      __syncthreads();
      if (t < N) E[t * N + t] = t_div(t_one<T>(), E[t * N + t]);
    }
  }
  /* my rows: CH q + 2 h + e, kept at x[2 q + e] */
  T x[NL], y[SOLVE ? 1 : NL];
  if (cs == 1) {
#pragma unroll
    for (int l = 0; l < NL; l++) {
      const int r = CH * (l >> 1) + 2 * h + (l & 1);
      x[l] = (r < nb && c < nrhs) ? b[r * rs + c] : t_zero<T>();
    }
    __syncthreads();
  } else {
    {
      constexpr int PER = N * TRR_RHS / TRR_THREADS;        /* 16 independent loads per thread, then the stores */
      T v[PER];
#pragma unroll
      for (int j = 0; j < PER; j++) {
        const int idx = j * TRR_THREADS + t, r = idx % N, cc = idx / N;
        v[j] = (r < nb && c0 + cc < nrhs) ? b[r * rs + (c0 + cc) * cs] : t_zero<T>();
      }
#pragma unroll
      for (int j = 0; j < PER; j++) {
        const int idx = j * TRR_THREADS + t, r = idx % N, cc = idx / N;
        Xs[cc * (N + 1) + r] = v[j];
      }
    }
    __syncthreads();
#pragma unroll
    for (int l = 0; l < NL; l++) x[l] = Xs[rl * (N + 1) + CH * (l >> 1) + 2 * h + (l & 1)];
  }
  if (SOLVE) {
#pragma unroll
    for (int l = 0; l < NL; l++) x[l] = t_mul(x[l], alpha);
  } else {
#pragma unroll
    for (int l = 0; l < NL; l++) y[l] = t_zero<T>();
  }
  const T *Eh = E + 2 * h;
  const unsigned group_base = (unsigned)(t & 31) & ~(unsigned)(LP - 1);
  /* column k of E for my rows of the chunks that still take part; the loads of column k + 1 are issued before the
   * updates of column k (ptxas keeps only a few of them in flight; with 8 per column that is enough) */
  Pair2<T> ecur[NQ], enext[NQ];
  T dcur = t_one<T>(), dnext = t_one<T>();
  auto load_column = [&](Pair2<T> (&e)[NQ], T &d, int k) {
    const int kq = k / CH;
#pragma unroll
    for (int q = 0; q < NQ; q++)
      if (LOWER ? (q >= kq) : (q <= kq)) e[q] = *reinterpret_cast<const Pair2<T> *>(Eh + k * N + CH * q);
    if (SOLVE) d = E[k * N + k];                    /* 1 / diagonal (1 for a unit diagonal and for the padding) */
  };
  load_column(ecur, dcur, LOWER ? 0 : N - 1);
#pragma unroll
  for (int kk = 0; kk < N; kk++) {
    const int k = LOWER ? kk : N - 1 - kk;
    const int kq = k / CH, kh = (k % CH) >> 1, kl = 2 * kq + (k & 1);      /* chunk, owner lane of the group, owner's slot of row k */
    if (kk + 1 < N) load_column(enext, dnext, LOWER ? k + 1 : k - 1);
    T xk = x[kl];
    if (SOLVE) xk = t_mul(xk, dcur);
    xk = t_shfl_idx(xk, group_base | (unsigned)kh);
    if (SOLVE && h == kh) x[kl] = xk;
#pragma unroll
    for (int q = 0; q < NQ; q++) {
      if (!(LOWER ? (q >= kq) : (q <= kq))) continue;
      const int r0 = CH * q + 2 * h;                /* my two rows of this chunk */
      if (SOLVE) {
        /* rows strictly beyond the pivot; only the pivot's own chunk needs the test */
        const bool on0 = q != kq || (LOWER ? r0 > k : r0 < k), on1 = q != kq || (LOWER ? r0 + 1 > k : r0 + 1 < k);
        if (on0) x[2 * q] = t_fnma(ecur[q].x, xk, x[2 * q]);
        if (on1) x[2 * q + 1] = t_fnma(ecur[q].y, xk, x[2 * q + 1]);
      } else {
        const bool on0 = q != kq || (LOWER ? r0 >= k : r0 <= k), on1 = q != kq || (LOWER ? r0 + 1 >= k : r0 + 1 <= k);
        if (on0) y[2 * q] = t_fma(ecur[q].x, xk, y[2 * q]);
        if (on1) y[2 * q + 1] = t_fma(ecur[q].y, xk, y[2 * q + 1]);
      }
    }
#pragma unroll
    for (int q = 0; q < NQ; q++) ecur[q] = enext[q];
    dcur = dnext;
  }
  if (!SOLVE) {
#pragma unroll
    for (int l = 0; l < NL; l++) x[l] = t_mul(alpha, y[l]);
  }
  if (cs == 1) {
#pragma unroll
    for (int l = 0; l < NL; l++) {
      const int r = CH * (l >> 1) + 2 * h + (l & 1);
      if (r < nb && c < nrhs) b[r * rs + c] = x[l];
    }
  } else {
#pragma unroll
    for (int l = 0; l < NL; l++) Xs[rl * (N + 1) + CH * (l >> 1) + 2 * h + (l & 1)] = x[l];
    __syncthreads();
    for (int idx = t; idx < N * TRR_RHS; idx += TRR_THREADS) {
      const int r = idx % N, cc = idx / N;
      if (r < nb && c0 + cc < nrhs) b[r * rs + (c0 + cc) * cs] = Xs[cc * (N + 1) + r];
    }
  }
}

template <class T, int N, int LP, int RHS, bool SOLVE, bool LOWER>
cudaError_t tri_block_reg_launch(int nb, int64_t nrhs, int unit, int cj, const void *f, int64_t fs_i, int64_t fs_k, T alpha, void *b, int64_t rs,
                                 int64_t cs, cudaStream_t s) {
  static bool configured = false;
  auto kern = tri_block_reg_kernel<T, N, LP, RHS, SOLVE, LOWER>;
  const size_t smem = ((size_t)N * N + (size_t)RHS * (N + 1)) * sizeof(T);
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  kern<<<(unsigned)((nrhs + RHS - 1) / RHS), LP * RHS, smem, s>>>(nb, nrhs, unit, cj, (const T *)f, fs_i, fs_k, alpha, (T *)b, rs, cs);
  return cudaGetLastError();
}
template <class T, int N, int LP, int RHS>
cudaError_t tri_block_reg_n(int solve, int nb, int64_t nrhs, int eff_lower, int unit, int cj, const void *f, int64_t fs_i, int64_t fs_k, T alpha,
                            void *b, int64_t rs, int64_t cs, cudaStream_t s) {
  if (solve) return eff_lower ? tri_block_reg_launch<T, N, LP, RHS, true, true>(nb, nrhs, unit, cj, f, fs_i, fs_k, alpha, b, rs, cs, s)
                              : tri_block_reg_launch<T, N, LP, RHS, true, false>(nb, nrhs, unit, cj, f, fs_i, fs_k, alpha, b, rs, cs, s);
  return eff_lower ? tri_block_reg_launch<T, N, LP, RHS, false, true>(nb, nrhs, unit, cj, f, fs_i, fs_k, alpha, b, rs, cs, s)
                   : tri_block_reg_launch<T, N, LP, RHS, false, false>(nb, nrhs, unit, cj, f, fs_i, fs_k, alpha, b, rs, cs, s);
}
template <class T>
cudaError_t tri_block_reg(int solve, int nb, int64_t nrhs, int eff_lower, int unit, const void *f, int64_t fs_i, int64_t fs_k, double ar, void *b,
                          int64_t rs, int64_t cs, cudaStream_t s) {
  if (nb <= 64) return tri_block_reg_n<T, 64, 4, 16>(solve, nb, nrhs, eff_lower, unit, 0, f, fs_i, fs_k, (T)ar, b, rs, cs, s);
  return tri_block_reg_n<T, 128, 8, 64>(solve, nb, nrhs, eff_lower, unit, 0, f, fs_i, fs_k, (T)ar, b, rs, cs, s);
}
/* complex: 64 x 64 blocks, 8 lanes x 8 rows per right-hand side, 32 right-hand sides per 256-thread CTA (two CTAs per SM) */
template <class T, class R>
cudaError_t tri_block_reg_cplx(int solve, int nb, int64_t nrhs, int eff_lower, int unit, int cj, const void *f, int64_t fs_i, int64_t fs_k, double ar,
                               double ai, void *b, int64_t rs, int64_t cs, cudaStream_t s) {
  T alpha; alpha.x = (R)ar; alpha.y = (R)ai;
  return tri_block_reg_n<T, 64, 8, 32>(solve, nb, nrhs, eff_lower, unit, cj, f, fs_i, fs_k, alpha, b, rs, cs, s);
}

template <class T, class R, bool SOLVE>
cudaError_t tri_block_t(int nb, int64_t nrhs, int eff_lower, int unit, int cj, const void *f, int64_t fs_i, int64_t fs_k, double ar,
                        double ai, void *b, int64_t rs, int64_t cs, cudaStream_t s) {
  static bool configured = false;
  auto kern = tri_block_kernel<T, R, SOLVE>;
  const size_t smem = ((size_t)TRI_NB * TRI_LDE + (size_t)TRI_RHS * TRI_LDX) * sizeof(T);
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  kern<<<(unsigned)((nrhs + TRI_RHS - 1) / TRI_RHS), TRI_RHS * TRI_LANES, smem, s>>>(nb, nrhs, eff_lower, unit, cj, (const T *)f, fs_i, fs_k, (R)ar, (R)ai,
                                                                       (T *)b, rs, cs);
  return cudaGetLastError();
}

template <class T>
cudaError_t expand_t(int uplo, int herm, const void *a, int64_t lda, void *out, int64_t ldo, int64_t i0, int64_t nr, int64_t j0, int64_t nc, cudaStream_t s) {
  const int64_t tiles = ((nr + 31) / 32) * ((nc + 7) / 8);
  const int64_t cap = (int64_t)sm_count() * 16;
  expand_symmetric_kernel<T><<<(unsigned)(tiles < cap ? tiles : cap), dim3(32, 8), 0, s>>>(uplo, herm, (const T *)a, lda, (T *)out, ldo, i0, nr, j0, nc);
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

/* ---- GEMM3M helpers: HBM-bound element-wise passes around three real GEMMs -------------------------------------
 * split3: one read of the complex operand (16 / 8 bytes per element), three real planes written; combine3: three real
 * planes read, C read only when beta != 0, C written.  O(mk + kn + mn) bytes against O(mnk) flops. */
template <class R2, class R>
__global__ void __launch_bounds__(256) split3_kernel(int64_t rows, int64_t cols, const R2 *__restrict__ x, int64_t ldx, R sign,
                                                     R *__restrict__ xr, R *__restrict__ xi, R *__restrict__ xs, int64_t ldp) {
  const int64_t tiles_r = (rows + 255) / 256;
  for (int64_t t = blockIdx.x; t < tiles_r * cols; t += gridDim.x) {
    const int64_t r = (t % tiles_r) * 256 + threadIdx.x, c = t / tiles_r;
    if (r < rows) {
      const R2 v = x[r + c * ldx];
      const R im = sign * v.y;
      xr[r + c * ldp] = v.x;
      xi[r + c * ldp] = im;
      xs[r + c * ldp] = v.x + im;
    }
  }
}
template <class R2, class R>
__global__ void __launch_bounds__(256) combine3_kernel(int64_t m, int64_t n, const R *__restrict__ t1, const R *__restrict__ t2,
                                                       const R *__restrict__ t3, int64_t ldt, R ar, R ai, R br, R bi, int use_beta,
                                                       R2 *__restrict__ c, int64_t ldc) {
  const int64_t tiles_r = (m + 255) / 256;
  for (int64_t t = blockIdx.x; t < tiles_r * n; t += gridDim.x) {
    const int64_t r = (t % tiles_r) * 256 + threadIdx.x, j = t / tiles_r;
    if (r < m) {
      const R a1 = t1[r + j * ldt], a2 = t2[r + j * ldt], a3 = t3[r + j * ldt];
      const R pr = a1 - a2, pi = (a3 - a1) - a2;          /* Re = XrYr - XiYi,  Im = (Xr + Xi)(Yr + Yi) - XrYr - XiYi */
      R2 o;
      o.x = ar * pr - ai * pi;
      o.y = ar * pi + ai * pr;
      if (use_beta) {
        const R2 old = c[r + j * ldc];
        o.x += br * old.x - bi * old.y;
        o.y += br * old.y + bi * old.x;
      }
      c[r + j * ldc] = o;
    }
  }
}
template <class R2, class R>
cudaError_t split3_t(int64_t rows, int64_t cols, const void *x, int64_t ldx, int conj, void *re, void *im, void *sum, int64_t ldp,
                     cudaStream_t s) {
  int64_t blocks = ((rows + 255) / 256) * cols;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  split3_kernel<R2, R><<<(unsigned)blocks, 256, 0, s>>>(rows, cols, (const R2 *)x, ldx, conj ? (R)-1 : (R)1, (R *)re, (R *)im, (R *)sum, ldp);
  return cudaGetLastError();
}
template <class R2, class R>
cudaError_t combine3_t(int64_t m, int64_t n, const void *t1, const void *t2, const void *t3, int64_t ldt, double ar, double ai, double br,
                       double bi, void *c, int64_t ldc, cudaStream_t s) {
  int64_t blocks = ((m + 255) / 256) * n;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  combine3_kernel<R2, R><<<(unsigned)blocks, 256, 0, s>>>(m, n, (const R *)t1, (const R *)t2, (const R *)t3, ldt, (R)ar, (R)ai, (R)br, (R)bi,
                                                          !(br == 0.0 && bi == 0.0), (R2 *)c, ldc);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_split3(int dtype, int64_t rows, int64_t cols, const void *x, int64_t ldx, int conj, void *re, void *im, void *sum,
                          int64_t ldp, cudaStream_t stream) {
  if (rows <= 0 || cols <= 0) return cudaSuccess;
  cudaError_t e;
  if (dtype == B200_C) e = split3_t<float2, float>(rows, cols, x, ldx, conj, re, im, sum, ldp, stream);
  else if (dtype == B200_Z) e = split3_t<double2, double>(rows, cols, x, ldx, conj, re, im, sum, ldp, stream);
  else return cudaErrorNotSupported;
  if (e == cudaSuccess) count_launch("split3");
  return e;
}
cudaError_t launch_combine3(int dtype, int64_t m, int64_t n, const void *t1, const void *t2, const void *t3, int64_t ldt, double alpha_re,
                            double alpha_im, double beta_re, double beta_im, void *c, int64_t ldc, cudaStream_t stream) {
  if (m <= 0 || n <= 0) return cudaSuccess;
  cudaError_t e;
  if (dtype == B200_C) e = combine3_t<float2, float>(m, n, t1, t2, t3, ldt, alpha_re, alpha_im, beta_re, beta_im, c, ldc, stream);
  else if (dtype == B200_Z) e = combine3_t<double2, double>(m, n, t1, t2, t3, ldt, alpha_re, alpha_im, beta_re, beta_im, c, ldc, stream);
  else return cudaErrorNotSupported;
  if (e == cudaSuccess) count_launch("combine3");
  return e;
}

cudaError_t launch_expand_symmetric(int dtype, int uplo, int herm, int64_t n, const void *a, int64_t lda, void *out,
                                    int64_t ldo, cudaStream_t stream, int64_t i0, int64_t nr, int64_t j0, int64_t nc) {
  if (nr < 0) { i0 = 0; nr = n; }
  if (nc < 0) { j0 = 0; nc = n; }
  if (nr <= 0 || nc <= 0) return cudaSuccess;
  cudaError_t e;
  switch (dtype) {
    case B200_S: e = expand_t<float>(uplo, 0, a, lda, out, ldo, i0, nr, j0, nc, stream); break;
    case B200_D: e = expand_t<double>(uplo, 0, a, lda, out, ldo, i0, nr, j0, nc, stream); break;
    case B200_C: e = expand_t<float2>(uplo, herm, a, lda, out, ldo, i0, nr, j0, nc, stream); break;
    case B200_Z: e = expand_t<double2>(uplo, herm, a, lda, out, ldo, i0, nr, j0, nc, stream); break;
    default: return cudaErrorNotSupported;
  }
  if (e == cudaSuccess) count_launch("expand_symmetric");
  return e;
}

/* largest diagonal block launch_tri_block takes: 128 for the real types (register kernel), 64 for the complex ones */
int tri_block_max(int dtype) {
  static const bool old_kernel = getenv("B200_TRI_KERNEL") && !strcmp(getenv("B200_TRI_KERNEL"), "smem");
  return (!old_kernel && (dtype == B200_S || dtype == B200_D)) ? 128 : TRI_NB;
}

cudaError_t launch_tri_block(int dtype, int solve, int nb, int64_t nrhs, int eff_lower, int unit, int cj, const void *f, int64_t fs_i,
                             int64_t fs_k, double ar, double ai, void *b, int64_t rs, int64_t cs, cudaStream_t stream) {
  if (nb <= 0 || nrhs <= 0) return cudaSuccess;
  if (nb > tri_block_max(dtype)) return cudaErrorInvalidValue;
  cudaError_t e;
#define TRI_CASE(T, R) (solve ? tri_block_t<T, R, true>(nb, nrhs, eff_lower, unit, cj, f, fs_i, fs_k, ar, ai, b, rs, cs, stream) \
                              : tri_block_t<T, R, false>(nb, nrhs, eff_lower, unit, cj, f, fs_i, fs_k, ar, ai, b, rs, cs, stream))
  static const bool old_kernel = getenv("B200_TRI_KERNEL") && !strcmp(getenv("B200_TRI_KERNEL"), "smem");   /* round 1's kernel, for comparison */
  const bool reg_ok = !old_kernel && (rs == 1 || cs == 1);
  if (!reg_ok && nb > TRI_NB) return cudaErrorInvalidValue;
  switch (dtype) {
    case B200_S: e = reg_ok ? tri_block_reg<float>(solve, nb, nrhs, eff_lower, unit, f, fs_i, fs_k, ar, b, rs, cs, stream) : TRI_CASE(float, float); break;
    case B200_D: e = reg_ok ? tri_block_reg<double>(solve, nb, nrhs, eff_lower, unit, f, fs_i, fs_k, ar, b, rs, cs, stream) : TRI_CASE(double, double); break;
    case B200_C: e = reg_ok ? tri_block_reg_cplx<float2, float>(solve, nb, nrhs, eff_lower, unit, cj, f, fs_i, fs_k, ar, ai, b, rs, cs, stream) : TRI_CASE(float2, float); break;
    case B200_Z: e = reg_ok ? tri_block_reg_cplx<double2, double>(solve, nb, nrhs, eff_lower, unit, cj, f, fs_i, fs_k, ar, ai, b, rs, cs, stream) : TRI_CASE(double2, double); break;
    default: return cudaErrorNotSupported;
  }
#undef TRI_CASE
  if (e == cudaSuccess) count_launch(solve ? "tri_block_solve" : "tri_block_multiply");
  return e;
}

template <class T>
__global__ void real_diagonal_kernel(int64_t n, T *__restrict__ c, int64_t ldc) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    c[i + i * ldc] = real_only(c[i + i * ldc]);
}

/* HERK / HER2K: the diagonal of the result is real by definition (zherkf.f:284, 321) */
cudaError_t launch_real_diagonal(int dtype, int64_t n, void *c, int64_t ldc, cudaStream_t stream) {
  if (n <= 0 || (dtype != B200_C && dtype != B200_Z)) return cudaSuccess;
  const unsigned blocks = (unsigned)((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024);
  if (dtype == B200_C) real_diagonal_kernel<float2><<<blocks, 256, 0, stream>>>(n, (float2 *)c, ldc);
  else real_diagonal_kernel<double2><<<blocks, 256, 0, stream>>>(n, (double2 *)c, ldc);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) count_launch("real_diagonal");
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
