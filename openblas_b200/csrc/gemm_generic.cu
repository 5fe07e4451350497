/*
 * gemm_generic.cu -- the latency / any-shape kernel: C <- alpha*op(A)*op(B) + beta*C for all
 * five precisions, all 16 op combinations, any m,n,k >= 0, any leading dimension and any
 * alignment.  It plays the role of the reference's small-matrix kernels
 * (interface/gemm.c:551-571, kernel/x86_64/dgemm_small_kernel_nn_skylakex.c) and is also the
 * fall-through for problems the roofline kernels cannot take.
 *
 * op() -- transpose and conjugation -- is applied when a tile is loaded into shared memory
 * (the reference does it by choosing an N or T packing routine, level3.c:62-78, and a
 * conj-variant micro-kernel, level3.c:80-93).  beta is fused into the epilogue; beta == 0
 * never reads C (kernel/generic/gemm_beta.c:52-71), alpha == 0 or k == 0 never reads A or B
 * (driver/level3/level3.c:252-259).
 */
#include "gemm_common.cuh"

namespace b200 {
namespace {

constexpr int TILE = 32;   /* C tile edge per CTA */
constexpr int KSTEP = 16;
constexpr int TPB = 256;   /* 16 x 16 threads, 2 x 2 outputs each */

__device__ __forceinline__ void mac(float &c, float a, float b) { c = fmaf(a, b, c); }
__device__ __forceinline__ void mac(double &c, double a, double b) { c = fma(a, b, c); }
__device__ __forceinline__ void mac(float2 &c, float2 a, float2 b) {
  c.x = fmaf(a.x, b.x, c.x); c.x = fmaf(-a.y, b.y, c.x);
  c.y = fmaf(a.x, b.y, c.y); c.y = fmaf(a.y, b.x, c.y);
}
__device__ __forceinline__ void mac(double2 &c, double2 a, double2 b) {
  c.x = fma(a.x, b.x, c.x); c.x = fma(-a.y, b.y, c.x);
  c.y = fma(a.x, b.y, c.y); c.y = fma(a.y, b.x, c.y);
}

/* element fetch with conjugation folded in; bf16 widens to fp32 exactly */
__device__ __forceinline__ float  fetch(const float *p, int64_t i, bool) { return p[i]; }
__device__ __forceinline__ double fetch(const double *p, int64_t i, bool) { return p[i]; }
__device__ __forceinline__ float  fetch(const uint16_t *p, int64_t i, bool) {
  return __uint_as_float(((uint32_t)p[i]) << 16);
}
__device__ __forceinline__ float2 fetch(const float2 *p, int64_t i, bool conj) {
  float2 v = p[i]; if (conj) v.y = -v.y; return v;
}
__device__ __forceinline__ double2 fetch(const double2 *p, int64_t i, bool conj) {
  double2 v = p[i]; if (conj) v.y = -v.y; return v;
}

template <class T> __device__ __forceinline__ T zero_of();
template <> __device__ __forceinline__ float zero_of<float>() { return 0.f; }
template <> __device__ __forceinline__ double zero_of<double>() { return 0.0; }
template <> __device__ __forceinline__ float2 zero_of<float2>() { return make_float2(0.f, 0.f); }
template <> __device__ __forceinline__ double2 zero_of<double2>() { return make_double2(0.0, 0.0); }

/* out = alpha*acc (+ beta*old when use_beta) */
__device__ __forceinline__ float epi(float acc, float old, float ar, float, float br, float, bool ub) {
  return ub ? fmaf(ar, acc, br * old) : ar * acc;
}
__device__ __forceinline__ double epi(double acc, double old, double ar, double, double br, double, bool ub) {
  return ub ? fma(ar, acc, br * old) : ar * acc;
}
__device__ __forceinline__ float2 epi(float2 acc, float2 old, float ar, float ai, float br, float bi, bool ub) {
  float2 r;
  r.x = ar * acc.x - ai * acc.y;
  r.y = ar * acc.y + ai * acc.x;
  if (ub) { r.x += br * old.x - bi * old.y; r.y += br * old.y + bi * old.x; }
  return r;
}
__device__ __forceinline__ double2 epi(double2 acc, double2 old, double ar, double ai, double br, double bi, bool ub) {
  double2 r;
  r.x = ar * acc.x - ai * acc.y;
  r.y = ar * acc.y + ai * acc.x;
  if (ub) { r.x += br * old.x - bi * old.y; r.y += br * old.y + bi * old.x; }
  return r;
}

/* one TILE x TILE tile of C at (m0, n0); every thread of the CTA takes part */
template <int DT>
__device__ __forceinline__ void generic_tile(const DeviceGemm &g, const int64_t m0, const int64_t n0,
                                             typename Traits<DT>::Out (*As)[TILE + 1], typename Traits<DT>::Out (*Bs)[TILE + 1]) {
  using In = typename Traits<DT>::In;
  using Out = typename Traits<DT>::Out;    /* also the accumulator type */
  using Real = typename Traits<DT>::Real;

  const In *A = (const In *)g.a;
  const In *B = (const In *)g.b;
  Out *C = (Out *)g.c;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const bool ta = g.transa & 1, ca = g.transa & 2;
  const bool tb = g.transb & 1, cb = g.transb & 2;

  Out acc[2][2];
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 2; j++) acc[i][j] = zero_of<Out>();

  const bool skip_product = (g.k == 0) || (g.alpha_re == 0.0 && g.alpha_im == 0.0);
  const int64_t kend = skip_product ? 0 : g.k;

  for (int64_t k0 = 0; k0 < kend; k0 += KSTEP) {
    /* 512 elements per operand tile, 2 per thread; thread -> element map follows the
     * contiguous direction of the stored matrix so global reads coalesce */
#pragma unroll
    for (int r = 0; r < 2; r++) {
      int i = threadIdx.x + r * TPB;
      int mm, kk;
      if (!ta) { mm = i % TILE; kk = i / TILE; } else { kk = i % KSTEP; mm = i / KSTEP; }
      int64_t gm = m0 + mm, gk = k0 + kk;
      Out v = zero_of<Out>();
      if (gm < g.m && gk < g.k) v = fetch(A, ta ? gk + gm * g.lda : gm + gk * g.lda, ca);
      As[kk][mm] = v;
      int nn;
      if (!tb) { kk = i % KSTEP; nn = i / KSTEP; } else { nn = i % TILE; kk = i / TILE; }
      int64_t gn = n0 + nn; gk = k0 + kk;
      v = zero_of<Out>();
      if (gn < g.n && gk < g.k) v = fetch(B, tb ? gn + gk * g.ldb : gk + gn * g.ldb, cb);
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < KSTEP; kk++) {
      Out a0 = As[kk][tx], a1 = As[kk][tx + 16];
      Out b0 = Bs[kk][ty], b1 = Bs[kk][ty + 16];
      mac(acc[0][0], a0, b0); mac(acc[1][0], a1, b0);
      mac(acc[0][1], a0, b1); mac(acc[1][1], a1, b1);
    }
    __syncthreads();
  }

  const Real ar = (Real)g.alpha_re, ai = (Real)g.alpha_im;
  const Real br = (Real)g.beta_re, bi = (Real)g.beta_im;
  const bool use_beta = !(g.beta_re == 0.0 && g.beta_im == 0.0);
#pragma unroll
  for (int j = 0; j < 2; j++) {
    int64_t gn = n0 + ty + 16 * j;
    if (gn >= g.n) continue;
#pragma unroll
    for (int i = 0; i < 2; i++) {
      int64_t gm = m0 + tx + 16 * i;
      if (gm >= g.m) continue;
      Out *dst = C + gm + gn * g.ldc;
      Out old = zero_of<Out>();
      if (use_beta) old = *dst;
      /* alpha == 0 / k == 0: the product term is dropped entirely, so a NaN alpha with
       * k == 0 cannot leak (level3.c:252-259 returns after the beta pass) */
      if (skip_product) {
        Out z = zero_of<Out>();
        *dst = epi(z, old, (Real)0, (Real)0, br, bi, use_beta);
      } else {
        *dst = epi(acc[i][j], old, ar, ai, br, bi, use_beta);
      }
    }
  }
}


/* Tiles are numbered along m first; a 1-D grid-stride walk, so any m and n launch (a 2-D grid would cap
 * n at 65535 tiles: m = 8, n = 3 000 000 is a legal call) */
template <int DT>
__global__ void __launch_bounds__(TPB)
gemm_generic_kernel(DeviceGemm g) {
  using Out = typename Traits<DT>::Out;
  __shared__ Out As[KSTEP][TILE + 1];
  __shared__ Out Bs[KSTEP][TILE + 1];
  const int64_t tiles_m = (g.m + TILE - 1) / TILE, tiles = tiles_m * ((g.n + TILE - 1) / TILE);
  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    generic_tile<DT>(g, (t % tiles_m) * TILE, (t / tiles_m) * TILE, As, Bs);
    __syncthreads();
  }
}

/* gemm_batch: ONE launch walks every tile of every matrix of the batch (interface/gemm_batch.c:322-366
 * queues one job per matrix for the thread pool; here the job list lives in device memory).
 * first_tile[i] = number of tiles of problems 0..i-1; a CTA finds its problem by bisection. */
template <int DT>
__global__ void __launch_bounds__(TPB)
gemm_grouped_kernel(const DeviceGemm *__restrict__ problems, const int64_t *__restrict__ first_tile, int count) {
  using Out = typename Traits<DT>::Out;
  __shared__ Out As[KSTEP][TILE + 1];
  __shared__ Out Bs[KSTEP][TILE + 1];
  __shared__ __align__(16) unsigned char g_bytes[sizeof(DeviceGemm)];     /* raw bytes: DeviceGemm has a default member initialiser */
  DeviceGemm &g = *reinterpret_cast<DeviceGemm *>(g_bytes);
  const int64_t tiles = first_tile[count];
  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    int lo = 0, hi = count - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (first_tile[mid] <= t) lo = mid; else hi = mid - 1;
    }
    if (threadIdx.x == 0) g = problems[lo];
    __syncthreads();
    const int64_t local = t - first_tile[lo], tiles_m = (g.m + TILE - 1) / TILE;
    generic_tile<DT>(g, (local % tiles_m) * TILE, (local / tiles_m) * TILE, As, Bs);
    __syncthreads();
  }
}

/* bf16 <-> fp32/fp64 conversion; rounding rule of kernel/x86_64/tobf16.c:46-96:
 * round-to-nearest-even, input denormals -> signed zero, NaN -> quiet NaN. */
__device__ __forceinline__ uint16_t f32_to_bf16(float f) {
  uint32_t u = __float_as_uint(f);
  switch (u & 0xff800000u) {
    case 0x00000000u: return 0x0000u;
    case 0x80000000u: return 0x8000u;
    case 0x7f800000u:
    case 0xff800000u: {
      uint16_t h = (uint16_t)(u >> 16);
      if (u & 0x007fffffu) h |= 0x0040u;
      return h;
    }
    default:
      u += ((u >> 16) & 1u) + 0x7fffu;
      return (uint16_t)(u >> 16);
  }
}
/* kernel/x86_64/bf16to.c:43-90: denormal inputs -> signed zero, NaN -> quiet NaN */
__device__ __forceinline__ float bf16_to_f32(uint16_t h) {
  switch (h & 0xff80u) {
    case 0x0000u: return __uint_as_float(0x00000000u);
    case 0x8000u: return __uint_as_float(0x80000000u);
    case 0x7f80u:
    case 0xff80u: {
      uint32_t u = ((uint32_t)h) << 16;
      if (h & 0x007fu) u |= 0x00400000u;
      return __uint_as_float(u);
    }
    default: return __uint_as_float(((uint32_t)h) << 16);
  }
}

__global__ void convert_kernel(int dir, int64_t n, const void *in, int64_t inc_in, void *out,
                               int64_t inc_out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    switch (dir) {
      case 0: ((uint16_t *)out)[i * inc_out] = f32_to_bf16(((const float *)in)[i * inc_in]); break;
      case 1: ((uint16_t *)out)[i * inc_out] = f32_to_bf16((float)((const double *)in)[i * inc_in]); break;
      case 2: ((float *)out)[i * inc_out] = bf16_to_f32(((const uint16_t *)in)[i * inc_in]); break;
      default: ((double *)out)[i * inc_out] = (double)bf16_to_f32(((const uint16_t *)in)[i * inc_in]); break;
    }
  }
}

}  // namespace

static unsigned grid_for(int64_t tiles) { return (unsigned)(tiles < 0x7fffffffll ? (tiles > 0 ? tiles : 1) : 0x7fffffffll); }

cudaError_t launch_generic(const DeviceGemm &g, cudaStream_t stream) {
  const unsigned grid = grid_for(((g.m + TILE - 1) / TILE) * ((g.n + TILE - 1) / TILE));
  switch (g.dtype) {
    case B200_S:  gemm_generic_kernel<B200_S><<<grid, TPB, 0, stream>>>(g); break;
    case B200_D:  gemm_generic_kernel<B200_D><<<grid, TPB, 0, stream>>>(g); break;
    case B200_C:  gemm_generic_kernel<B200_C><<<grid, TPB, 0, stream>>>(g); break;
    case B200_Z:  gemm_generic_kernel<B200_Z><<<grid, TPB, 0, stream>>>(g); break;
    case B200_SB: gemm_generic_kernel<B200_SB><<<grid, TPB, 0, stream>>>(g); break;
    default: return cudaErrorInvalidValue;
  }
  count_launch("gemm_generic");
  return cudaGetLastError();
}

int64_t generic_tile_count(int64_t m, int64_t n) { return ((m + TILE - 1) / TILE) * ((n + TILE - 1) / TILE); }

cudaError_t launch_grouped(int dtype, const DeviceGemm *problems_dev, const int64_t *first_tile_dev, int count, int64_t tiles,
                           cudaStream_t stream) {
  if (count <= 0 || tiles <= 0) return cudaSuccess;
  const unsigned grid = grid_for(tiles);
  switch (dtype) {
    case B200_S:  gemm_grouped_kernel<B200_S><<<grid, TPB, 0, stream>>>(problems_dev, first_tile_dev, count); break;
    case B200_D:  gemm_grouped_kernel<B200_D><<<grid, TPB, 0, stream>>>(problems_dev, first_tile_dev, count); break;
    case B200_C:  gemm_grouped_kernel<B200_C><<<grid, TPB, 0, stream>>>(problems_dev, first_tile_dev, count); break;
    case B200_Z:  gemm_grouped_kernel<B200_Z><<<grid, TPB, 0, stream>>>(problems_dev, first_tile_dev, count); break;
    case B200_SB: gemm_grouped_kernel<B200_SB><<<grid, TPB, 0, stream>>>(problems_dev, first_tile_dev, count); break;
    default: return cudaErrorInvalidValue;
  }
  count_launch("gemm_grouped");
  return cudaGetLastError();
}

cudaError_t launch_convert(int dir, int64_t n, const void *in, int64_t inc_in, void *out,
                           int64_t inc_out, cudaStream_t stream) {
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  convert_kernel<<<(unsigned)blocks, 256, 0, stream>>>(dir, n, in, inc_in, out, inc_out);
  count_launch("bf16_convert");
  return cudaGetLastError();
}

}  // namespace b200
