/*
 * cgemm_ffma.cu -- CGEMM on the FP32 FMA pipe (no TF32 / tensor-core substitution), complex
 * numbers kept interleaved.  A complex multiply-accumulate is two packed FFMA2 (fma.rn.f32x2
 * with a broadcast scalar) into TWO accumulators, with no swap / negate in the loop:
 *
 *     P += {ar, ai} * br   ->  P = sum {ar*br, ai*br}
 *     Q += {ar, ai} * bi   ->  Q = sum {ar*bi, ai*bi}
 *
 * and only the epilogue combines them, which is also where conjugation (op codes 2, 3) is
 * applied as two signs sA, sB:   re = P.x - sA*sB*Q.y ,  im = sA*P.y + sB*Q.x .
 * (The reference builds four micro-kernel variants for this: cgemm_kernel_8x2_skylakex.c
 * with -DNN/-DCN/-DNC/-DCC, kernel/Makefile.L3:877-916.)  The 16 op combinations share four
 * kernels (stored orientation of A x stored orientation of B).
 *
 * 64x64 complex C tile per CTA, k step 16, 256 threads with 4x4 complex outputs each, 3-stage
 * ring; same data movement scheme as sgemm_ffma.cu: S[k][mn] in shared memory, cp.async for
 * operands stored mn-contiguous, register-staged transposing stores for k-contiguous ones.
 */
#include "gemm_common.cuh"
#include "async_copy.cuh"
#include "sgemm_ws.cuh"
#include <cstdlib>

namespace b200 {
namespace {

constexpr int BM = 64, BN = 64, BK = 16;
constexpr int STAGES = 3;
constexpr int THREADS = 256;
constexpr int LDS = BM + 2;                          /* complex units; 528 B rows */
constexpr int OPERAND_ELEMS = BK * LDS;
constexpr int STAGE_ELEMS = 2 * OPERAND_ELEMS;
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_ELEMS * sizeof(float2);   /* 50688 B */

typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ void ffma2(u64 &c, u64 a, u64 b) {
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b));
}

/* Element-wise fallback for an mn-contiguous operand that is not 16-byte aligned. */
__device__ __noinline__ void load_mn_unaligned(float2 *s, const float2 *__restrict__ g, int64_t ld, int64_t mn0,
                                               int64_t k0, int64_t mn_end, int64_t k_end, int tid) {
  for (int i = 0; i < BM * BK / THREADS; i++) {
    int idx = tid + i * THREADS;
    int k = idx / BM, mn = idx % BM;
    int64_t gk = k0 + k, gmn = mn0 + mn;
    int bytes = (gk < k_end && gmn < mn_end) ? 8 : 0;
    const float2 *src = bytes ? g + gmn + gk * ld : g;
    cp_async8(s + k * LDS + mn, src, bytes);
  }
}

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(THREADS, 2)
cgemm_ffma_kernel(DeviceGemm g, int vec_a, int vec_b, int vec_c) {
  extern __shared__ __align__(16) float2 csmem[];
  const float2 *__restrict__ A = (const float2 *)g.a;
  const float2 *__restrict__ B = (const float2 *)g.b;
  float2 *__restrict__ C = (float2 *)g.c;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int pm, pn;
  warp_tile_position(lane, pm, pn);
  const int tm = (warp & 1) * 8 + pm;               /* 0..15: rows tm*2+{0,1}, 32+tm*2+{0,1} */
  const int tn = (warp >> 1) * 4 + pn;              /* 0..15 */
  const float sgn_a = (g.transa & 2) ? -1.f : 1.f;    /* conj(A) */
  const float sgn_b = (g.transb & 2) ? -1.f : 1.f;    /* conj(B) */

  const int64_t tiles_m = (g.m + BM - 1) / BM, tiles_n = (g.n + BN - 1) / BN;
  const int64_t tiles = g.tri ? tri_tile_count(tiles_m) : tiles_m * tiles_n;      /* tri: m == n, square tiles */
  const int64_t ktiles = (g.k + BK - 1) / BK;
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(csmem);

  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    int64_t bm, bn;
    if (g.tri) tri_tile_coords(t, g.tri, bm, bn); else banded_tile_coords<32>(t, tiles_m, tiles_n, bm, bn);
    const int64_t m0 = bm * BM, n0 = bn * BN;
    if (tri_outside(g.tri, m0, BM, n0, BN)) continue;      /* uniform over the CTA: no barrier is skipped by part of it */
    const bool masked = tri_partial(g.tri, m0, BM, n0, BN);

    u64 accp[4][4], accq[4][4];   /* [row i][col j]: P = sum a*br, Q = sum a*bi (see header) */
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) accp[i][j] = accq[i][j] = 0ull;

    TileLoader<true, 8, BM, BK, LDS, THREADS> la_mn, lb_mn;
    KStager<float2, BM, BK, LDS, THREADS> la_k, lb_k;
    if (A_MN) { if (vec_a) la_mn.init(A, g.lda, m0, g.m, tid); } else la_k.init(A, g.lda, m0, g.m, vec_a, tid);
    if (B_MN) { if (vec_b) lb_mn.init(B, g.ldb, n0, g.n, tid); } else lb_k.init(B, g.ldb, n0, g.n, vec_b, tid);
    float2 ra[4], rb[4];
    int req_slot = 0, dep_slot = 0, use_slot = 0;   /* ring positions: next request / deposit / consume */
    auto request = [&](int64_t kt_load) {
      const int stage = req_slot;
      req_slot = (req_slot + 1 == STAGES) ? 0 : req_slot + 1;
      const int64_t k_left = g.k - kt_load * BK;
      float2 *sa = csmem + stage * STAGE_ELEMS, *sb = sa + OPERAND_ELEMS;
      const uint32_t ua = smem_base + (uint32_t)(stage * STAGE_ELEMS * 8), ub = ua + (uint32_t)(OPERAND_ELEMS * 8);
      if (A_MN) {
        if (vec_a) { if (k_left < BK) la_mn.issue_tail(ua, (int)k_left); else la_mn.issue(ua); la_mn.advance(); }
        else load_mn_unaligned(sa, A, g.lda, m0, kt_load * BK, g.m, g.k, tid);
      } else { la_k.fetch(ra, k_left < BK ? (int)k_left : BK); la_k.advance(); }
      if (B_MN) {
        if (vec_b) { if (k_left < BK) lb_mn.issue_tail(ub, (int)k_left); else lb_mn.issue(ub); lb_mn.advance(); }
        else load_mn_unaligned(sb, B, g.ldb, n0, kt_load * BK, g.n, g.k, tid);
      } else { lb_k.fetch(rb, k_left < BK ? (int)k_left : BK); lb_k.advance(); }
    };
    auto deposit = [&](int64_t kt_load) {
      float2 *sa = csmem + dep_slot * STAGE_ELEMS, *sb = sa + OPERAND_ELEMS;
      if (!A_MN) la_k.store(sa, ra);
      if (!B_MN) lb_k.store(sb, rb);
      dep_slot = (dep_slot + 1 == STAGES) ? 0 : dep_slot + 1;
    };
#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
      if (s < ktiles) { request(s); deposit(s); }
      cp_async_commit();
    }

    for (int64_t kt = 0; kt < ktiles; kt++) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      const int64_t nk = kt + STAGES - 1;
      const bool refill = nk < ktiles;
      if (refill) request(nk);
      cp_async_commit();

      const float2 *sa = csmem + use_slot * STAGE_ELEMS + tm * 2;
      const float2 *sb = csmem + use_slot * STAGE_ELEMS + OPERAND_ELEMS + tn * 2;
      use_slot = (use_slot + 1 == STAGES) ? 0 : use_slot + 1;
#pragma unroll
      for (int k = 0; k < BK; k++) {
        ulonglong2 a01 = *reinterpret_cast<const ulonglong2 *>(sa + k * LDS);
        ulonglong2 a23 = *reinterpret_cast<const ulonglong2 *>(sa + k * LDS + 32);
        float4 b01 = *reinterpret_cast<const float4 *>(sb + k * LDS);
        float4 b23 = *reinterpret_cast<const float4 *>(sb + k * LDS + 32);
        const u64 a2[4] = {a01.x, a01.y, a23.x, a23.y};
        const float br[4] = {b01.x, b01.z, b23.x, b23.z};
        const float bi[4] = {b01.y, b01.w, b23.y, b23.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
          u64 brr = pack2(br[j], br[j]), bii = pack2(bi[j], bi[j]);
#pragma unroll
          for (int i = 0; i < 4; i++) {
            ffma2(accp[i][j], a2[i], brr);
            ffma2(accq[i][j], a2[i], bii);
          }
        }
      }
      if (refill) deposit(nk);
    }
    cp_async_wait<0>();
    __syncthreads();

    const float alr = (float)g.alpha_re, ali = (float)g.alpha_im, ber = (float)g.beta_re, bei = (float)g.beta_im;
    const bool use_beta = !(g.beta_re == 0.0 && g.beta_im == 0.0);
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int64_t n = n0 + (j < 2 ? tn * 2 + j : 32 + tn * 2 + (j - 2));
      if (n >= g.n) continue;
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int64_t m = m0 + h * 32 + tm * 2;
        if (m >= g.m) continue;
        float2 out[2];
#pragma unroll
        for (int e = 0; e < 2; e++) {
          float px, py, qx, qy;
          unpack2(accp[2 * h + e][j], px, py);
          unpack2(accq[2 * h + e][j], qx, qy);
          const float xr = px - sgn_a * sgn_b * qy, xi = sgn_a * py + sgn_b * qx;
          out[e].x = alr * xr - ali * xi;
          out[e].y = alr * xi + ali * xr;
        }
        float2 *p = C + m + n * g.ldc;
        if (vec_c && m + 1 < g.m && !masked) {
          if (use_beta) {
            float4 old = *reinterpret_cast<const float4 *>(p);
            out[0].x += ber * old.x - bei * old.y; out[0].y += ber * old.y + bei * old.x;
            out[1].x += ber * old.z - bei * old.w; out[1].y += ber * old.w + bei * old.z;
          }
          *reinterpret_cast<float4 *>(p) = make_float4(out[0].x, out[0].y, out[1].x, out[1].y);
        } else {
#pragma unroll
          for (int e = 0; e < 2; e++) {
            if (m + e >= g.m) break;
            if (!tri_keep(g.tri, m + e, n)) continue;
            if (use_beta) {
              float2 old = p[e];
              out[e].x += ber * old.x - bei * old.y; out[e].y += ber * old.y + bei * old.x;
            }
            p[e] = out[e];
          }
        }
      }
    }
  }
}

template <bool A_MN, bool B_MN>
cudaError_t launch_variant(const DeviceGemm &g, cudaStream_t stream, int vec_a, int vec_b, int vec_c) {
  static bool configured = false;
  auto kern = cgemm_ffma_kernel<A_MN, B_MN>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  int64_t tiles = ((g.m + BM - 1) / BM) * ((g.n + BN - 1) / BN);
  if (g.tri) tiles = tri_tile_count((g.m + BM - 1) / BM);
  int64_t cap = (int64_t)sm_count() * 2;
  int grid = (int)(tiles < cap ? tiles : cap);
  kern<<<grid, THREADS, SMEM_BYTES, stream>>>(g, vec_a, vec_b, vec_c);
  return cudaGetLastError();
}

typedef sws::Cfg<16, 2, 4, 5, 1, 2, 4, 224, 56, true, false, true> CwsConfig;

}  // namespace

cudaError_t launch_cgemm_ffma(const DeviceGemm &g, cudaStream_t stream) {
  if (g.dtype != B200_C) return cudaErrorNotSupported;
  if (((uintptr_t)g.a | (uintptr_t)g.b | (uintptr_t)g.c) & 7) return cudaErrorNotSupported;
  if (g.tri && g.m != g.n) return cudaErrorNotSupported;
  const bool a_mn = !(g.transa & 1), b_mn = (g.transb & 1);
  const int vec_a = (((uintptr_t)g.a & 15) == 0) && (g.lda % 2 == 0);
  const int vec_b = (((uintptr_t)g.b & 15) == 0) && (g.ldb % 2 == 0);
  const int vec_c = (((uintptr_t)g.c & 15) == 0) && (g.ldc % 2 == 0);
  cudaError_t e;
  /* Full grids go to the warp-specialised TMA kernel shared with SGEMM (sgemm_ws.cuh, CPLX: 128 x 64 complex
   * tiles, 8 x 4 complex per thread): same wave-count rule as launch_sgemm_ffma.  B200_CGEMM_TILE=64|128 forces
   * the 64 x 64 kernel / the TMA kernel. */
  {
    const char *ev = getenv("B200_CGEMM_TILE");
    const int forced = ev ? atoi(ev) : 0;
    const int64_t sms = sm_count();
    int64_t t64 = ((g.m + 63) / 64) * ((g.n + 63) / 64), tws = ((g.m + 127) / 128) * ((g.n + 63) / 64);
    if (g.tri) { t64 = (t64 + 1) / 2; tws = (tws + 1) / 2; }
    /* per-SM time in units of one 64 x 64 tile: the 64 x 64 kernel keeps two CTAs per SM, so its tiles go out in waves
     * of 2 * sms that take two units each (2048^3: 1024 tiles = 4 waves of 296 -> 8 units, measured 46 TFLOP/s against
     * 55 on full waves); a 128 x 64 tile of the TMA kernel is two units at 0.9 of the cost */
    const double est64 = 2.0 * (double)((t64 + 2 * sms - 1) / (2 * sms)), estws = 2.0 * 0.9 * (double)((tws + sms - 1) / sms);
    if (sws::eligible(g) && (forced == 128 || (forced == 0 && estws <= est64))) {
      e = sws::launch<CwsConfig, true>(g, stream);
      if (e == cudaSuccess) { count_launch("cgemm_ffma2_ws_tma_128x64x16"); return e; }
      if (e != cudaErrorNotSupported) return e;
    }
  }
  if (a_mn && b_mn) e = launch_variant<true, true>(g, stream, vec_a, vec_b, vec_c);
  else if (a_mn && !b_mn) e = launch_variant<true, false>(g, stream, vec_a, vec_b, vec_c);
  else if (!a_mn && b_mn) e = launch_variant<false, true>(g, stream, vec_a, vec_b, vec_c);
  else e = launch_variant<false, false>(g, stream, vec_a, vec_b, vec_c);
  if (e == cudaSuccess) count_launch("cgemm_ffma2_64x64x16");
  return e;
}

}  // namespace b200
