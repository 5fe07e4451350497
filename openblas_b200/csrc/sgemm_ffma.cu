/*
 * sgemm_ffma.cu -- SGEMM on the FP32 FMA pipe, register-blocked, NO TF32 / tensor-core
 * substitution (the north star forbids it): every product is an IEEE fp32 fused multiply-add,
 * issued as the Blackwell packed form FFMA2 (PTX fma.rn.f32x2: two independent fp32 FMAs on a
 * 64-bit register pair), which halves the issue slots the FMA stream needs and leaves room
 * for the shared-memory loads.
 *
 * Replaces sgemm_kernel_16x4_skylakex_3.c (AVX-512 register tile), the sgemm_{n,t}copy packing
 * (level3.c:62-78) and sgemm_beta (fused epilogue).  128x128 C tile per CTA, k step 16,
 * 256 threads, 8x8 outputs per thread held as 4x8 row-pairs; 3-stage ring.
 *
 * Both operands are kept in shared memory as S[k][mn] (row stride 132 floats) so that a thread
 * fetches its 8 rows / 8 columns for one k with two 16-byte loads each.  How a tile gets there
 * depends on the STORED orientation -- this is where op() is absorbed:
 *   mn-contiguous in memory (A not transposed, B transposed): cp.async straight into S[k][mn]
 *   k-contiguous in memory  (A transposed, B not transposed): 16-byte global loads into
 *       registers issued before the FMA block, scattered into S[k][mn] after it
 *       (the transpose happens in the register -> shared store, never as a separate pass)
 */
#include "gemm_common.cuh"
#include "async_copy.cuh"
#include "sgemm_ws.cuh"
#include <cstdlib>

namespace b200 {
namespace {

#ifndef SGEMM_BK
#define SGEMM_BK 16
#endif
constexpr int BK = SGEMM_BK;
constexpr int STAGES = 3;
/* Square CTA tile TILE x TILE, 8 x 8 outputs per thread.  TILE = 128: 256 threads, 2 CTAs per SM (the
 * roofline shape).  TILE = 64: 64 threads, up to 8 CTAs per SM -- same warps per SM, four times as
 * many tiles, for grids that would leave SMs idle at 128 x 128 (1024^3 is only 64 such tiles). */
template <int TILE>
struct SC {
  static constexpr int BM = TILE, BN = TILE;
  static constexpr int THREADS = TILE * TILE / 64;
  static constexpr int WARPS_M = TILE / 64;               /* a warp spans 8 (m) x 4 (n) thread positions */
  static constexpr int MINB = TILE == 128 ? 2 : 8;
  static constexpr int LDS = TILE + 4;                    /* 132 / 68 floats: 16-byte multiples */
  static constexpr int OPERAND_FLOATS = BK * LDS;
  static constexpr int STAGE_FLOATS = 2 * OPERAND_FLOATS;
  static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_FLOATS * sizeof(float);   /* 50688 / 26112 B */
};

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
template <int BM, int THREADS, int LDS>
__device__ __noinline__ void load_mn_unaligned(float *s, const float *__restrict__ g, int64_t ld, int64_t mn0,
                                               int64_t k0, int64_t mn_end, int64_t k_end, int tid) {
  for (int i = 0; i < BM * BK / THREADS; i++) {
    int idx = tid + i * THREADS;
    int k = idx / BM, mn = idx % BM;
    int64_t gk = k0 + k, gmn = mn0 + mn;
    int bytes = (gk < k_end && gmn < mn_end) ? 4 : 0;
    const float *src = bytes ? g + gmn + gk * ld : g;
    cp_async4(s + k * LDS + mn, src, bytes);
  }
}

/* PACKED: accumulate with FFMA2 (fma.rn.f32x2) on row pairs; otherwise with scalar FFMA. */
template <int TILE, bool A_MN, bool B_MN, int PACKED>
__global__ void __launch_bounds__(SC<TILE>::THREADS, SC<TILE>::MINB)
sgemm_ffma_kernel(DeviceGemm g, int vec_a, int vec_b, int vec_c, int probe) {
  constexpr int BM = SC<TILE>::BM, BN = SC<TILE>::BN, THREADS = SC<TILE>::THREADS, LDS = SC<TILE>::LDS;
  constexpr int OPERAND_FLOATS = SC<TILE>::OPERAND_FLOATS, STAGE_FLOATS = SC<TILE>::STAGE_FLOATS;
  constexpr int HM = BM / 2, HN = BN / 2;                 /* a thread owns rows tm*4..+3 and HM + tm*4..+3; columns likewise */
  extern __shared__ __align__(16) float fsmem[];
  const float *__restrict__ A = (const float *)g.a;
  const float *__restrict__ B = (const float *)g.b;
  float *__restrict__ C = (float *)g.c;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  /* warp = 8 (m) x 4 (n) threads; warps WARPS_M (m) x rest (n): a (TILE/8) x (TILE/8) grid of threads */
  int pm, pn;
  warp_tile_position(lane, pm, pn);
  const int tm = (warp % SC<TILE>::WARPS_M) * 8 + pm;
  const int tn = (warp / SC<TILE>::WARPS_M) * 4 + pn;

  const int64_t tiles_m = (g.m + BM - 1) / BM, tiles_n = (g.n + BN - 1) / BN;
  const int64_t tiles = g.tri ? tri_tile_count(tiles_m) : tiles_m * tiles_n;      /* tri: m == n, square tiles */
  const int64_t ktiles = (g.k + BK - 1) / BK;
  const float alpha = (float)g.alpha_re, beta = (float)g.beta_re;
  const bool use_beta = beta != 0.f;
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(fsmem);
  /* (tried: opaque per-lane fragment addresses + inline ld.shared to stop the compiler re-deriving
   * them from %tid every k tile -- ptxas then rotates the accumulators through 56 MOVs per k tile
   * and the kernel drops from 56.5 to 52.3 TFLOP/s; left as plain C++ loads) */

  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    int64_t bm, bn;
    if (g.tri) tri_tile_coords(t, g.tri, bm, bn); else banded_tile_coords<16>(t, tiles_m, tiles_n, bm, bn);
    const int64_t m0 = bm * BM, n0 = bn * BN;
    if (tri_outside(g.tri, m0, BM, n0, BN)) continue;      /* uniform over the CTA */
    const bool masked = tri_partial(g.tri, m0, BM, n0, BN);

    /* acc[p][j]: rows (2p, 2p+1) of the thread's 8 rows, column j of its 8 columns */
    u64 acc[4][8];
    float accs[PACKED ? 1 : 8][PACKED ? 1 : 8];
#pragma unroll
    for (int p = 0; p < 4; p++)
#pragma unroll
      for (int j = 0; j < 8; j++) acc[p][j] = 0ull;
    if (!PACKED) {
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) accs[PACKED ? 0 : i][PACKED ? 0 : j] = 0.f;
    }

    /* loaders with hoisted addresses; k tiles are requested strictly in order */
    TileLoader<true, 4, BM, BK, LDS, THREADS> la_mn, lb_mn;
    KStager<float, BM, BK, LDS, THREADS> la_k, lb_k;
    if (A_MN) { if (vec_a) la_mn.init(A, g.lda, m0, g.m, tid); } else la_k.init(A, g.lda, m0, g.m, vec_a, tid);
    if (B_MN) { if (vec_b) lb_mn.init(B, g.ldb, n0, g.n, tid); } else lb_k.init(B, g.ldb, n0, g.n, vec_b, tid);
    float ra[BM * BK / THREADS], rb[BN * BK / THREADS];
    /* issue the asynchronous part of k tile `kt_load` (cp.async for mn-contiguous operands, global
     * loads into registers for k-contiguous ones) */
    int req_slot = 0, dep_slot = 0, use_slot = 0;   /* ring positions: next request / deposit / consume */
    auto request = [&](int64_t kt_load) {
      const int stage = req_slot;
      req_slot = (req_slot + 1 == STAGES) ? 0 : req_slot + 1;
      const int64_t k_left = g.k - kt_load * BK;
      float *sa = fsmem + stage * STAGE_FLOATS, *sb = sa + OPERAND_FLOATS;
      const uint32_t ua = smem_base + (uint32_t)(stage * STAGE_FLOATS * 4), ub = ua + (uint32_t)(OPERAND_FLOATS * 4);
      if (A_MN) {
        if (vec_a) { if (k_left < BK) la_mn.issue_tail(ua, (int)k_left); else la_mn.issue(ua); la_mn.advance(); }
        else load_mn_unaligned<BM, THREADS, LDS>(sa, A, g.lda, m0, kt_load * BK, g.m, g.k, tid);
      } else { la_k.fetch(ra, k_left < BK ? (int)k_left : BK); la_k.advance(); }
      if (B_MN) {
        if (vec_b) { if (k_left < BK) lb_mn.issue_tail(ub, (int)k_left); else lb_mn.issue(ub); lb_mn.advance(); }
        else load_mn_unaligned<BM, THREADS, LDS>(sb, B, g.ldb, n0, kt_load * BK, g.n, g.k, tid);
      } else { lb_k.fetch(rb, k_left < BK ? (int)k_left : BK); lb_k.advance(); }
    };
    auto deposit = [&](int64_t kt_load) {   /* register-staged operands: transpose into S[k][mn] */
      float *sa = fsmem + dep_slot * STAGE_FLOATS, *sb = sa + OPERAND_FLOATS;
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
      /* probe (B200_SGEMM_PROBE, timing experiments only, results are garbage): 1 = no operand
       * traffic after the first ring fill, 2 = additionally no CTA barrier */
      if (!(probe == 2 && kt >= STAGES)) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
      }
      const int64_t nk = kt + STAGES - 1;
      const bool refill = nk < ktiles && !(probe && kt >= STAGES);
      if (refill) request(nk);
      cp_async_commit();

      const float *sa = fsmem + use_slot * STAGE_FLOATS + tm * 4;
      const float *sb = fsmem + use_slot * STAGE_FLOATS + OPERAND_FLOATS + tn * 4;
      use_slot = (use_slot + 1 == STAGES) ? 0 : use_slot + 1;
#pragma unroll
      for (int k = 0; k < BK; k++) {
        /* rows tm*4..+3 and HM+tm*4..+3 as two 64-bit pairs each; columns likewise as floats */
        float4 b_lo = *reinterpret_cast<const float4 *>(sb + k * LDS);
        float4 b_hi = *reinterpret_cast<const float4 *>(sb + k * LDS + HN);
        float bv[8] = {b_lo.x, b_lo.y, b_lo.z, b_lo.w, b_hi.x, b_hi.y, b_hi.z, b_hi.w};
        if (PACKED) {
          ulonglong2 a_lo = *reinterpret_cast<const ulonglong2 *>(sa + k * LDS);
          ulonglong2 a_hi = *reinterpret_cast<const ulonglong2 *>(sa + k * LDS + HM);
          u64 ap[4] = {a_lo.x, a_lo.y, a_hi.x, a_hi.y};
          if (PACKED == 2) {        /* row pair outer: the 64-bit A operand is the one kept in the reuse cache */
#pragma unroll
            for (int p = 0; p < 4; p++)
#pragma unroll
              for (int j = 0; j < 8; j++) ffma2(acc[p][j], ap[p], pack2(bv[j], bv[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 8; j++) {
              u64 bb = pack2(bv[j], bv[j]);
#pragma unroll
              for (int p = 0; p < 4; p++) ffma2(acc[p][j], ap[p], bb);
            }
          }
        } else {
          float4 a_lo = *reinterpret_cast<const float4 *>(sa + k * LDS);
          float4 a_hi = *reinterpret_cast<const float4 *>(sa + k * LDS + HM);
          float av[8] = {a_lo.x, a_lo.y, a_lo.z, a_lo.w, a_hi.x, a_hi.y, a_hi.z, a_hi.w};
#pragma unroll
          for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 8; j++)
              accs[PACKED ? 0 : i][PACKED ? 0 : j] = fmaf(av[i], bv[j], accs[PACKED ? 0 : i][PACKED ? 0 : j]);
        }
      }
      if (refill) deposit(nk);
    }
    cp_async_wait<0>();
    __syncthreads();

    /* epilogue: column j -> n, row pair p -> m */
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int64_t n = n0 + (j < 4 ? tn * 4 + j : HN + tn * 4 + (j - 4));
      if (n >= g.n) continue;
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int64_t m = m0 + h * HM + tm * 4;
        if (m >= g.m) continue;
        float v[4];
        if (PACKED) {
          unpack2(acc[2 * h][j], v[0], v[1]);
          unpack2(acc[2 * h + 1][j], v[2], v[3]);
        } else {
#pragma unroll
          for (int e = 0; e < 4; e++) v[e] = accs[PACKED ? 0 : 4 * h + e][PACKED ? 0 : j];
        }
        float *p = C + m + n * g.ldc;
        if (vec_c && m + 3 < g.m && !masked) {
          float4 o = make_float4(alpha * v[0], alpha * v[1], alpha * v[2], alpha * v[3]);
          if (use_beta) {
            float4 old = *reinterpret_cast<const float4 *>(p);
            o.x = fmaf(beta, old.x, o.x); o.y = fmaf(beta, old.y, o.y);
            o.z = fmaf(beta, old.z, o.z); o.w = fmaf(beta, old.w, o.w);
          }
          *reinterpret_cast<float4 *>(p) = o;
        } else {
#pragma unroll
          for (int e = 0; e < 4; e++) {
            if (m + e >= g.m) break;
            if (!tri_keep(g.tri, m + e, n)) continue;
            float o = alpha * v[e];
            if (use_beta) o = fmaf(beta, p[e], o);
            p[e] = o;
          }
        }
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * Producer-warp variant (B200_SGEMM_CFG=1; NOT the default: measured 43-47 TFLOP/s against 57 for the
 * kernel above -- with 12 consumer warps the LDS.128 wavefronts, ~3.4 per load, become the limiter).
 * Kept as the starting point for a larger register tile.  ncu on the kernel above: the FMA pipe is the busiest unit but
 * only ~77 % active -- the rest goes to barrier phases, LDS waits and the loader's integer work,
 * which shares the FMA pipe (IMAD).  Here 12 consumer warps (3 x 4, warp tile 64 x 32, 8x8 outputs per
 * thread -> 192 x 128 x 16 CTA tile) execute nothing but LDS.128 + FFMA2, with the next k step's
 * fragments prefetched into registers; two producer warps (A tiles / B tiles) fill a 4-stage ring
 * with cp.async -- 16-byte copies for operands stored mn-contiguous, 4-byte TRANSPOSING copies
 * (global k-contiguous -> shared S[k][mn]) for operands stored k-contiguous, so op() is still
 * absorbed in the load path -- and signal full[stage] mbarriers (cp.async.mbarrier.arrive.noinc);
 * consumers release stages through empty[stage].  No CTA-wide barrier in the main loop, any
 * m, n, k, any 4-byte aligned operands (zero fill at the edges via the cp.async src-size). */
namespace pw {
constexpr int BM = 192, BN = 128, BK = 16, STAGES = 4;
constexpr int CONSUMER_WARPS = 12, THREADS = (CONSUMER_WARPS + 4) * 32;   /* 2 producers + 2 idle: 128 regs each */
constexpr int LDA = BM + 4, LDB = BN + 4;                                /* floats; 16-byte multiples */
constexpr int A_FLOATS = BK * LDA, B_FLOATS = BK * LDB, STAGE_FLOATS = A_FLOATS + B_FLOATS;
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_FLOATS * 4 + 64;
}  // namespace pw

/* one operand tile (ROWS mn x 16 k floats) -> S[k][mn]; every lane arrives once on `bar` */
template <bool MN_CONTIG, int ROWS, int LD>
__device__ __forceinline__ void sgemm_produce(uint32_t s_tile, const float *__restrict__ g, int64_t ld, int64_t mn0,
                                              int64_t k0, int64_t mn_end, int64_t k_end, bool vec, uint32_t bar, int lane) {
  constexpr int BK = pw::BK;
  const int64_t mn_left = mn_end - mn0, k_left = k_end - k0;
  if (MN_CONTIG && vec) {
    constexpr int CPR = ROWS / 4;                       /* 16-byte chunks per k row */
    for (int c = lane; c < CPR * BK; c += 32) {
      const int k = c / CPR, mn = (c % CPR) * 4;
      const int64_t left = (k < k_left) ? mn_left - mn : 0;
      const int bytes = left >= 4 ? 16 : (left > 0 ? (int)left * 4 : 0);
      const float *src = bytes ? g + mn0 + mn + (k0 + k) * ld : g;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s_tile + (uint32_t)((k * LD + mn) * 4)), "l"(src), "r"(bytes) : "memory");
    }
  } else if (MN_CONTIG) {
    for (int c = lane; c < ROWS * BK; c += 32) {
      const int k = c / ROWS, mn = c % ROWS;
      const int bytes = (k < k_left && mn < mn_left) ? 4 : 0;
      const float *src = bytes ? g + mn0 + mn + (k0 + k) * ld : g;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(s_tile + (uint32_t)((k * LD + mn) * 4)), "l"(src), "r"(bytes) : "memory");
    }
  } else {
    /* k-contiguous in memory: lanes walk k (64 contiguous bytes per mn row), 2 rows per warp copy;
     * the 4-byte copy lands transposed at S[k][mn] */
    const int k = lane % BK, r0 = lane / BK;
    const float *src = g + k0 + k + (mn0 + r0) * ld;
    uint32_t dst = s_tile + (uint32_t)((k * LD + r0) * 4);
    const int64_t src_step = 2 * ld;
    const bool k_ok = k < k_left;
#pragma unroll 8
    for (int i = 0; i < ROWS / 2; i++) {
      const int bytes = (k_ok && r0 + 2 * i < mn_left) ? 4 : 0;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(bytes ? src : g), "r"(bytes) : "memory");
      src += src_step;
      dst += 8;
    }
  }
  cp_async_mbar_arrive_noinc(bar);
}

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(pw::THREADS, 1)
sgemm_ffma_pw_kernel(DeviceGemm g, int vec_a, int vec_b, int vec_c) {
  constexpr int BM = pw::BM, BN = pw::BN, BK = pw::BK, STAGES = pw::STAGES, LDA = pw::LDA, LDB = pw::LDB;
  constexpr int A_FLOATS = pw::A_FLOATS, STAGE_FLOATS = pw::STAGE_FLOATS;
  extern __shared__ __align__(16) float fsmem[];
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(fsmem);
  const uint32_t bars = smem_base + (uint32_t)(STAGES * STAGE_FLOATS * 4);
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  const float *__restrict__ A = (const float *)g.a;
  const float *__restrict__ B = (const float *)g.b;
  float *__restrict__ C = (float *)g.c;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t tiles_m = (g.m + BM - 1) / BM, tiles_n = (g.n + BN - 1) / BN, tiles = tiles_m * tiles_n;
  const int64_t ktiles = (g.k + BK - 1) / BK;

  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(full_bar(s), 64); mbar_init(empty_bar(s), pw::CONSUMER_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp >= pw::CONSUMER_WARPS) {
    if (warp >= pw::CONSUMER_WARPS + 2) return;          /* padding warps: only there for the register budget */
    const bool feeds_a = warp == pw::CONSUMER_WARPS;
    int slot = 0; uint32_t phase = 0;
    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
      int64_t bm, bn;
      banded_tile_coords<12>(t, tiles_m, tiles_n, bm, bn);
      const int64_t m0 = bm * BM, n0 = bn * BN;
      for (int64_t kt = 0; kt < ktiles; kt++) {
        mbar_wait(empty_bar(slot), phase ^ 1);
        const uint32_t sa = smem_base + (uint32_t)(slot * STAGE_FLOATS * 4), sb = sa + (uint32_t)(A_FLOATS * 4);
        if (feeds_a) sgemm_produce<A_MN, BM, LDA>(sa, A, g.lda, m0, kt * BK, g.m, g.k, vec_a, full_bar(slot), lane);
        else         sgemm_produce<B_MN, BN, LDB>(sb, B, g.ldb, n0, kt * BK, g.n, g.k, vec_b, full_bar(slot), lane);
        __syncwarp();
        if (++slot == STAGES) { slot = 0; phase ^= 1; }
      }
    }
    return;
  }

  /* consumers: warp grid 3 (m) x 4 (n), lanes 8 (m) x 4 (n) */
  const int wm = (warp % 3) * 64, wn = (warp / 3) * 32;
  int lm, ln;
  warp_tile_position(lane, lm, ln);
  const int a_off = wm + lm * 4, b_off = wn + ln * 4;     /* rows a_off..+3 and a_off+32..+35; cols b_off..+3, b_off+16..+19 */
  const float alpha = (float)g.alpha_re, beta = (float)g.beta_re;
  const bool use_beta = beta != 0.f;

  int slot = 0; uint32_t phase = 0;
  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    int64_t bm, bn;
    banded_tile_coords<12>(t, tiles_m, tiles_n, bm, bn);
    const int64_t m0 = bm * BM, n0 = bn * BN;

    u64 acc[4][8];
#pragma unroll
    for (int p = 0; p < 4; p++)
#pragma unroll
      for (int j = 0; j < 8; j++) acc[p][j] = 0ull;

    for (int64_t kt = 0; kt < ktiles; kt++) {
      mbar_wait(full_bar(slot), phase);
      const float *sa = fsmem + slot * STAGE_FLOATS + a_off;
      const float *sb = fsmem + slot * STAGE_FLOATS + A_FLOATS + b_off;
      ulonglong2 a_lo = *reinterpret_cast<const ulonglong2 *>(sa);
      ulonglong2 a_hi = *reinterpret_cast<const ulonglong2 *>(sa + 32);
      float4 b_lo = *reinterpret_cast<const float4 *>(sb);
      float4 b_hi = *reinterpret_cast<const float4 *>(sb + 16);
#pragma unroll
      for (int k = 0; k < BK; k++) {
        const u64 ap[4] = {a_lo.x, a_lo.y, a_hi.x, a_hi.y};
        const float bv[8] = {b_lo.x, b_lo.y, b_lo.z, b_lo.w, b_hi.x, b_hi.y, b_hi.z, b_hi.w};
        if (k + 1 < BK) {                                  /* prefetch the next k step's fragments */
          a_lo = *reinterpret_cast<const ulonglong2 *>(sa + (k + 1) * LDA);
          a_hi = *reinterpret_cast<const ulonglong2 *>(sa + (k + 1) * LDA + 32);
          b_lo = *reinterpret_cast<const float4 *>(sb + (k + 1) * LDB);
          b_hi = *reinterpret_cast<const float4 *>(sb + (k + 1) * LDB + 16);
        }
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const u64 bb = pack2(bv[j], bv[j]);
#pragma unroll
          for (int p = 0; p < 4; p++) ffma2(acc[p][j], ap[p], bb);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_bar(slot));
      if (++slot == STAGES) { slot = 0; phase ^= 1; }
    }

    /* epilogue: column j -> n, row quad h -> m */
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int64_t n = n0 + b_off + (j < 4 ? j : 16 + (j - 4));
      if (n >= g.n) continue;
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int64_t m = m0 + a_off + h * 32;
        if (m >= g.m) continue;
        float v[4];
        unpack2(acc[2 * h][j], v[0], v[1]);
        unpack2(acc[2 * h + 1][j], v[2], v[3]);
        float *p = C + m + n * g.ldc;
        if (vec_c && m + 3 < g.m) {
          float4 o = make_float4(alpha * v[0], alpha * v[1], alpha * v[2], alpha * v[3]);
          if (use_beta) {
            float4 old = *reinterpret_cast<const float4 *>(p);
            o.x = fmaf(beta, old.x, o.x); o.y = fmaf(beta, old.y, o.y);
            o.z = fmaf(beta, old.z, o.z); o.w = fmaf(beta, old.w, o.w);
          }
          *reinterpret_cast<float4 *>(p) = o;
        } else {
#pragma unroll
          for (int e = 0; e < 4; e++) {
            if (m + e >= g.m) break;
            float o = alpha * v[e];
            if (use_beta) o = fmaf(beta, p[e], o);
            p[e] = o;
          }
        }
      }
    }
  }
}

template <bool A_MN, bool B_MN>
cudaError_t launch_pw_variant(const DeviceGemm &g, cudaStream_t stream, int vec_a, int vec_b, int vec_c) {
  static bool configured = false;
  auto kern = sgemm_ffma_pw_kernel<A_MN, B_MN>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pw::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  int64_t tiles = ((g.m + pw::BM - 1) / pw::BM) * ((g.n + pw::BN - 1) / pw::BN);
  int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  kern<<<grid, pw::THREADS, pw::SMEM_BYTES, stream>>>(g, vec_a, vec_b, vec_c);
  return cudaGetLastError();
}

template <int TILE, bool A_MN, bool B_MN, int PACKED>
cudaError_t launch_variant(const DeviceGemm &g, cudaStream_t stream, int vec_a, int vec_b, int vec_c) {
  static bool configured = false;
  using S = SC<TILE>;
  auto kern = sgemm_ffma_kernel<TILE, A_MN, B_MN, PACKED>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  int64_t tiles = ((g.m + S::BM - 1) / S::BM) * ((g.n + S::BN - 1) / S::BN);
  if (g.tri) tiles = tri_tile_count((g.m + S::BM - 1) / S::BM);
  int64_t cap = (int64_t)sm_count() * S::MINB;
  int grid = (int)(tiles < cap ? tiles : cap);
  const char *pv = getenv("B200_SGEMM_PROBE");
  kern<<<grid, S::THREADS, S::SMEM_BYTES, stream>>>(g, vec_a, vec_b, vec_c, pv ? atoi(pv) : 0);
  return cudaGetLastError();
}

template <int TILE, int PACKED>
cudaError_t launch_ops(const DeviceGemm &g, cudaStream_t stream, int vec_a, int vec_b, int vec_c) {
  const bool a_mn = !(g.transa & 1), b_mn = (g.transb & 1);
  if (a_mn && b_mn) return launch_variant<TILE, true, true, PACKED>(g, stream, vec_a, vec_b, vec_c);
  if (a_mn && !b_mn) return launch_variant<TILE, true, false, PACKED>(g, stream, vec_a, vec_b, vec_c);
  if (!a_mn && b_mn) return launch_variant<TILE, false, true, PACKED>(g, stream, vec_a, vec_b, vec_c);
  return launch_variant<TILE, false, false, PACKED>(g, stream, vec_a, vec_b, vec_c);
}

/* 8 consumer warps (2 x 4, warp tile 128 x 32, 16 x 8 per thread, 224 registers via setmaxnreg) + a producer
 * warpgroup (one feeder warp per operand plus one helper warp each for the transposition of k-contiguous tiles, 56
 * registers), 5-stage ring of 24 KB stages, fragments of the next stage prefetched under the last FMA block of the
 * current one. */
typedef sws::Cfg<16, 2, 4, 5, 1, 2, 4, 224, 56, true, false, true> WsConfig;

}  // namespace

cudaError_t launch_sgemm_ffma(const DeviceGemm &g, cudaStream_t stream) {
  if (g.dtype != B200_S) return cudaErrorNotSupported;
  if (((uintptr_t)g.a | (uintptr_t)g.b | (uintptr_t)g.c) & 3) return cudaErrorNotSupported;
  const bool a_mn = !(g.transa & 1), b_mn = (g.transb & 1);
  const int vec_a = (((uintptr_t)g.a & 15) == 0) && (g.lda % 4 == 0);
  const int vec_b = (((uintptr_t)g.b & 15) == 0) && (g.ldb % 4 == 0);
  const int vec_c = (((uintptr_t)g.c & 15) == 0) && (g.ldc % 4 == 0);
  static int packed = -1, cfg = -1;
  if (packed < 0) { const char *ev = getenv("B200_SGEMM_PACKED"); packed = ev ? atoi(ev) : 1; }
  if (cfg < 0) { const char *ev = getenv("B200_SGEMM_CFG"); cfg = ev ? atoi(ev) : 0; }
  cudaError_t e;
  if (g.tri && (g.m != g.n || cfg == 1 || packed != 1)) return cudaErrorNotSupported;   /* only the default kernels mask */
  if (cfg == 1) {
    if (a_mn && b_mn) e = launch_pw_variant<true, true>(g, stream, vec_a, vec_b, vec_c);
    else if (a_mn && !b_mn) e = launch_pw_variant<true, false>(g, stream, vec_a, vec_b, vec_c);
    else if (!a_mn && b_mn) e = launch_pw_variant<false, true>(g, stream, vec_a, vec_b, vec_c);
    else e = launch_pw_variant<false, false>(g, stream, vec_a, vec_b, vec_c);
    if (e == cudaSuccess) count_launch("sgemm_ffma2_pw_192x128x16");
    return e;
  }
  /* Tile choice: a 128x128 tile costs four 64x64 tiles and CTAs spread evenly over the SMs.  Few 128x128
   * tiles (or a mostly empty last wave) leave FMA pipes idle, so the 64x64 tiling can win despite twice
   * the shared-memory traffic per flop.  Measured at 1024^3 (64 against 128): NT 35.8/21.6, NN 25.5/21.0,
   * TT 26.4/20.3, TN 15.5/18.9 TFLOP/s; on full grids the small tile reaches 85-95 % of the large one
   * (NT best, register-staged k-contiguous operands worst), hence a penalty per op and never for TN.
   * B200_SGEMM_TILE=64|128 forces one. */
  const char *tv = getenv("B200_SGEMM_TILE");
  const int forced = tv ? atoi(tv) : 0;
  const int64_t sms = sm_count();
  int64_t t128 = ((g.m + 127) / 128) * ((g.n + 127) / 128), t64 = ((g.m + 63) / 64) * ((g.n + 63) / 64);
  int64_t t256 = ((g.m + 255) / 256) * ((g.n + 127) / 128);
  if (g.tri) { t128 = (t128 + 1) / 2; t64 = (t64 + 1) / 2; t256 = (t256 + 1) / 2; }
  const double penalty = (a_mn && b_mn) ? 1.15 : (!a_mn && !b_mn) ? 1e9 : 1.3;
  const double est128 = 4.0 * (double)((t128 + sms - 1) / sms), est64 = penalty * (double)((t64 + sms - 1) / sms);   /* per-SM work */
  /* The warp-specialised TMA kernel (sgemm_ws.cuh: 256 x 128 tiles, 16 x 8 per thread, one CTA per SM) runs
   * at 61-62 TFLOP/s on full grids against 53-55 for the 128 x 128 kernel (profiles/r02_sgemm_lab_and_ffma_probe.txt),
   * i.e. a 256 x 128 tile costs 8 x 0.87 units of 64 x 64 work; it takes the problem when that beats the
   * other tilings' wave counts and TMA can address the operands.  B200_SGEMM_TILE=256 forces it. */
  const double est256 = 8.0 * 0.87 * (double)((t256 + sms - 1) / sms);
  if (packed == 1 && sws::eligible(g) && (forced == 256 || (forced == 0 && est256 <= est128 && est256 <= est64))) {
    e = sws::launch<WsConfig>(g, stream);
    if (e == cudaSuccess) { count_launch("sgemm_ffma2_ws_tma_256x128x16"); return e; }
    if (e != cudaErrorNotSupported) return e;
  }
  const bool small_tile = packed == 1 && (forced == 64 || (forced != 128 && forced != 256 && est64 < est128));
  if (small_tile) {
    e = launch_ops<64, 1>(g, stream, vec_a, vec_b, vec_c);
    if (e == cudaSuccess) count_launch("sgemm_ffma2_64x64x16");
    return e;
  }
  if (packed == 2) e = launch_ops<128, 2>(g, stream, vec_a, vec_b, vec_c);
  else if (packed) e = launch_ops<128, 1>(g, stream, vec_a, vec_b, vec_c);
  else e = launch_ops<128, 0>(g, stream, vec_a, vec_b, vec_c);
  if (e == cudaSuccess) count_launch(packed ? "sgemm_ffma2_128x128x16" : "sgemm_ffma_128x128x16");
  return e;
}

}  // namespace b200
