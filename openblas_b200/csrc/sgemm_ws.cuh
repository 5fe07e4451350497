/*
 * sgemm_ws.cuh -- warp-specialised, TMA-fed SGEMM on the FP32 FMA pipe (FFMA2 = fma.rn.f32x2, every
 * product an IEEE fp32 FMA: no TF32, no tensor cores).
 *
 * Why a second kernel: ncu on the round-1 kernel (every thread loads, CTA-wide barrier per k tile)
 * showed the FMA pipe only 78 % active with the issue slots half empty: the loss is lock-step
 * behaviour (all warps of a CTA meet at the barrier, then all wait for their first LDS), the
 * loaders' integer work (IMAD shares the FMA pipe) and the registers the register-staged
 * transposition of k-contiguous operands took from the FMA block.  Here
 *
 *   consumer warps   execute nothing but LDS.128 + FFMA2 and one mbarrier wait / arrive per k tile;
 *                    they drift apart, so one warp's k-tile turnaround hides behind the others
 *   producer warp    one lane issues TMA tile copies (cp.async.bulk.tensor.2d); OOB zero fill does
 *                    all edge handling.  An operand stored mn-contiguous lands directly as S[k][mn].
 *                    An operand stored k-contiguous cannot be transposed by TMA (4-byte elements, 16-byte
 *                    inner box), so it lands as a 64B-swizzled [mn][k] staging tile and the producer warp
 *                    rewrites it as S[k][mn]: LDS.128 along k (4 wavefronts = the minimum for 512 B,
 *                    thanks to the swizzle) + 4 conflict-free STS.32.  op() is absorbed in the load
 *                    path; the consumers run the same code for all four op combinations.
 *
 * Replaces sgemm_kernel_16x4_skylakex_3.c, sgemm_{n,t}copy (level3.c:62-78) and sgemm_beta
 * (fused epilogue; beta == 0 never reads C).
 */
#pragma once
#include <cuda.h>
#include "gemm_common.cuh"
#include "async_copy.cuh"

namespace b200 {
namespace sws {

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
__device__ __forceinline__ void mbar_expect_tx_only(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
#ifdef B200_LAB_SPIN_LIMIT
/* lab builds: a barrier that never completes traps instead of hanging the box */
__device__ __forceinline__ void mbar_wait_ws(uint32_t bar, uint32_t parity) {
  for (unsigned spin = 0;; spin++) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
    if (spin > B200_LAB_SPIN_LIMIT) __trap();
  }
}
#else
__device__ __forceinline__ void mbar_wait_ws(uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); }
#endif

/* TM rows x 8 columns per thread (rows as TM/2 f32x2 pairs); a warp covers 8 x 4 thread positions in
 * the 2 x 2-quad order of warp_tile_position(), i.e. a (8 TM) x 32 warp tile; WARPS_M x WARPS_N
 * consumer warps + 1 producer warp per CTA; MINB CTAs per SM. */
/* PW = 1: one extra producer warp, every warp keeps the launch register count.  PW = 4: a whole producer
 * warpgroup (only its first warp works) so that setmaxnreg can move registers: producers shrink to RP,
 * the consumer warpgroups grow to RC (8 * CONSUMERS * RC + 4 * RP must not exceed the CTA's launch pool). */
/* HELP (needs PW = 4): the two spare warps of the producer warpgroup each take half of the rows of a staged
 * (k-contiguous) tile, so the register -> shared transposition of one stage is shared by two warps per operand (ncu:
 * with one warp the consumers spent 4 % of their time waiting for full[] in NN, nothing in NT) */
template <int TM_, int WARPS_M_, int WARPS_N_, int STAGES_, int MINB_, int STG_ = 2, int PW_ = 1, int RC_ = 0, int RP_ = 0, bool XPF_ = false, bool SNAKE_ = false, bool HELP_ = false>
struct Cfg {
  static constexpr bool XPF = XPF_, SNAKE = SNAKE_, HELP = HELP_ && PW_ == 4;
  static_assert(!XPF_ || (16 % 2 == 0), "XPF needs an even BK (fragment buffers alternate)");
  static constexpr int PW = PW_, RC = RC_, RP = RP_;
  static_assert(PW == 1 || (PW == 4 && (WARPS_M_ * WARPS_N_) % 4 == 0), "setmaxnreg works on whole warpgroups");
  static constexpr int TM = TM_, TN = 8, BK = 16;
  static constexpr int WARPS_M = WARPS_M_, WARPS_N = WARPS_N_;
  static constexpr int CONSUMERS = WARPS_M * WARPS_N;
  static constexpr int THREADS = (CONSUMERS + PW) * 32;
  static constexpr int WM = 8 * TM, WN = 32;
  static constexpr int BM = WARPS_M * WM, BN = WARPS_N * WN;
  static constexpr int STAGES = STAGES_, MINB = MINB_, STG = STG_;
  static constexpr int A_FLOATS = BM * BK, B_FLOATS = BN * BK, STAGE_FLOATS = A_FLOATS + B_FLOATS;
  static constexpr int BAR_BYTES = 8 * (2 * STAGES + 4 * STG);      /* full, empty, staging landed (2 operands), staging consumed (2 operands) */
  static constexpr int FEEDERS = PW_ == 4 ? 2 : 1;          /* producer warps that actually work */
  static_assert(TM % 4 == 0 && BM <= 256 && BN <= 256, "TMA boxes are at most 256 elements per dimension");
  static constexpr size_t smem_bytes(bool a_mn, bool b_mn) {
    return 1024 + sizeof(float) * ((size_t)STAGES * STAGE_FLOATS + (a_mn ? 0 : (size_t)STG * A_FLOATS) + (b_mn ? 0 : (size_t)STG * B_FLOATS)) + BAR_BYTES;
  }
};

/* The C tiles a CTA owns, in the order it visits them; producer and consumers walk the same list. */
template <int BM, int BN>
struct TileWalk {
  int64_t tiles_m, tiles_n, tiles, t;
  int tri;
  __device__ __forceinline__ void init(const DeviceGemm &g) {
    tiles_m = (g.m + BM - 1) / BM; tiles_n = (g.n + BN - 1) / BN;
    tri = g.tri;
    tiles = (tri && BM == BN) ? tri_tile_count(tiles_m) : (tri && BM == 2 * BN) ? tri21_tile_count(tri, tiles_m, tiles_n) : tiles_m * tiles_n;
    t = (int64_t)blockIdx.x - (int64_t)gridDim.x;
  }
  /* next tile of this CTA; false when there is none */
  __device__ __forceinline__ bool next(int64_t &m0, int64_t &n0) {
    for (;;) {
      t += gridDim.x;
      if (t >= tiles) return false;
      int64_t bm, bn;
      if (tri && BM == BN) tri_tile_coords(t, tri, bm, bn);
      else if (tri && BM == 2 * BN) tri21_tile_coords(t, tri, bm, bn);
      else banded_tile_coords<16>(t, tiles_m, tiles_n, bm, bn);
      m0 = bm * BM; n0 = bn * BN;
      if (!tri_outside(tri, m0, BM, n0, BN)) return true;
    }
  }
};

/* producer warp: staging tile [ROWS][16 floats], 64B-swizzled by TMA  ->  S[k][mn] (row stride ROWS floats) */
template <int ROWS, bool CPLX, int PART = 0>
__device__ __forceinline__ void transpose_staged(const float *stg, float *dst, int lane) {
  constexpr int RB = ROWS / 32, RB0 = PART == 2 ? RB / 2 : 0, RB1 = PART == 1 ? RB / 2 : RB;
  if (!CPLX) {
#pragma unroll
    for (int rb = RB0; rb < RB1; rb++) {
      const int row = rb * 32 + lane;
      const int sw = (row >> 1) & 3;                       /* Swizzle<2,4,3>: 16-byte chunk ^= address bits 7..8 */
      float4 v[4];
#pragma unroll
      for (int kq = 0; kq < 4; kq++) v[kq] = *reinterpret_cast<const float4 *>(stg + row * 16 + ((kq ^ sw) << 2));
#pragma unroll
      for (int kq = 0; kq < 4; kq++) {
        dst[(kq * 4 + 0) * ROWS + row] = v[kq].x;
        dst[(kq * 4 + 1) * ROWS + row] = v[kq].y;
        dst[(kq * 4 + 2) * ROWS + row] = v[kq].z;
        dst[(kq * 4 + 3) * ROWS + row] = v[kq].w;
      }
    }
  } else {
    /* complex: ROWS complex rows of 16 complex k (128 B, Swizzle<3,4,3>: chunk ^= row & 7) -> S[k][mn] of
     * interleaved (re, im): 8-byte elements move as STS.64 */
    float2 *d2 = reinterpret_cast<float2 *>(dst);
#pragma unroll
    for (int rb = RB0; rb < RB1; rb++) {
      const int row = rb * 32 + lane;
      const int sw = row & 7;
#pragma unroll
      for (int half = 0; half < 2; half++) {
        float4 v[4];
#pragma unroll
        for (int c = 0; c < 4; c++) v[c] = *reinterpret_cast<const float4 *>(stg + row * 32 + (((half * 4 + c) ^ sw) << 2));
#pragma unroll
        for (int c = 0; c < 4; c++) {
          d2[((half * 4 + c) * 2 + 0) * ROWS + row] = make_float2(v[c].x, v[c].y);
          d2[((half * 4 + c) * 2 + 1) * ROWS + row] = make_float2(v[c].z, v[c].w);
        }
      }
    }
  }
}

/* CPLX: CGEMM on the same loop.  The float view of an interleaved complex operand has twice the rows; a row
 * pair of the real kernel is then one complex element (re, im) of A and the 8 scalars of a thread are 4
 * complex elements (br, bi) of B, so FFMA2 builds P = a * br and Q = a * bi and the epilogue combines
 * re = P.re -/+ Q.im, im = P.im +/- Q.re (signs carry conj(A) / conj(B)): 4 real FMAs per complex MAC, no
 * swap or negate in the loop.  BM, BN stay FLOAT extents (256 x 128 floats = 128 x 64 complex). */
template <class C, bool A_MN, bool B_MN, bool CPLX>
__global__ void __launch_bounds__(C::THREADS, C::MINB)
sgemm_ws_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, DeviceGemm g, int vec_c) {
  constexpr int TM = C::TM, BK = C::BK, BM = C::BM, BN = C::BN, STAGES = C::STAGES, STG = C::STG;
  constexpr int EM = CPLX ? 2 : 1, TILE_M = BM / EM, TILE_N = BN / EM;     /* tile extents in matrix elements */
  constexpr int A_FLOATS = C::A_FLOATS, B_FLOATS = C::B_FLOATS, STAGE_FLOATS = C::STAGE_FLOATS;
  constexpr int NP = TM / 2, NG = TM / 4;                 /* row pairs, row groups of 4 per thread */
  extern __shared__ __align__(16) uint8_t raw_smem[];
  /* 1024-byte alignment by OFFSET from the shared array, so the compiler keeps the shared state space */
  float *ring = reinterpret_cast<float *>(raw_smem + ((1024u - ((uint32_t)__cvta_generic_to_shared(raw_smem) & 1023u)) & 1023u));
  float *stg_a = ring + (size_t)STAGES * STAGE_FLOATS;
  float *stg_b = stg_a + (A_MN ? 0 : STG * A_FLOATS);
  float *bar_mem = stg_b + (B_MN ? 0 : STG * B_FLOATS);
  const uint32_t ring_u = (uint32_t)__cvta_generic_to_shared(ring);
  const uint32_t stg_a_u = (uint32_t)__cvta_generic_to_shared(stg_a), stg_b_u = (uint32_t)__cvta_generic_to_shared(stg_b);
  const uint32_t bars = (uint32_t)__cvta_generic_to_shared(bar_mem);
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto stg_bar = [&](int w, int s) { return bars + 8u * (2 * STAGES + w * STG + s); };
  auto stg_free = [&](int w, int s) { return bars + 8u * (2 * STAGES + 2 * STG + w * STG + s); };     /* HELP: the helper is done with staging buffer s */
  constexpr int HELPERS = C::HELP ? ((A_MN ? 0 : 1) + (B_MN ? 0 : 1)) : 0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t ktiles = (g.k + BK - 1) / BK;

  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(full_bar(s), C::FEEDERS + HELPERS); mbar_init(empty_bar(s), C::CONSUMERS); }
    for (int s = 0; s < 4 * STG; s++) mbar_init(stg_bar(0, s), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  TileWalk<TILE_M, TILE_N> walk;
  walk.init(g);

  if (warp >= C::CONSUMERS) {
    /* ------------------------------------------------------------------ producer warps
     * PW == 1: one warp feeds both operands.  PW == 4: warp 0 of the producer warpgroup feeds A, warp 1
     * feeds B (two independent pipelines, each arrives once per stage on full[]), warps 2-3 only donate
     * registers. */
    if (C::RP) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C::RP ? C::RP : 24));
    const int pw = warp - C::CONSUMERS;
    if (C::HELP && pw >= 2) {
      /* helper of feeder pw - 2: the second half of the rows of every staged tile of that operand */
      const int pf = pw - 2;
      if ((pf == 0 && A_MN) || (pf == 1 && B_MN)) return;
      uint32_t it = 0;
      int64_t m0, n0;
      while (walk.next(m0, n0)) {
        for (int64_t kt = 0; kt < ktiles; kt++, it++) {
          const int slot = (int)(it % STAGES), j = (int)(it % STG);
          mbar_wait_ws(empty_bar(slot), ((it / STAGES) & 1) ^ 1);
          mbar_wait_ws(stg_bar(pf, j), (it / STG) & 1);
          if (pf == 0) transpose_staged<TILE_M, CPLX, 2>(stg_a + j * A_FLOATS, ring + slot * STAGE_FLOATS, lane);
          else transpose_staged<TILE_N, CPLX, 2>(stg_b + j * B_FLOATS, ring + slot * STAGE_FLOATS + A_FLOATS, lane);
          __syncwarp();
          if (lane == 0) { mbar_arrive(full_bar(slot)); mbar_arrive(stg_free(pf, j)); }
        }
      }
      return;
    }
    if (pw >= C::FEEDERS) return;
    const bool do_a = C::FEEDERS == 1 || pw == 0, do_b = C::FEEDERS == 1 || pw == 1;
    const bool stage_a = do_a && !A_MN, stage_b = do_b && !B_MN;
    const uint32_t direct_bytes = ((do_a && A_MN ? A_FLOATS : 0) + (do_b && B_MN ? B_FLOATS : 0)) * 4;
    const uint32_t staged_bytes = ((stage_a ? A_FLOATS : 0) + (stage_b ? B_FLOATS : 0)) * 4;
    /* the staging loads run LOOK k tiles ahead of the ring fill, on their own walk of the tile list */
    constexpr int LOOK = STG - 1;
    TileWalk<TILE_M, TILE_N> ahead = walk;
    int64_t am0 = 0, an0 = 0, akt = 0;
    bool ahead_ok = false;
    uint32_t stg_issue = 0;                               /* staging steps issued so far */
    auto stage_next = [&]() {                             /* issue the staging copies of the next k tile of `ahead` */
      if (!staged_bytes) return;
      if (!ahead_ok || akt >= ktiles) { ahead_ok = ahead.next(am0, an0); akt = 0; if (!ahead_ok) return; }
      if (C::HELP && stg_issue >= (uint32_t)STG) mbar_wait_ws(stg_free(pw, (int)(stg_issue % STG)), ((stg_issue / STG) - 1) & 1);   /* the helper has read the buffer's previous tile */
      if (lane == 0) {
        const int j = (int)(stg_issue % STG);
        mbar_expect_tx(stg_bar(pw, j), staged_bytes);
        if (stage_a) tma_load_2d(stg_a_u + (uint32_t)(j * A_FLOATS * 4), &map_a, stg_bar(pw, j), (int)(akt * BK * EM), (int)am0);
        if (stage_b) tma_load_2d(stg_b_u + (uint32_t)(j * B_FLOATS * 4), &map_b, stg_bar(pw, j), (int)(akt * BK * EM), (int)an0);
      }
      stg_issue++; akt++;
    };
    for (int i = 0; i < LOOK; i++) stage_next();

    uint32_t it = 0;
    int64_t m0, n0;
    while (walk.next(m0, n0)) {
      for (int64_t kt = 0; kt < ktiles; kt++, it++) {
        const int slot = (int)(it % STAGES);
        const uint32_t ph = (it / STAGES) & 1;
        stage_next();                                     /* staging buffer (it + LOOK) % STG: its previous tile was transposed by this warp already */
        mbar_wait_ws(empty_bar(slot), ph ^ 1);
        const uint32_t sa = ring_u + (uint32_t)(slot * STAGE_FLOATS * 4), sb = sa + (uint32_t)(A_FLOATS * 4);
        if (direct_bytes && lane == 0) {
          mbar_expect_tx_only(full_bar(slot), direct_bytes);
          if (do_a && A_MN) tma_load_2d(sa, &map_a, full_bar(slot), (int)(m0 * EM), (int)(kt * BK));
          if (do_b && B_MN) tma_load_2d(sb, &map_b, full_bar(slot), (int)(n0 * EM), (int)(kt * BK));
        }
        if (staged_bytes) {
          const int j = (int)(it % STG);
          mbar_wait_ws(stg_bar(pw, j), (it / STG) & 1);
          if (stage_a) transpose_staged<TILE_M, CPLX, C::HELP ? 1 : 0>(stg_a + j * A_FLOATS, ring + slot * STAGE_FLOATS, lane);
          if (stage_b) transpose_staged<TILE_N, CPLX, C::HELP ? 1 : 0>(stg_b + j * B_FLOATS, ring + slot * STAGE_FLOATS + A_FLOATS, lane);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(full_bar(slot));
      }
    }
    return;
  }

  /* -------------------------------------------------------------------- consumer warps */
  if (C::RC) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(C::RC ? C::RC : 24));
  int pm, pn;
  warp_tile_position(lane, pm, pn);
  const int wm0 = (warp % C::WARPS_M) * C::WM, wn0 = (warp / C::WARPS_M) * C::WN;
  const int a_off = wm0 + pm * 4, b_off = wn0 + pn * 4;   /* rows a_off + 32 g .. + 3; columns b_off + 16 h .. + 3 */
  const float alpha = (float)g.alpha_re, beta = (float)g.beta_re;
  const bool use_beta = beta != 0.f;
  float *__restrict__ Cp = (float *)g.c;

  uint32_t it = 0;
  int64_t m0, n0;
  while (walk.next(m0, n0)) {
    const bool masked = tri_partial(g.tri, m0, TILE_M, n0, TILE_N);
    u64 acc[NP][8];
#pragma unroll
    for (int p = 0; p < NP; p++)
#pragma unroll
      for (int j = 0; j < 8; j++) acc[p][j] = 0ull;

    ulonglong2 af[2][NG];
    float4 bf[2][2];
    auto load_frags = [&](int buf, const float *sa, const float *sb) {
#pragma unroll
      for (int gi = 0; gi < NG; gi++) af[buf][gi] = *reinterpret_cast<const ulonglong2 *>(sa + gi * 32);
      bf[buf][0] = *reinterpret_cast<const float4 *>(sb);
      bf[buf][1] = *reinterpret_cast<const float4 *>(sb + 16);
    };
    if (C::XPF) {                                          /* fragments of k = 0 of the tile's first stage */
      const int slot = (int)(it % STAGES);
      mbar_wait_ws(full_bar(slot), (it / STAGES) & 1);
      load_frags(0, ring + slot * STAGE_FLOATS + a_off, ring + slot * STAGE_FLOATS + A_FLOATS + b_off);
    }
    for (int64_t kt = 0; kt < ktiles; kt++, it++) {
      const int slot = (int)(it % STAGES);
      const float *sa = ring + slot * STAGE_FLOATS + a_off;
      const float *sb = ring + slot * STAGE_FLOATS + A_FLOATS + b_off;
      if (!C::XPF) {
        mbar_wait_ws(full_bar(slot), (it / STAGES) & 1);
        load_frags(0, sa, sb);
      }
#pragma unroll
      for (int k = 0; k < BK; k++) {
        const int cur = k & 1;
        if (k + 1 < BK) {
          load_frags(cur ^ 1, sa + (k + 1) * BM, sb + (k + 1) * BN);
        } else if (C::XPF && kt + 1 < ktiles) {
          /* XPF: the first fragments of the NEXT stage are fetched under the last FMA block of this one,
           * so a warp's k-tile turnaround (barrier poll + LDS latency) is off its critical path */
          const int nslot = (int)((it + 1) % STAGES);
          mbar_wait_ws(full_bar(nslot), ((it + 1) / STAGES) & 1);
          load_frags(cur ^ 1, ring + nslot * STAGE_FLOATS + a_off, ring + nslot * STAGE_FLOATS + A_FLOATS + b_off);
        }
        const float bv[8] = {bf[cur][0].x, bf[cur][0].y, bf[cur][0].z, bf[cur][0].w, bf[cur][1].x, bf[cur][1].y, bf[cur][1].z, bf[cur][1].w};
        u64 ap[NP];
#pragma unroll
        for (int gi = 0; gi < NG; gi++) { ap[2 * gi] = af[cur][gi].x; ap[2 * gi + 1] = af[cur][gi].y; }
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const u64 bb = pack2(bv[j], bv[j]);
          /* SNAKE: walk the row pairs back and forth so that consecutive FFMA2 share either the scalar
           * (inside a column) or the row pair (at the turn): one new register operand per instruction */
#pragma unroll
          for (int q = 0; q < NP; q++) {
            const int pp = (C::SNAKE && (j & 1)) ? NP - 1 - q : q;
            ffma2(acc[pp][j], ap[pp], bb);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_bar(slot));
    }

    if (!CPLX) {
      /* epilogue: column j -> n, row group gi -> 4 consecutive m */
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int64_t n = n0 + b_off + (j < 4 ? j : 16 + (j - 4));
        if (n >= g.n) continue;
#pragma unroll
        for (int gi = 0; gi < NG; gi++) {
          const int64_t m = m0 + a_off + gi * 32;
          if (m >= g.m) continue;
          float v[4];
          unpack2(acc[2 * gi][j], v[0], v[1]);
          unpack2(acc[2 * gi + 1][j], v[2], v[3]);
          float *p = Cp + m + n * g.ldc;
          if (vec_c && m + 3 < g.m && !masked) {
            float4 o = make_float4(alpha * v[0], alpha * v[1], alpha * v[2], alpha * v[3]);
            if (use_beta) {
              const float4 old = *reinterpret_cast<const float4 *>(p);
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
    } else {
      /* complex epilogue: scalars (2 cc, 2 cc + 1) = (br, bi) of complex column cc; row pair = one complex row */
      const float sa = (g.transa & 2) ? -1.f : 1.f, sb = (g.transb & 2) ? -1.f : 1.f, sab = sa * sb;
      const float ai = (float)g.alpha_im, bi = (float)g.beta_im;
      const bool use_cbeta = use_beta || bi != 0.f;
      float2 *__restrict__ Cc = (float2 *)g.c;
#pragma unroll
      for (int cc = 0; cc < 4; cc++) {
        const int64_t n = n0 + (b_off + (cc < 2 ? 0 : 16)) / 2 + (cc & 1);
        if (n >= g.n) continue;
#pragma unroll
        for (int gi = 0; gi < NG; gi++) {
          const int64_t m = m0 + (a_off + gi * 32) / 2;
          if (m >= g.m) continue;
          float2 out[2];
#pragma unroll
          for (int e = 0; e < 2; e++) {
            float pr, pi, qr, qi;
            unpack2(acc[2 * gi + e][2 * cc], pr, pi);         /* (sum ar br, sum ai br) */
            unpack2(acc[2 * gi + e][2 * cc + 1], qr, qi);     /* (sum ar bi, sum ai bi) */
            const float re = pr - sab * qi, im = sb * qr + sa * pi;
            out[e] = make_float2(alpha * re - ai * im, alpha * im + ai * re);
          }
          float2 *p = Cc + m + n * g.ldc;
          if (vec_c && m + 1 < g.m && !masked) {
            if (use_cbeta) {
              const float4 old = *reinterpret_cast<const float4 *>(p);
              out[0].x += beta * old.x - bi * old.y; out[0].y += beta * old.y + bi * old.x;
              out[1].x += beta * old.z - bi * old.w; out[1].y += beta * old.w + bi * old.z;
            }
            *reinterpret_cast<float4 *>(p) = make_float4(out[0].x, out[0].y, out[1].x, out[1].y);
          } else {
#pragma unroll
            for (int e = 0; e < 2; e++) {
              if (m + e >= g.m) break;
              if (!tri_keep(g.tri, m + e, n)) continue;
              float2 o = out[e];
              if (use_cbeta) { const float2 old = p[e]; o.x += beta * old.x - bi * old.y; o.y += beta * old.y + bi * old.x; }
              p[e] = o;
            }
          }
        }
      }
    }
  }
}

/* ------------------------------------------------------------------------------ host side */
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}
/* 2-D fp32 map: dim0 = the contiguous dimension (extent0 floats), dim1 strided by pitch_bytes */
inline bool make_map_f32(CUtensorMap *map, const void *ptr, uint64_t extent0, uint64_t extent1, uint64_t pitch_bytes, uint32_t box0,
                         uint32_t box1, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {extent0, extent1};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

/* TMA needs 16-byte aligned bases and pitches; tile coordinates are int32 (in floats) */
inline bool eligible(const DeviceGemm &g) {
  const bool cplx = g.dtype == B200_C;
  if ((g.dtype != B200_S && !cplx) || g.k <= 0) return false;
  const int es = cplx ? 8 : 4;
  if ((((uintptr_t)g.a | (uintptr_t)g.b) & 15) || ((uintptr_t)g.c & (es - 1))) return false;
  if (((g.lda * es) % 16) || ((g.ldb * es) % 16)) return false;
  if (g.m >= (1ll << 30) || g.n >= (1ll << 30) || g.k >= (1ll << 30)) return false;
  return true;
}

template <class C, bool A_MN, bool B_MN, bool CPLX>
cudaError_t launch_variant(const DeviceGemm &g, cudaStream_t stream) {
  static bool configured = false;
  auto kern = sgemm_ws_kernel<C, A_MN, B_MN, CPLX>;
  constexpr size_t SMEM = C::smem_bytes(A_MN, B_MN);
  constexpr int EM = CPLX ? 2 : 1;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  CUtensorMap map_a, map_b;
  /* stored A is (A_MN ? m x k : k x m), stored B is (B_MN ? n x k : k x n), both column-major; in floats the
   * contiguous extent doubles for complex.  k-contiguous operands: one staging row = BK elements = 64 / 128 B */
  const CUtensorMapSwizzle stage_swz = CPLX ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  const uint64_t pa = (uint64_t)g.lda * 4 * EM, pb = (uint64_t)g.ldb * 4 * EM;
  bool ok = A_MN ? make_map_f32(&map_a, g.a, (uint64_t)g.m * EM, (uint64_t)g.k, pa, C::BM, C::BK, CU_TENSOR_MAP_SWIZZLE_NONE)
                 : make_map_f32(&map_a, g.a, (uint64_t)g.k * EM, (uint64_t)g.m, pa, C::BK * EM, C::BM / EM, stage_swz);
  ok = ok && (B_MN ? make_map_f32(&map_b, g.b, (uint64_t)g.n * EM, (uint64_t)g.k, pb, C::BN, C::BK, CU_TENSOR_MAP_SWIZZLE_NONE)
                   : make_map_f32(&map_b, g.b, (uint64_t)g.k * EM, (uint64_t)g.n, pb, C::BK * EM, C::BN / EM, stage_swz));
  if (!ok) return cudaErrorNotSupported;
  const int64_t tm = (g.m + C::BM / EM - 1) / (C::BM / EM), tn = (g.n + C::BN / EM - 1) / (C::BN / EM);
  const int64_t tiles = (g.tri && C::BM == C::BN) ? tri_tile_count(tm) : (g.tri && C::BM == 2 * C::BN) ? tri21_tile_count(g.tri, tm, tn) : tm * tn;
  const int64_t cap = (int64_t)sm_count() * C::MINB;
  const int grid = (int)(tiles < cap ? tiles : cap);
  const int vec_c = (((uintptr_t)g.c & 15) == 0) && ((g.ldc * 4 * EM) % 16 == 0);
  kern<<<grid, C::THREADS, SMEM, stream>>>(map_a, map_b, g, vec_c);
  return cudaGetLastError();
}

template <class C, bool CPLX = false>
cudaError_t launch(const DeviceGemm &g, cudaStream_t stream) {
  if (!eligible(g) || (g.dtype == B200_C) != CPLX) return cudaErrorNotSupported;
  if (g.tri && g.m != g.n) return cudaErrorNotSupported;
  const bool a_mn = !(g.transa & 1), b_mn = (g.transb & 1);
  if (a_mn && b_mn) return launch_variant<C, true, true, CPLX>(g, stream);
  if (a_mn && !b_mn) return launch_variant<C, true, false, CPLX>(g, stream);
  if (!a_mn && b_mn) return launch_variant<C, false, true, CPLX>(g, stream);
  return launch_variant<C, false, false, CPLX>(g, stream);
}

}  // namespace sws
}  // namespace b200
