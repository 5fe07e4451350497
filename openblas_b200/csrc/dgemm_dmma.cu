/*
 * dgemm_dmma.cu -- DGEMM on the FP64 tensor pipe (DMMA.8x8x4, PTX mma.sync.m8n8k4.f64).
 *
 * Replaces, for double precision, the reference's whole level-3 inner machinery:
 *   driver/level3/level3.c:288-406      GEMM_R/Q/P loop nest          -> persistent CTAs walking
 *                                                                        128x128 C tiles in an
 *                                                                        L2-friendly order
 *   ?gemm_itcopy/incopy/oncopy/otcopy   packing (level3.c:62-78)      -> none: cp.async lands A
 *                                                                        and B tiles in padded
 *                                                                        shared memory in their
 *                                                                        STORED orientation; op()
 *                                                                        is only a different
 *                                                                        fragment address
 *   dgemm_kernel_16x2_skylakex.c        AVX-512 register tile         -> 64x32 warp tile of
 *                                                                        DMMA 8x8x4 fragments
 *   dgemm_beta_skylakex.c               separate C = beta*C pass      -> fused epilogue; beta == 0
 *                                                                        never reads C
 *
 * The MMA is issued on the TRANSPOSED tile (MMA "A" operand = op(B)^T fragment, MMA "B"
 * operand = op(A)^T fragment) so that the two accumulator values a lane owns are adjacent
 * rows of column-major C and can be stored as one 16-byte word.
 *
 * Shared-memory layouts (doubles), chosen so every fragment load is bank-conflict free
 * (a lane reads X[idx = lane/4][k = lane%4]; a half-warp covers 4 idx x 4 k):
 *   stored "mn-contiguous" (A not transposed / B transposed):  S[k][mn], row stride 132
 *   stored "k-contiguous"  (A transposed / B not transposed):  S[mn][k], row stride 20
 * Both strides are 4 mod 16, which maps the 16 (idx,k) pairs onto 16 distinct 8-byte banks.
 *
 * Results are deterministic: one CTA owns a C tile for the whole k range, fixed k order.
 */
#include "gemm_common.cuh"
#include "async_copy.cuh"
#include <cstdlib>

namespace b200 {
namespace {


/* Tile configuration.  DMMA.8x8x4 can be re-issued by the same warp only every ~32 cycles while
 * the pipe needs 16 per instruction (measured: ncu shows the warp parked on the NOP that follows
 * every DMMA), so two warps per scheduler can only just saturate the pipe and every LDS / barrier
 * stall shows up as idle pipe time (80 % with 8 warps of 64x32).  Four warps per scheduler with
 * 32x32 warp tiles (<= 128 registers) leave slack to hide them. */
template <int BM_, int BN_, int WM_, int WN_, int BK_, int STAGES_, int MINB_>
struct Cfg {
  static constexpr int BM = BM_, BN = BN_, WM = WM_, WN = WN_, BK = BK_, STAGES = STAGES_, MINB = MINB_;
  static constexpr int LD_K = BK + 4;                     /* S[mn][k] row stride, = 4 mod 16 */
  static constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN;
  static constexpr int THREADS = WARPS_M * WARPS_N * 32;
  static constexpr int FM = WM / 8, FN = WN / 8;          /* 8x8 fragments per warp tile */
  static constexpr int LDA_MN = BM + 4, LDB_MN = BN + 4;  /* S[k][mn] row strides, = 4 mod 16 */
  static constexpr int A_DOUBLES = (BK * LDA_MN > BM * LD_K) ? BK * LDA_MN : BM * LD_K;
  static constexpr int B_DOUBLES = (BK * LDB_MN > BN * LD_K) ? BK * LDB_MN : BN * LD_K;
  static constexpr int STAGE_DOUBLES = A_DOUBLES + B_DOUBLES;
  static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_DOUBLES * sizeof(double);
  static_assert(BM % 16 == 0 && BN % 16 == 0 && BK % 16 == 0, "strides must stay 4 mod 16");
  static_assert(SMEM_BYTES <= 227 * 1024, "tile ring does not fit in shared memory");
};

/* Element-wise fallback loader for operands whose base or leading dimension is not 16-byte
 * aligned (8-byte cp.async per element, bounds checked per element).  The aligned case uses
 * TileLoader (async_copy.cuh).  MN_CONTIG: element (mn, k) lives at g[mn + k*ld], else at
 * g[k + mn*ld]. */
template <bool MN_CONTIG, int ROWS, int BK, int LD_MN, int THREADS>
__device__ __noinline__ void load_tile_unaligned(double *s, const double *__restrict__ g, int64_t ld, int64_t mn0,
                                                 int64_t k0, int64_t mn_end, int64_t k_end, int tid) {
  for (int i = 0; i < ROWS * BK / THREADS; i++) {
    int idx = tid + i * THREADS;
    int k = MN_CONTIG ? idx / ROWS : idx % BK, mn = MN_CONTIG ? idx % ROWS : idx / BK;
    int64_t gk = k0 + k, gmn = mn0 + mn;
    int bytes = (gk < k_end && gmn < mn_end) ? 8 : 0;
    const double *src = bytes ? (MN_CONTIG ? g + gmn + gk * ld : g + gk + gmn * ld) : g;
    cp_async8(MN_CONTIG ? s + k * LD_MN + mn : s + mn * (BK + 4) + k, src, bytes);
  }
}

/* A_MN: op(A) tile is stored mn-contiguous (A not transposed); B_MN: B transposed. */
template <class C_, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(C_::THREADS, C_::MINB)
dgemm_dmma_kernel(DeviceGemm g, int vec_a, int vec_b, int vec_c) {
  constexpr int BM = C_::BM, BN = C_::BN, BK = C_::BK, LD_K = C_::LD_K, STAGES = C_::STAGES, THREADS = C_::THREADS;
  constexpr int FM = C_::FM, FN = C_::FN;
  constexpr int LDA_MN = C_::LDA_MN, LDB_MN = C_::LDB_MN;
  constexpr int A_DOUBLES = C_::A_DOUBLES, STAGE_DOUBLES = C_::STAGE_DOUBLES;
  extern __shared__ __align__(16) double smem[];
  const double *__restrict__ A = (const double *)g.a;
  const double *__restrict__ B = (const double *)g.b;
  double *__restrict__ C = (double *)g.c;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp % C_::WARPS_M) * C_::WM, wn = (warp / C_::WARPS_M) * C_::WN;
  const int fi = lane >> 2, fk = lane & 3;          /* fragment index / k within a k4 step */

  const int64_t tiles_m = (g.m + BM - 1) / BM, tiles_n = (g.n + BN - 1) / BN;
  const int64_t tiles = tiles_m * tiles_n;
  const int64_t ktiles = (g.k + BK - 1) / BK;
  const double alpha = g.alpha_re, beta = g.beta_re;
  const bool use_beta = beta != 0.0;

  /* per-lane fragment offsets inside an operand tile */
  const int a_off = A_MN ? (fk * LDA_MN + wm + fi) : ((wm + fi) * LD_K + fk);
  const int b_off = B_MN ? (fk * LDB_MN + wn + fi) : ((wn + fi) * LD_K + fk);
  constexpr int A_MT = A_MN ? 8 : 8 * LD_K;          /* step to the next 8-row m fragment */
  constexpr int B_NT = B_MN ? 8 : 8 * LD_K;
  constexpr int A_K4 = A_MN ? 4 * LDA_MN : 4;        /* step to the next k4 slice */
  constexpr int B_K4 = B_MN ? 4 * LDB_MN : 4;

  using LoadA = TileLoader<A_MN, 8, BM, BK, A_MN ? LDA_MN : LD_K, THREADS>;
  using LoadB = TileLoader<B_MN, 8, BN, BK, B_MN ? LDB_MN : LD_K, THREADS>;
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem);
  const int k_tail = (int)(g.k % BK);               /* != 0: the last k tile is partial */

  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    int64_t bm, bn;
    banded_tile_coords<16>(t, tiles_m, tiles_n, bm, bn);
    const int64_t m0 = bm * BM, n0 = bn * BN;

    LoadA la; LoadB lb;
    if (vec_a) la.init(A, g.lda, m0, g.m, tid);
    if (vec_b) lb.init(B, g.ldb, n0, g.n, tid);
    /* k tiles are requested strictly in order, so the loaders just advance */
    int load_slot = 0;                               /* ring slot of the next k tile to request */
    auto load_stage = [&](int64_t kt_load) {
      const int stage = load_slot;
      load_slot = (load_slot + 1 == STAGES) ? 0 : load_slot + 1;
      const bool tail = k_tail != 0 && kt_load == ktiles - 1;
      const uint32_t sa = smem_base + (uint32_t)(stage * STAGE_DOUBLES * 8), sb = sa + (uint32_t)(A_DOUBLES * 8);
      if (vec_a) { if (tail) la.issue_tail(sa, k_tail); else la.issue(sa); la.advance(); }
      else load_tile_unaligned<A_MN, BM, BK, LDA_MN, THREADS>(smem + stage * STAGE_DOUBLES, A, g.lda, m0, kt_load * BK, g.m, g.k, tid);
      if (vec_b) { if (tail) lb.issue_tail(sb, k_tail); else lb.issue(sb); lb.advance(); }
      else load_tile_unaligned<B_MN, BN, BK, LDB_MN, THREADS>(smem + stage * STAGE_DOUBLES + A_DOUBLES, B, g.ldb, n0, kt_load * BK, g.n, g.k, tid);
    };

    double acc[FM][FN][2];
#pragma unroll
    for (int i = 0; i < FM; i++)
#pragma unroll
      for (int j = 0; j < FN; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    /* prologue: STAGES-1 tiles in flight */
#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
      if (s < ktiles) load_stage(s);
      cp_async_commit();
    }

    int slot = 0;                                    /* ring slot of the k tile being consumed */
    for (int64_t kt = 0; kt < ktiles; kt++) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      {
        /* refill the stage consumed in the previous iteration */
        int64_t nk = kt + STAGES - 1;
        if (nk < ktiles) load_stage(nk);
        cp_async_commit();
      }
      const double *sa = smem + slot * STAGE_DOUBLES + a_off;
      const double *sb = smem + slot * STAGE_DOUBLES + A_DOUBLES + b_off;
      slot = (slot + 1 == STAGES) ? 0 : slot + 1;
#pragma unroll
      for (int k4 = 0; k4 < BK / 4; k4++) {
        double af[FM], bf[FN];
#pragma unroll
        for (int i = 0; i < FM; i++) af[i] = sa[k4 * A_K4 + i * A_MT];
#pragma unroll
        for (int j = 0; j < FN; j++) bf[j] = sb[k4 * B_K4 + j * B_NT];
#pragma unroll
        for (int i = 0; i < FM; i++)
#pragma unroll
          for (int j = 0; j < FN; j++) dmma884(acc[i][j][0], acc[i][j][1], bf[j], af[i]);
      }
    }
    cp_async_wait<0>();
    __syncthreads();   /* all warps done with the last stages before the next tile's prologue */

    /* epilogue: lane owns C[m .. m+1][n], m = m0+wm+8i+2*fk, n = n0+wn+8j+fi */
#pragma unroll
    for (int j = 0; j < FN; j++) {
      const int64_t n = n0 + wn + 8 * j + fi;
      if (n >= g.n) continue;
#pragma unroll
      for (int i = 0; i < FM; i++) {
        const int64_t m = m0 + wm + 8 * i + 2 * fk;
        if (m >= g.m) continue;
        double *p = C + m + n * g.ldc;
        double r0 = alpha * acc[i][j][0], r1 = alpha * acc[i][j][1];
        if (vec_c && m + 1 < g.m) {
          if (use_beta) {
            double2 old = *reinterpret_cast<const double2 *>(p);
            r0 = fma(beta, old.x, r0); r1 = fma(beta, old.y, r1);
          }
          *reinterpret_cast<double2 *>(p) = make_double2(r0, r1);
        } else {
          if (use_beta) r0 = fma(beta, p[0], r0);
          p[0] = r0;
          if (m + 1 < g.m) {
            if (use_beta) r1 = fma(beta, p[1], r1);
            p[1] = r1;
          }
        }
      }
    }
  }
}

template <class C_, bool A_MN, bool B_MN>
cudaError_t launch_variant(const DeviceGemm &g, cudaStream_t stream, int vec_a, int vec_b, int vec_c) {
  static bool configured = false;   /* per instantiation; benign race: same value every time */
  auto kern = dgemm_dmma_kernel<C_, A_MN, B_MN>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C_::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  int64_t tiles = ((g.m + C_::BM - 1) / C_::BM) * ((g.n + C_::BN - 1) / C_::BN);
  int64_t cap = (int64_t)sm_count() * C_::MINB;
  int grid = (int)(tiles < cap ? tiles : cap);
  kern<<<grid, C_::THREADS, C_::SMEM_BYTES, stream>>>(g, vec_a, vec_b, vec_c);
  return cudaGetLastError();
}

template <class C_>
cudaError_t launch_cfg(const DeviceGemm &g, cudaStream_t stream) {
  const bool a_mn = !(g.transa & 1), b_mn = (g.transb & 1);
  const int vec_a = (((uintptr_t)g.a & 15) == 0) && (g.lda % 2 == 0);
  const int vec_b = (((uintptr_t)g.b & 15) == 0) && (g.ldb % 2 == 0);
  const int vec_c = (((uintptr_t)g.c & 15) == 0) && (g.ldc % 2 == 0);
  if (a_mn && b_mn) return launch_variant<C_, true, true>(g, stream, vec_a, vec_b, vec_c);
  if (a_mn && !b_mn) return launch_variant<C_, true, false>(g, stream, vec_a, vec_b, vec_c);
  if (!a_mn && b_mn) return launch_variant<C_, false, true>(g, stream, vec_a, vec_b, vec_c);
  return launch_variant<C_, false, false>(g, stream, vec_a, vec_b, vec_c);
}

/* ------------------------------------------------------------------------------------------
 * Producer-warp variant (any m, n, k; A and B columns 16-byte aligned): the cp.async ring above costs every thread 8 LDGSTS plus their
 * address registers per k tile and a CTA-wide barrier per k tile (ncu: stall_barrier 3 %,
 * long_scoreboard 4 % on the address registers, short_scoreboard 6 %).  Here ONE extra warp is the
 * producer: each of its lanes copies whole tile rows with the TMA engine's 1-D bulk copy
 * (cp.async.bulk, SASS UBLKCP) into the same padded, conflict-free shared layout and completion
 * is tracked by mbarriers (full[stage], empty[stage] with one arrival per consumer warp), so the
 * 8 DMMA warps never execute a load instruction for global memory, never hit a CTA-wide barrier
 * in the main loop, and the producers run ahead across C tiles (the next tile's first stages are in
 * flight during the epilogue).  Two producer warps: one per operand (see produce_operand). */
template <class C_, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(C_::THREADS + 64, C_::MINB)
dgemm_dmma_bulk_kernel(DeviceGemm g, int vec_c, int probe_noload) {
  constexpr int BM = C_::BM, BN = C_::BN, BK = C_::BK, LD_K = C_::LD_K, STAGES = C_::STAGES;
  constexpr int FM = C_::FM, FN = C_::FN;
  constexpr int LDA_MN = C_::LDA_MN, LDB_MN = C_::LDB_MN;
  constexpr int A_DOUBLES = C_::A_DOUBLES, STAGE_DOUBLES = C_::STAGE_DOUBLES;
  constexpr uint32_t FULL_ARRIVALS = 64;   /* every lane of both producer warps arrives once per stage */
  extern __shared__ __align__(16) double smem[];
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t bars = smem_base + (uint32_t)(STAGES * STAGE_DOUBLES * 8);
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };

  const double *__restrict__ A = (const double *)g.a;
  const double *__restrict__ B = (const double *)g.b;
  double *__restrict__ C = (double *)g.c;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t tiles_m = (g.m + BM - 1) / BM, tiles_n = (g.n + BN - 1) / BN;
  const int64_t tiles = g.tri ? tri_tile_count(tiles_m) : tiles_m * tiles_n;      /* tri: m == n, square tiles */
  const int64_t ktiles = (g.k + BK - 1) / BK;

  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(full_bar(s), FULL_ARRIVALS); mbar_init(empty_bar(s), C_::THREADS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp >= C_::THREADS / 32) {
    /* ============================== producer warps: warp 8 feeds A tiles, warp 9 feeds B tiles */
    const bool feeds_a = warp == C_::THREADS / 32;
    int slot = 0; uint32_t phase = 0;
    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
      int64_t bm, bn;
      if (g.tri) tri_tile_coords(t, g.tri, bm, bn); else banded_tile_coords<16>(t, tiles_m, tiles_n, bm, bn);
      const int64_t m0 = bm * BM, n0 = bn * BN;
      if (tri_outside(g.tri, m0, BM, n0, BN)) continue;      /* consumers skip the same tiles */
      for (int64_t kt = 0; kt < ktiles; kt++) {
        const int64_t k0 = kt * BK;
        mbar_wait(empty_bar(slot), phase ^ 1);
        const uint32_t sa = smem_base + (uint32_t)(slot * STAGE_DOUBLES * 8), sb = sa + (uint32_t)(A_DOUBLES * 8);
        if (probe_noload && kt >= STAGES) {           /* measurement probe: pure compute, stale tiles */
          mbar_arrive(full_bar(slot));
        } else if (feeds_a) {
          produce_operand<8, A_MN, BM, BK, LDA_MN, LD_K>(sa, A, g.lda, m0, k0, g.m, g.k, full_bar(slot), lane);
        } else {
          produce_operand<8, B_MN, BN, BK, LDB_MN, LD_K>(sb, B, g.ldb, n0, k0, g.n, g.k, full_bar(slot), lane);
        }
        __syncwarp();
        if (++slot == STAGES) { slot = 0; phase ^= 1; }
      }
    }
    return;
  }

  /* ================================================================= consumer warps */
  const int wm = (warp % C_::WARPS_M) * C_::WM, wn = (warp / C_::WARPS_M) * C_::WN;
  const int fi = lane >> 2, fk = lane & 3;
  const double alpha = g.alpha_re, beta = g.beta_re;
  const bool use_beta = beta != 0.0;
  const int a_off = A_MN ? (fk * LDA_MN + wm + fi) : ((wm + fi) * LD_K + fk);
  const int b_off = B_MN ? (fk * LDB_MN + wn + fi) : ((wn + fi) * LD_K + fk);
  constexpr int A_MT = A_MN ? 8 : 8 * LD_K;
  constexpr int B_NT = B_MN ? 8 : 8 * LD_K;
  constexpr int A_K4 = A_MN ? 4 * LDA_MN : 4;
  constexpr int B_K4 = B_MN ? 4 * LDB_MN : 4;

  int slot = 0; uint32_t phase = 0;
  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    int64_t bm, bn;
    if (g.tri) tri_tile_coords(t, g.tri, bm, bn); else banded_tile_coords<16>(t, tiles_m, tiles_n, bm, bn);
    const int64_t m0 = bm * BM, n0 = bn * BN;
    if (tri_outside(g.tri, m0, BM, n0, BN)) continue;
    const bool masked = tri_partial(g.tri, m0, BM, n0, BN);

    double acc[FM][FN][2];
#pragma unroll
    for (int i = 0; i < FM; i++)
#pragma unroll
      for (int j = 0; j < FN; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int64_t kt = 0; kt < ktiles; kt++) {
      mbar_wait(full_bar(slot), phase);
      const double *sa = smem + slot * STAGE_DOUBLES + a_off;
      const double *sb = smem + slot * STAGE_DOUBLES + A_DOUBLES + b_off;
#pragma unroll
      for (int k4 = 0; k4 < BK / 4; k4++) {
        double af[FM], bf[FN];
#pragma unroll
        for (int i = 0; i < FM; i++) af[i] = sa[k4 * A_K4 + i * A_MT];
#pragma unroll
        for (int j = 0; j < FN; j++) bf[j] = sb[k4 * B_K4 + j * B_NT];
#pragma unroll
        for (int i = 0; i < FM; i++)
#pragma unroll
          for (int j = 0; j < FN; j++) dmma884(acc[i][j][0], acc[i][j][1], bf[j], af[i]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_bar(slot));     /* this warp is done reading the stage */
      if (++slot == STAGES) { slot = 0; phase ^= 1; }
    }

    /* epilogue: lane owns C[m .. m+1][n]; 16-byte stores when ldc is even and C 16-byte aligned */
#pragma unroll
    for (int j = 0; j < FN; j++) {
      const int64_t n = n0 + wn + 8 * j + fi;
      if (n >= g.n) continue;
#pragma unroll
      for (int i = 0; i < FM; i++) {
        const int64_t m = m0 + wm + 8 * i + 2 * fk;
        if (m >= g.m) continue;
        double *p = C + m + n * g.ldc;
        double r0 = alpha * acc[i][j][0], r1 = alpha * acc[i][j][1];
        if (vec_c && m + 1 < g.m && !masked) {
          if (use_beta) {
            double2 old = *reinterpret_cast<const double2 *>(p);
            r0 = fma(beta, old.x, r0); r1 = fma(beta, old.y, r1);
          }
          *reinterpret_cast<double2 *>(p) = make_double2(r0, r1);
        } else {
          if (tri_keep(g.tri, m, n)) {
            if (use_beta) r0 = fma(beta, p[0], r0);
            p[0] = r0;
          }
          if (m + 1 < g.m && tri_keep(g.tri, m + 1, n)) {
            if (use_beta) r1 = fma(beta, p[1], r1);
            p[1] = r1;
          }
        }
      }
    }
  }
}

template <class C_, bool A_MN, bool B_MN>
cudaError_t launch_bulk_variant(const DeviceGemm &g, cudaStream_t stream) {
  static bool configured = false;
  auto kern = dgemm_dmma_bulk_kernel<C_, A_MN, B_MN>;
  constexpr size_t smem_bytes = C_::SMEM_BYTES + 64;
  static_assert(smem_bytes <= 227 * 1024, "bulk ring does not fit");
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  int64_t tiles = ((g.m + C_::BM - 1) / C_::BM) * ((g.n + C_::BN - 1) / C_::BN);
  if (g.tri) tiles = tri_tile_count((g.m + C_::BM - 1) / C_::BM);
  const int64_t cap = (int64_t)sm_count() * C_::MINB;
  int grid = (int)(tiles < cap ? tiles : cap);
  const int vec_c = (((uintptr_t)g.c & 15) == 0) && (g.ldc % 2 == 0);
  static int probe = -1;   /* B200_DGEMM_PROBE_NOLOAD=1: timing probe only, results are garbage */
  if (probe < 0) { const char *e = getenv("B200_DGEMM_PROBE_NOLOAD"); probe = e ? atoi(e) : 0; }
  kern<<<grid, C_::THREADS + 64, smem_bytes, stream>>>(g, vec_c, probe);
  return cudaGetLastError();
}

/* the producer-warp kernel needs 16-byte aligned A and B columns; any m, n, k */
template <class C_>
bool bulk_eligible(const DeviceGemm &g) {
  return g.lda % 2 == 0 && g.ldb % 2 == 0 && ((((uintptr_t)g.a | (uintptr_t)g.b) & 15) == 0);
}

template <class C_>
cudaError_t launch_bulk(const DeviceGemm &g, cudaStream_t stream) {
  const bool a_mn = !(g.transa & 1), b_mn = (g.transb & 1);
  if (a_mn && b_mn) return launch_bulk_variant<C_, true, true>(g, stream);
  if (a_mn && !b_mn) return launch_bulk_variant<C_, true, false>(g, stream);
  if (!a_mn && b_mn) return launch_bulk_variant<C_, false, true>(g, stream);
  return launch_bulk_variant<C_, false, false>(g, stream);
}

using CfgWide   = Cfg<128, 128, 64, 32, 16, 4, 1>;   /* 8 warps, 64x32 warp tiles, k step 16, 1 CTA/SM            */
using CfgDual   = Cfg<128, 64, 32, 32, 16, 3, 2>;    /* 8 warps, 32x32 warp tiles, 2 CTAs/SM: 16 warps per SM     */
using CfgBig    = Cfg<128, 128, 32, 32, 16, 4, 1>;   /* 16 warps, 32x32 warp tiles, 1 CTA/SM                      */
using CfgWide32 = Cfg<128, 128, 64, 32, 32, 3, 1>;   /* as Wide with k step 32: half as many barriers per flop    */
using CfgWide32b = Cfg<128, 128, 64, 32, 32, 2, 1>;  /* k step 32, double buffer                                  */
using CfgMid    = Cfg<64, 64, 32, 16, 32, 5, 1>;     /* producer-warp kernel for grids that leave SMs idle at 128x128 */
using CfgMid2   = Cfg<64, 64, 32, 16, 32, 3, 2>;     /* the same with two CTAs per SM: one CTA's pipeline fill and epilogue hide behind the other's DMMAs */

}  // namespace

cudaError_t launch_dgemm_dmma(const DeviceGemm &g, cudaStream_t stream) {
  if (g.dtype != B200_D) return cudaErrorNotSupported;
  if (((uintptr_t)g.a | (uintptr_t)g.b | (uintptr_t)g.c) & 7) return cudaErrorNotSupported;
  static int cfg = -1;
  if (cfg < 0) { const char *e = getenv("B200_DGEMM_CFG"); cfg = e ? atoi(e) : 5; }
  cudaError_t e;
  const char *name;
  if (g.tri && (g.m != g.n || !(cfg >= 5 && bulk_eligible<CfgWide32>(g)))) return cudaErrorNotSupported;   /* only the producer-warp kernels mask */
  if (cfg >= 5 && bulk_eligible<CfgWide32>(g)) {
    /* Tile choice.  A 128x128 tile costs four 64x64 tiles; the grid runs in waves of one tile per SM.
     * When the 128x128 tiling leaves SMs idle (few tiles, or a mostly empty last wave) the 64x64
     * tiling finishes sooner even though it re-reads operands twice as often (33 against 35.5 TFLOP/s
     * on large grids: penalty 1.08; 1024^3: 26 against 14.6 TFLOP/s, 2048^3: 32.7 against 30.4).  B200_DGEMM_TILE=64|128 forces one. */
    const char *t = getenv("B200_DGEMM_TILE");
    const int forced = t ? atoi(t) : 0;
    const int64_t sms = sm_count();
    int64_t t128 = ((g.m + 127) / 128) * ((g.n + 127) / 128), t64 = ((g.m + 63) / 64) * ((g.n + 63) / 64);
    if (g.tri) { t128 = (t128 + 1) / 2; t64 = (t64 + 1) / 2; }      /* about half the tiles are skipped */
    const double est128 = 4.0 * (double)((t128 + sms - 1) / sms), est64 = 1.08 * (double)((t64 + sms - 1) / sms);
    const bool small_tile = forced == 64 || (forced != 128 && est64 < est128);
    if (small_tile) {
      /* two CTAs per SM (3-stage rings) when every tile is resident at once: one CTA's pipeline fill and epilogue hide
       * behind the other's DMMAs (1024^3: 26.1 -> 27.2 TFLOP/s); with more tiles than 2 x SMs the deeper ring of the
       * one-CTA form wins (2048^3: 32.7 against 29.4).  B200_DGEMM_MID=1|2 forces one. */
      static int mid = -1;
      if (mid < 0) { const char *ev = getenv("B200_DGEMM_MID"); mid = ev ? atoi(ev) : 0; }
      if (mid == 2 || (mid == 0 && t64 <= 2 * sms)) {
        e = launch_bulk<CfgMid2>(g, stream);
        if (e == cudaSuccess) count_launch("dgemm_dmma_pw_64x64x32_w32x16_2cta");
        return e;
      }
      e = launch_bulk<CfgMid>(g, stream);
      if (e == cudaSuccess) count_launch("dgemm_dmma_pw_64x64x32_w32x16");
      return e;
    }
    e = launch_bulk<CfgWide32>(g, stream);
    if (e == cudaSuccess) count_launch("dgemm_dmma_bulk_128x128x32_w64x32");
    return e;
  }
  switch (cfg) {
    case 1:  e = launch_cfg<CfgDual>(g, stream); name = "dgemm_dmma_128x64x16_w32x32_2cta"; break;
    case 2:  e = launch_cfg<CfgBig>(g, stream);  name = "dgemm_dmma_128x128x16_w32x32"; break;
    case 0:  e = launch_cfg<CfgWide>(g, stream); name = "dgemm_dmma_128x128x16_w64x32"; break;
    case 4:  e = launch_cfg<CfgWide32b>(g, stream); name = "dgemm_dmma_128x128x32_w64x32_2stage"; break;
    default: e = launch_cfg<CfgWide32>(g, stream); name = "dgemm_dmma_128x128x32_w64x32"; break;
  }
  if (e == cudaSuccess) count_launch(name);
  return e;
}

}  // namespace b200
