/*
 * zgemm_dmma.cu -- ZGEMM on the FP64 tensor pipe (DMMA.8x8x4).
 *
 * Stands in for the reference's zgemm_kernel_4x2_skylakex.c and its four conjugation variants
 * (_n/_l/_r/_b built with -DNN/-DCN/-DNC/-DCC, kernel/Makefile.L3:877-916, chosen at
 * level3.c:80-93) plus the zgemm_{i,o}{n,t}copy packing and zgemm_beta.  Complex operands stay
 * INTERLEAVED (re,im) in global and shared memory -- there is no planar re-layout pass; a lane
 * loads one complex number (16 bytes) per fragment element and the four real products of a
 * complex multiply-accumulate become four DMMAs:
 *
 *     re += ar*br ;  re += (-ai')*bi' ;  im += ar*bi' ;  im += ai'*br
 *
 * where ai' / bi' carry the conjugation sign (op codes 2,3), applied to the fragment in
 * registers, so all 16 op combinations share one of four kernels (stored orientation of A
 * and B), 8*m*n*k real flops all on the DMMA pipe.
 *
 * CTA tile 64 (m) x 128 (n) complex, k step 8, 8 warps as 2 x 4 with 32 x 32 warp tiles,
 * 4-stage cp.async ring.  As in dgemm_dmma.cu the MMA runs on the transposed tile so a lane
 * owns two adjacent rows of a column of C.  Shared layouts in units of one complex (16 B):
 * S[k][mn] with row stride = 2 mod 8 (66 / 130) or S[mn][k] with row stride 12 (= 4 mod 8):
 * the 8 lanes of a quarter-warp (2 idx x 4 k) then hit 8 distinct 16-byte bank groups.
 */
#include "gemm_common.cuh"
#include "async_copy.cuh"
#include <cstdlib>

namespace b200 {
namespace {

constexpr int BM = 64, BN = 128, BK = 8;
constexpr int STAGES = 4;
constexpr int THREADS = 256;
constexpr int WM = 32, WN = 32;
constexpr int LDA_MN = BM + 2, LDB_MN = BN + 2, LD_K = BK + 4;
constexpr int A_ELEMS = (BK * LDA_MN > BM * LD_K) ? BK * LDA_MN : BM * LD_K;   /* 768  */
constexpr int B_ELEMS = (BK * LDB_MN > BN * LD_K) ? BK * LDB_MN : BN * LD_K;   /* 1536 */
constexpr int STAGE_ELEMS = A_ELEMS + B_ELEMS;
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_ELEMS * sizeof(double2);  /* 144 KB */

__device__ __forceinline__ double xor_sign(double x, int mask) {
  return __hiloint2double(__double2hiint(x) ^ mask, __double2loint(x));
}

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(THREADS, 1)
zgemm_dmma_kernel(DeviceGemm g) {
  extern __shared__ __align__(16) double2 zsmem[];
  const double2 *__restrict__ A = (const double2 *)g.a;
  const double2 *__restrict__ B = (const double2 *)g.b;
  double2 *__restrict__ C = (double2 *)g.c;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp & 1) * WM, wn = (warp >> 1) * WN;
  const int fi = lane >> 2, fk = lane & 3;
  /* conjugation = flipping the sign bit of the imaginary part: done with integer XOR on the
   * high word so that the FP64 pipe only ever sees DMMAs */
  const int flip_a = (g.transa & 2) ? (int)0x80000000 : 0;
  const int flip_b = (g.transb & 2) ? (int)0x80000000 : 0;

  const int64_t tiles_m = (g.m + BM - 1) / BM, tiles_n = (g.n + BN - 1) / BN;
  const int64_t tiles = tiles_m * tiles_n;
  const int64_t ktiles = (g.k + BK - 1) / BK;

  const int a_off = A_MN ? (fk * LDA_MN + wm + fi) : ((wm + fi) * LD_K + fk);
  const int b_off = B_MN ? (fk * LDB_MN + wn + fi) : ((wn + fi) * LD_K + fk);
  constexpr int A_MT = A_MN ? 8 : 8 * LD_K;
  constexpr int B_NT = B_MN ? 8 : 8 * LD_K;
  constexpr int A_K4 = A_MN ? 4 * LDA_MN : 4;
  constexpr int B_K4 = B_MN ? 4 * LDB_MN : 4;

  /* one complex = one 16-byte chunk; all address arithmetic hoisted out of the k loop */
  using LoadA = TileLoader<A_MN, 16, BM, BK, A_MN ? LDA_MN : LD_K, THREADS>;
  using LoadB = TileLoader<B_MN, 16, BN, BK, B_MN ? LDB_MN : LD_K, THREADS>;
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(zsmem);
  const int k_tail = (int)(g.k % BK);

  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    int64_t bm, bn;
    banded_tile_coords<32>(t, tiles_m, tiles_n, bm, bn);
    const int64_t m0 = bm * BM, n0 = bn * BN;

    LoadA la; LoadB lb;
    la.init(A, g.lda, m0, g.m, tid);
    lb.init(B, g.ldb, n0, g.n, tid);
    int load_slot = 0;
    auto load_stage = [&](int64_t kt_load) {
      const uint32_t sa = smem_base + (uint32_t)(load_slot * STAGE_ELEMS * 16), sb = sa + (uint32_t)(A_ELEMS * 16);
      load_slot = (load_slot + 1 == STAGES) ? 0 : load_slot + 1;
      if (k_tail != 0 && kt_load == ktiles - 1) { la.issue_tail(sa, k_tail); lb.issue_tail(sb, k_tail); }
      else { la.issue(sa); lb.issue(sb); }
      la.advance(); lb.advance();
    };

    double re[4][4][2], im[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) re[i][j][0] = re[i][j][1] = im[i][j][0] = im[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
      if (s < ktiles) load_stage(s);
      cp_async_commit();
    }

    for (int64_t kt = 0; kt < ktiles; kt++) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      {
        int64_t nk = kt + STAGES - 1;
        if (nk < ktiles) load_stage(nk);
        cp_async_commit();
      }
      const int slot = (int)(kt % STAGES);   /* STAGES is a power of two */
      const double2 *sa = zsmem + slot * STAGE_ELEMS + a_off;
      const double2 *sb = zsmem + slot * STAGE_ELEMS + A_ELEMS + b_off;
#pragma unroll
      for (int k4 = 0; k4 < BK / 4; k4++) {
        double ar[4], ai[4], nai[4], br[4], bi[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          double2 v = sa[k4 * A_K4 + i * A_MT];
          ar[i] = v.x; ai[i] = xor_sign(v.y, flip_a); nai[i] = xor_sign(ai[i], (int)0x80000000);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
          double2 v = sb[k4 * B_K4 + j * B_NT];
          br[j] = v.x; bi[j] = xor_sign(v.y, flip_b);
        }
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) {
            dmma884(re[i][j][0], re[i][j][1], br[j], ar[i]);
            dmma884(im[i][j][0], im[i][j][1], bi[j], ar[i]);
            dmma884(re[i][j][0], re[i][j][1], bi[j], nai[i]);
            dmma884(im[i][j][0], im[i][j][1], br[j], ai[i]);
          }
      }
    }
    cp_async_wait<0>();
    __syncthreads();

    const double alr = g.alpha_re, ali = g.alpha_im, ber = g.beta_re, bei = g.beta_im;
    const bool use_beta = !(ber == 0.0 && bei == 0.0);
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int64_t n = n0 + wn + 8 * j + fi;
      if (n >= g.n) continue;
#pragma unroll
      for (int i = 0; i < 4; i++) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int64_t m = m0 + wm + 8 * i + 2 * fk + h;
          if (m >= g.m) continue;
          double2 *p = C + m + n * g.ldc;
          double xr = re[i][j][h], xi = im[i][j][h];
          double2 out;
          out.x = alr * xr - ali * xi;
          out.y = alr * xi + ali * xr;
          if (use_beta) {
            double2 old = *p;
            out.x += ber * old.x - bei * old.y;
            out.y += ber * old.y + bei * old.x;
          }
          *p = out;
        }
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * Producer-warp variant (the default): same structure as dgemm_dmma.cu's -- two producer warps
 * (one per operand: bulk copies for mn-contiguous tiles, cp.async for k-contiguous and edge tiles,
 * both reporting to full[stage] mbarriers) feed 8 consumer warps that only execute LDS + DMMA and
 * release stages through empty[stage]; no CTA-wide barrier in the main loop.  k step 16
 * (k-contiguous rows of 256 bytes), 3 stages. */
namespace pw {
constexpr int BK = 16, STAGES = 3;
constexpr int LDA_MN = BM + 2, LDB_MN = BN + 2, LD_K = BK + 4;                /* = 2 mod 8 / = 4 mod 8 */
constexpr int A_ELEMS = (BK * LDA_MN > BM * LD_K) ? BK * LDA_MN : BM * LD_K;   /* 1280 */
constexpr int B_ELEMS = (BK * LDB_MN > BN * LD_K) ? BK * LDB_MN : BN * LD_K;   /* 2560 */
constexpr int STAGE_ELEMS = A_ELEMS + B_ELEMS;
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_ELEMS * sizeof(double2) + 64;   /* 180 KB + barriers */
static_assert(SMEM_BYTES <= 227 * 1024, "ring does not fit");
}  // namespace pw

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(THREADS + 64, 1)
zgemm_dmma_pw_kernel(DeviceGemm g) {
  constexpr int BK = pw::BK, STAGES = pw::STAGES, LDA_MN = pw::LDA_MN, LDB_MN = pw::LDB_MN, LD_K = pw::LD_K;
  constexpr int A_ELEMS = pw::A_ELEMS, STAGE_ELEMS = pw::STAGE_ELEMS;
  extern __shared__ __align__(16) double2 zsmem[];
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(zsmem);
  const uint32_t bars = smem_base + (uint32_t)(STAGES * STAGE_ELEMS * 16);
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  const double2 *__restrict__ A = (const double2 *)g.a;
  const double2 *__restrict__ B = (const double2 *)g.b;
  double2 *__restrict__ C = (double2 *)g.c;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t tiles_m = (g.m + BM - 1) / BM, tiles_n = (g.n + BN - 1) / BN;
  /* triangle (SYRK family, GEMMT): only the tiles that touch it are enumerated (BN == 2 BM: tri12_tile_coords), so the
   * static round-robin over persistent CTAs stays balanced; round 1 walked all tiles and skipped */
  static_assert(BN == 2 * BM, "tri12 enumeration assumes tiles twice as wide as tall");
  const int64_t tiles = g.tri ? tri12_tile_count(g.tri, tiles_m, tiles_n) : tiles_m * tiles_n;
  const int64_t ktiles = (g.k + BK - 1) / BK;

  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(full_bar(s), 64); mbar_init(empty_bar(s), THREADS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp >= THREADS / 32) {
    const bool feeds_a = warp == THREADS / 32;
    int slot = 0; uint32_t phase = 0;
    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
      int64_t bm, bn;
      if (g.tri) tri12_tile_coords(t, g.tri, bm, bn); else banded_tile_coords<32>(t, tiles_m, tiles_n, bm, bn);
      const int64_t m0 = bm * BM, n0 = bn * BN;
      if (tri_outside(g.tri, m0, BM, n0, BN)) continue;      /* (never on the triangle's own enumeration) */
      for (int64_t kt = 0; kt < ktiles; kt++) {
        mbar_wait(empty_bar(slot), phase ^ 1);
        const uint32_t sa = smem_base + (uint32_t)(slot * STAGE_ELEMS * 16), sb = sa + (uint32_t)(A_ELEMS * 16);
        if (feeds_a) produce_operand<16, A_MN, BM, BK, LDA_MN, LD_K>(sa, A, g.lda, m0, kt * BK, g.m, g.k, full_bar(slot), lane);
        else         produce_operand<16, B_MN, BN, BK, LDB_MN, LD_K>(sb, B, g.ldb, n0, kt * BK, g.n, g.k, full_bar(slot), lane);
        __syncwarp();
        if (++slot == STAGES) { slot = 0; phase ^= 1; }
      }
    }
    return;
  }

  const int wm = (warp & 1) * WM, wn = (warp >> 1) * WN;
  const int fi = lane >> 2, fk = lane & 3;
  const int flip_a = (g.transa & 2) ? (int)0x80000000 : 0;
  const int flip_b = (g.transb & 2) ? (int)0x80000000 : 0;
  const int a_off = A_MN ? (fk * LDA_MN + wm + fi) : ((wm + fi) * LD_K + fk);
  const int b_off = B_MN ? (fk * LDB_MN + wn + fi) : ((wn + fi) * LD_K + fk);
  constexpr int A_MT = A_MN ? 8 : 8 * LD_K;
  constexpr int B_NT = B_MN ? 8 : 8 * LD_K;
  constexpr int A_K4 = A_MN ? 4 * LDA_MN : 4;
  constexpr int B_K4 = B_MN ? 4 * LDB_MN : 4;

  int slot = 0; uint32_t phase = 0;
  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    int64_t bm, bn;
    if (g.tri) tri12_tile_coords(t, g.tri, bm, bn); else banded_tile_coords<32>(t, tiles_m, tiles_n, bm, bn);
    const int64_t m0 = bm * BM, n0 = bn * BN;
    if (tri_outside(g.tri, m0, BM, n0, BN)) continue;

    double re[4][4][2], im[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) re[i][j][0] = re[i][j][1] = im[i][j][0] = im[i][j][1] = 0.0;

    for (int64_t kt = 0; kt < ktiles; kt++) {
      mbar_wait(full_bar(slot), phase);
      const double2 *sa = zsmem + slot * STAGE_ELEMS + a_off;
      const double2 *sb = zsmem + slot * STAGE_ELEMS + A_ELEMS + b_off;
#pragma unroll
      for (int k4 = 0; k4 < BK / 4; k4++) {
        double ar[4], ai[4], nai[4], br[4], bi[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          double2 v = sa[k4 * A_K4 + i * A_MT];
          ar[i] = v.x; ai[i] = xor_sign(v.y, flip_a); nai[i] = xor_sign(ai[i], (int)0x80000000);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
          double2 v = sb[k4 * B_K4 + j * B_NT];
          br[j] = v.x; bi[j] = xor_sign(v.y, flip_b);
        }
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) {
            dmma884(re[i][j][0], re[i][j][1], br[j], ar[i]);
            dmma884(im[i][j][0], im[i][j][1], bi[j], ar[i]);
            dmma884(re[i][j][0], re[i][j][1], bi[j], nai[i]);
            dmma884(im[i][j][0], im[i][j][1], br[j], ai[i]);
          }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_bar(slot));
      if (++slot == STAGES) { slot = 0; phase ^= 1; }
    }

    const double alr = g.alpha_re, ali = g.alpha_im, ber = g.beta_re, bei = g.beta_im;
    const bool use_beta = !(ber == 0.0 && bei == 0.0);
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int64_t n = n0 + wn + 8 * j + fi;
      if (n >= g.n) continue;
#pragma unroll
      for (int i = 0; i < 4; i++) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int64_t m = m0 + wm + 8 * i + 2 * fk + h;
          if (m >= g.m || !tri_keep(g.tri, m, n)) continue;
          double2 *p = C + m + n * g.ldc;
          double xr = re[i][j][h], xi = im[i][j][h];
          double2 out;
          out.x = alr * xr - ali * xi;
          out.y = alr * xi + ali * xr;
          if (use_beta) {
            double2 old = *p;
            out.x += ber * old.x - bei * old.y;
            out.y += ber * old.y + bei * old.x;
          }
          *p = out;
        }
      }
    }
  }
}

template <bool A_MN, bool B_MN>
cudaError_t launch_pw_variant(const DeviceGemm &g, cudaStream_t stream) {
  static bool configured = false;
  auto kern = zgemm_dmma_pw_kernel<A_MN, B_MN>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pw::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  int64_t tiles = ((g.m + BM - 1) / BM) * ((g.n + BN - 1) / BN);
  int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  kern<<<grid, THREADS + 64, pw::SMEM_BYTES, stream>>>(g);
  return cudaGetLastError();
}

template <bool A_MN, bool B_MN>
cudaError_t launch_variant(const DeviceGemm &g, cudaStream_t stream) {
  static bool configured = false;
  auto kern = zgemm_dmma_kernel<A_MN, B_MN>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  int64_t tiles = ((g.m + BM - 1) / BM) * ((g.n + BN - 1) / BN);
  int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  kern<<<grid, THREADS, SMEM_BYTES, stream>>>(g);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_zgemm_dmma(const DeviceGemm &g, cudaStream_t stream) {
  if (g.dtype != B200_Z) return cudaErrorNotSupported;
  /* one complex = one 16-byte cp.async: needs 16-byte aligned bases (any ld) */
  if (((uintptr_t)g.a | (uintptr_t)g.b | (uintptr_t)g.c) & 15) return cudaErrorNotSupported;
  const bool a_mn = !(g.transa & 1), b_mn = (g.transb & 1);
  static int cfg = -1;
  if (cfg < 0) { const char *ev = getenv("B200_ZGEMM_CFG"); cfg = ev ? atoi(ev) : 1; }
  cudaError_t e;
  if (g.tri && cfg != 1) return cudaErrorNotSupported;       /* only the producer-warp kernel masks */
  if (cfg == 1) {
    if (a_mn && b_mn) e = launch_pw_variant<true, true>(g, stream);
    else if (a_mn && !b_mn) e = launch_pw_variant<true, false>(g, stream);
    else if (!a_mn && b_mn) e = launch_pw_variant<false, true>(g, stream);
    else e = launch_pw_variant<false, false>(g, stream);
    if (e == cudaSuccess) count_launch("zgemm_dmma_pw_64x128x16");
    return e;
  }
  if (a_mn && b_mn) e = launch_variant<true, true>(g, stream);
  else if (a_mn && !b_mn) e = launch_variant<true, false>(g, stream);
  else if (!a_mn && b_mn) e = launch_variant<false, true>(g, stream);
  else e = launch_variant<false, false>(g, stream);
  if (e == cudaSuccess) count_launch("zgemm_dmma_64x128x8");
  return e;
}

}  // namespace b200
