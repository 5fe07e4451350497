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

namespace b200 {
namespace {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int STAGES = 4;
constexpr int THREADS = 256;                 /* 8 warps: 2 (m) x 4 (n), warp tile 64 x 32 */
constexpr int WM = 64, WN = 32;
constexpr int LD_MN = BM + 4;                /* S[k][mn] row stride */
constexpr int LD_K = BK + 4;                 /* S[mn][k] row stride */
constexpr int OPERAND_DOUBLES = (BK * LD_MN > BM * LD_K) ? BK * LD_MN : BM * LD_K;  /* 2560 */
constexpr int STAGE_DOUBLES = 2 * OPERAND_DOUBLES;
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_DOUBLES * sizeof(double);      /* 160 KB */

/* Copy one (128 mn) x (16 k) operand tile global -> shared, zero-filling everything outside
 * the matrix.  MN_CONTIG: element (mn, k) lives at g[mn + k*ld], else at g[k + mn*ld].
 * vec16: the operand's base and ld allow 16-byte copies. */
template <bool MN_CONTIG>
__device__ __forceinline__ void load_tile(double *s, const double *__restrict__ g, int64_t ld, int64_t mn0,
                                          int64_t k0, int64_t mn_end, int64_t k_end, bool vec16, int tid) {
  if (MN_CONTIG) {
    if (vec16) {
#pragma unroll
      for (int i = 0; i < (BM / 2) * BK / THREADS; i++) {   /* 4 */
        int idx = tid + i * THREADS;
        int k = idx / (BM / 2), mn = (idx % (BM / 2)) * 2;
        int64_t gk = k0 + k, gmn = mn0 + mn;
        int64_t left = (gk < k_end) ? (mn_end - gmn) : 0;
        int bytes = left >= 2 ? 16 : (left == 1 ? 8 : 0);
        const double *src = bytes ? g + gmn + gk * ld : g;
        cp_async16(s + k * LD_MN + mn, src, bytes);
      }
    } else {
#pragma unroll
      for (int i = 0; i < BM * BK / THREADS; i++) {         /* 8 */
        int idx = tid + i * THREADS;
        int k = idx / BM, mn = idx % BM;
        int64_t gk = k0 + k, gmn = mn0 + mn;
        int bytes = (gk < k_end && gmn < mn_end) ? 8 : 0;
        const double *src = bytes ? g + gmn + gk * ld : g;
        cp_async8(s + k * LD_MN + mn, src, bytes);
      }
    }
  } else {
    if (vec16) {
#pragma unroll
      for (int i = 0; i < BM * (BK / 2) / THREADS; i++) {   /* 4 */
        int idx = tid + i * THREADS;
        int k = (idx % (BK / 2)) * 2, mn = idx / (BK / 2);
        int64_t gk = k0 + k, gmn = mn0 + mn;
        int64_t left = (gmn < mn_end) ? (k_end - gk) : 0;
        int bytes = left >= 2 ? 16 : (left == 1 ? 8 : 0);
        const double *src = bytes ? g + gk + gmn * ld : g;
        cp_async16(s + mn * LD_K + k, src, bytes);
      }
    } else {
#pragma unroll
      for (int i = 0; i < BM * BK / THREADS; i++) {         /* 8 */
        int idx = tid + i * THREADS;
        int k = idx % BK, mn = idx / BK;
        int64_t gk = k0 + k, gmn = mn0 + mn;
        int bytes = (gk < k_end && gmn < mn_end) ? 8 : 0;
        const double *src = bytes ? g + gk + gmn * ld : g;
        cp_async8(s + mn * LD_K + k, src, bytes);
      }
    }
  }
}

/* A_MN: op(A) tile is stored mn-contiguous (A not transposed); B_MN: B transposed. */
template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(THREADS, 1)
dgemm_dmma_kernel(DeviceGemm g, int vec_a, int vec_b, int vec_c) {
  extern __shared__ __align__(16) double smem[];
  const double *__restrict__ A = (const double *)g.a;
  const double *__restrict__ B = (const double *)g.b;
  double *__restrict__ C = (double *)g.c;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp & 1) * WM, wn = (warp >> 1) * WN;
  const int fi = lane >> 2, fk = lane & 3;          /* fragment index / k within a k4 step */

  const int64_t tiles_m = (g.m + BM - 1) / BM, tiles_n = (g.n + BN - 1) / BN;
  const int64_t tiles = tiles_m * tiles_n;
  const int64_t ktiles = (g.k + BK - 1) / BK;
  const double alpha = g.alpha_re, beta = g.beta_re;
  const bool use_beta = beta != 0.0;

  /* per-lane fragment offsets inside an operand tile */
  const int a_off = A_MN ? (fk * LD_MN + wm + fi) : ((wm + fi) * LD_K + fk);
  const int b_off = B_MN ? (fk * LD_MN + wn + fi) : ((wn + fi) * LD_K + fk);
  constexpr int A_MT = A_MN ? 8 : 8 * LD_K;          /* step to the next 8-row m fragment */
  constexpr int B_NT = B_MN ? 8 : 8 * LD_K;
  constexpr int A_K4 = A_MN ? 4 * LD_MN : 4;         /* step to the next k4 slice */
  constexpr int B_K4 = B_MN ? 4 * LD_MN : 4;

  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    int64_t bm, bn;
    banded_tile_coords<16>(t, tiles_m, tiles_n, bm, bn);
    const int64_t m0 = bm * BM, n0 = bn * BN;

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    /* prologue: STAGES-1 tiles in flight */
#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
      if (s < ktiles) {
        double *sa = smem + s * STAGE_DOUBLES, *sb = sa + OPERAND_DOUBLES;
        load_tile<A_MN>(sa, A, g.lda, m0, (int64_t)s * BK, g.m, g.k, vec_a, tid);
        load_tile<B_MN>(sb, B, g.ldb, n0, (int64_t)s * BK, g.n, g.k, vec_b, tid);
      }
      cp_async_commit();
    }

    for (int64_t kt = 0; kt < ktiles; kt++) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      {
        /* refill the stage consumed in the previous iteration */
        int64_t nk = kt + STAGES - 1;
        if (nk < ktiles) {
          double *sa = smem + (nk % STAGES) * STAGE_DOUBLES, *sb = sa + OPERAND_DOUBLES;
          load_tile<A_MN>(sa, A, g.lda, m0, nk * BK, g.m, g.k, vec_a, tid);
          load_tile<B_MN>(sb, B, g.ldb, n0, nk * BK, g.n, g.k, vec_b, tid);
        }
        cp_async_commit();
      }
      const double *sa = smem + (kt % STAGES) * STAGE_DOUBLES + a_off;
      const double *sb = smem + (kt % STAGES) * STAGE_DOUBLES + OPERAND_DOUBLES + b_off;
#pragma unroll
      for (int k4 = 0; k4 < BK / 4; k4++) {
        double af[8], bf[4];
#pragma unroll
        for (int i = 0; i < 8; i++) af[i] = sa[k4 * A_K4 + i * A_MT];
#pragma unroll
        for (int j = 0; j < 4; j++) bf[j] = sb[k4 * B_K4 + j * B_NT];
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], bf[j], af[i]);
      }
    }
    cp_async_wait<0>();
    __syncthreads();   /* all warps done with the last stages before the next tile's prologue */

    /* epilogue: lane owns C[m .. m+1][n], m = m0+wm+8i+2*fk, n = n0+wn+8j+fi */
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int64_t n = n0 + wn + 8 * j + fi;
      if (n >= g.n) continue;
#pragma unroll
      for (int i = 0; i < 8; i++) {
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

template <bool A_MN, bool B_MN>
cudaError_t launch_variant(const DeviceGemm &g, cudaStream_t stream, int vec_a, int vec_b, int vec_c) {
  static bool configured = false;   /* per instantiation; benign race: same value every time */
  auto kern = dgemm_dmma_kernel<A_MN, B_MN>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  int64_t tiles = ((g.m + BM - 1) / BM) * ((g.n + BN - 1) / BN);
  int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  kern<<<grid, THREADS, SMEM_BYTES, stream>>>(g, vec_a, vec_b, vec_c);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_dgemm_dmma(const DeviceGemm &g, cudaStream_t stream) {
  if (g.dtype != B200_D) return cudaErrorNotSupported;
  if (((uintptr_t)g.a | (uintptr_t)g.b | (uintptr_t)g.c) & 7) return cudaErrorNotSupported;
  const bool a_mn = !(g.transa & 1), b_mn = (g.transb & 1);
  const int vec_a = (((uintptr_t)g.a & 15) == 0) && (g.lda % 2 == 0);
  const int vec_b = (((uintptr_t)g.b & 15) == 0) && (g.ldb % 2 == 0);
  const int vec_c = (((uintptr_t)g.c & 15) == 0) && (g.ldc % 2 == 0);
  cudaError_t e;
  if (a_mn && b_mn) e = launch_variant<true, true>(g, stream, vec_a, vec_b, vec_c);
  else if (a_mn && !b_mn) e = launch_variant<true, false>(g, stream, vec_a, vec_b, vec_c);
  else if (!a_mn && b_mn) e = launch_variant<false, true>(g, stream, vec_a, vec_b, vec_c);
  else e = launch_variant<false, false>(g, stream, vec_a, vec_b, vec_c);
  if (e == cudaSuccess) count_launch("dgemm_dmma_128x128x16");
  return e;
}

}  // namespace b200
