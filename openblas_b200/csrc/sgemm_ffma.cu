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

namespace b200 {
namespace {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int STAGES = 3;
constexpr int THREADS = 256;
constexpr int LDS = BM + 4;                          /* 132 floats = 528 B (16-byte multiple) */
constexpr int OPERAND_FLOATS = BK * LDS;
constexpr int STAGE_FLOATS = 2 * OPERAND_FLOATS;
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_FLOATS * sizeof(float);   /* 50688 B */

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

/* mn-contiguous operand: element (mn,k) at g[mn + k*ld] -> S[k][mn] by cp.async */
__device__ __forceinline__ void load_mn_async(float *s, const float *__restrict__ g, int64_t ld, int64_t mn0,
                                              int64_t k0, int64_t mn_end, int64_t k_end, bool vec, int tid) {
  if (vec) {
#pragma unroll
    for (int i = 0; i < (BM / 4) * BK / THREADS; i++) {   /* 2 */
      int idx = tid + i * THREADS;
      int k = idx / (BM / 4), mn = (idx % (BM / 4)) * 4;
      int64_t gk = k0 + k, gmn = mn0 + mn;
      int64_t left = (gk < k_end) ? (mn_end - gmn) : 0;
      int bytes = left >= 4 ? 16 : (left > 0 ? (int)left * 4 : 0);
      const float *src = bytes ? g + gmn + gk * ld : g;
      cp_async16(s + k * LDS + mn, src, bytes);
    }
  } else {
#pragma unroll
    for (int i = 0; i < BM * BK / THREADS; i++) {         /* 8 */
      int idx = tid + i * THREADS;
      int k = idx / BM, mn = idx % BM;
      int64_t gk = k0 + k, gmn = mn0 + mn;
      int bytes = (gk < k_end && gmn < mn_end) ? 4 : 0;
      const float *src = bytes ? g + gmn + gk * ld : g;
      cp_async4(s + k * LDS + mn, src, bytes);
    }
  }
}

/* k-contiguous operand: element (mn,k) at g[k + mn*ld].  Each thread owns two (mn, 4 k) quads. */
__device__ __forceinline__ void fetch_k(float (&r)[8], const float *__restrict__ g, int64_t ld, int64_t mn0,
                                        int64_t k0, int64_t mn_end, int64_t k_end, bool vec, int tid) {
#pragma unroll
  for (int i = 0; i < 2; i++) {
    int idx = tid + i * THREADS;
    int kq = (idx % 4) * 4, mn = idx / 4;
    int64_t gk = k0 + kq, gmn = mn0 + mn;
    if (gmn < mn_end && gk + 3 < k_end && vec) {
      float4 v = *reinterpret_cast<const float4 *>(g + gk + gmn * ld);
      r[4 * i + 0] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; e++)
        r[4 * i + e] = (gmn < mn_end && gk + e < k_end) ? g[gk + e + gmn * ld] : 0.f;
    }
  }
}
__device__ __forceinline__ void store_k(float *s, const float (&r)[8], int tid) {
#pragma unroll
  for (int i = 0; i < 2; i++) {
    int idx = tid + i * THREADS;
    int kq = (idx % 4) * 4, mn = idx / 4;
#pragma unroll
    for (int e = 0; e < 4; e++) s[(kq + e) * LDS + mn] = r[4 * i + e];
  }
}

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(THREADS, 2)
sgemm_ffma_kernel(DeviceGemm g, int vec_a, int vec_b, int vec_c) {
  extern __shared__ __align__(16) float fsmem[];
  const float *__restrict__ A = (const float *)g.a;
  const float *__restrict__ B = (const float *)g.b;
  float *__restrict__ C = (float *)g.c;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  /* warp = 8 (m) x 4 (n) threads; warps 2 (m) x 4 (n): thread coordinates in a 16 x 16 grid */
  const int tm = (warp & 1) * 8 + (lane & 7);
  const int tn = (warp >> 1) * 4 + (lane >> 3);

  const int64_t tiles_m = (g.m + BM - 1) / BM, tiles_n = (g.n + BN - 1) / BN;
  const int64_t tiles = tiles_m * tiles_n;
  const int64_t ktiles = (g.k + BK - 1) / BK;
  const float alpha = (float)g.alpha_re, beta = (float)g.beta_re;
  const bool use_beta = beta != 0.f;

  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    int64_t bm, bn;
    banded_tile_coords<16>(t, tiles_m, tiles_n, bm, bn);
    const int64_t m0 = bm * BM, n0 = bn * BN;

    /* acc[p][j]: rows (2p, 2p+1) of the thread's 8 rows, column j of its 8 columns */
    u64 acc[4][8];
#pragma unroll
    for (int p = 0; p < 4; p++)
#pragma unroll
      for (int j = 0; j < 8; j++) acc[p][j] = 0ull;

    float ra[8], rb[8];
#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
      if (s < ktiles) {
        float *sa = fsmem + s * STAGE_FLOATS, *sb = sa + OPERAND_FLOATS;
        if (A_MN) load_mn_async(sa, A, g.lda, m0, (int64_t)s * BK, g.m, g.k, vec_a, tid);
        else { fetch_k(ra, A, g.lda, m0, (int64_t)s * BK, g.m, g.k, vec_a, tid); store_k(sa, ra, tid); }
        if (B_MN) load_mn_async(sb, B, g.ldb, n0, (int64_t)s * BK, g.n, g.k, vec_b, tid);
        else { fetch_k(rb, B, g.ldb, n0, (int64_t)s * BK, g.n, g.k, vec_b, tid); store_k(sb, rb, tid); }
      }
      cp_async_commit();
    }

    for (int64_t kt = 0; kt < ktiles; kt++) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      const int64_t nk = kt + STAGES - 1;
      const bool refill = nk < ktiles;
      float *na = fsmem + (nk % STAGES) * STAGE_FLOATS, *nb = na + OPERAND_FLOATS;
      if (refill) {
        if (A_MN) load_mn_async(na, A, g.lda, m0, nk * BK, g.m, g.k, vec_a, tid);
        else fetch_k(ra, A, g.lda, m0, nk * BK, g.m, g.k, vec_a, tid);
        if (B_MN) load_mn_async(nb, B, g.ldb, n0, nk * BK, g.n, g.k, vec_b, tid);
        else fetch_k(rb, B, g.ldb, n0, nk * BK, g.n, g.k, vec_b, tid);
      }
      cp_async_commit();

      const float *sa = fsmem + (kt % STAGES) * STAGE_FLOATS + tm * 4;
      const float *sb = fsmem + (kt % STAGES) * STAGE_FLOATS + OPERAND_FLOATS + tn * 4;
#pragma unroll
      for (int k = 0; k < BK; k++) {
        /* rows tm*4..+3 and 64+tm*4..+3 as two 64-bit pairs each; columns likewise as floats */
        ulonglong2 a_lo = *reinterpret_cast<const ulonglong2 *>(sa + k * LDS);
        ulonglong2 a_hi = *reinterpret_cast<const ulonglong2 *>(sa + k * LDS + 64);
        float4 b_lo = *reinterpret_cast<const float4 *>(sb + k * LDS);
        float4 b_hi = *reinterpret_cast<const float4 *>(sb + k * LDS + 64);
        u64 ap[4] = {a_lo.x, a_lo.y, a_hi.x, a_hi.y};
        float bv[8] = {b_lo.x, b_lo.y, b_lo.z, b_lo.w, b_hi.x, b_hi.y, b_hi.z, b_hi.w};
#pragma unroll
        for (int j = 0; j < 8; j++) {
          u64 bb = pack2(bv[j], bv[j]);
#pragma unroll
          for (int p = 0; p < 4; p++) ffma2(acc[p][j], ap[p], bb);
        }
      }
      if (refill) {
        if (!A_MN) store_k(na, ra, tid);
        if (!B_MN) store_k(nb, rb, tid);
      }
    }
    cp_async_wait<0>();
    __syncthreads();

    /* epilogue: column j -> n, row pair p -> m */
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int64_t n = n0 + (j < 4 ? tn * 4 + j : 64 + tn * 4 + (j - 4));
      if (n >= g.n) continue;
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int64_t m = m0 + h * 64 + tm * 4;
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
cudaError_t launch_variant(const DeviceGemm &g, cudaStream_t stream, int vec_a, int vec_b, int vec_c) {
  static bool configured = false;
  auto kern = sgemm_ffma_kernel<A_MN, B_MN>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  int64_t tiles = ((g.m + BM - 1) / BM) * ((g.n + BN - 1) / BN);
  int64_t cap = (int64_t)sm_count() * 2;
  int grid = (int)(tiles < cap ? tiles : cap);
  kern<<<grid, THREADS, SMEM_BYTES, stream>>>(g, vec_a, vec_b, vec_c);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_sgemm_ffma(const DeviceGemm &g, cudaStream_t stream) {
  if (g.dtype != B200_S) return cudaErrorNotSupported;
  if (((uintptr_t)g.a | (uintptr_t)g.b | (uintptr_t)g.c) & 3) return cudaErrorNotSupported;
  const bool a_mn = !(g.transa & 1), b_mn = (g.transb & 1);
  const int vec_a = (((uintptr_t)g.a & 15) == 0) && (g.lda % 4 == 0);
  const int vec_b = (((uintptr_t)g.b & 15) == 0) && (g.ldb % 4 == 0);
  const int vec_c = (((uintptr_t)g.c & 15) == 0) && (g.ldc % 4 == 0);
  cudaError_t e;
  if (a_mn && b_mn) e = launch_variant<true, true>(g, stream, vec_a, vec_b, vec_c);
  else if (a_mn && !b_mn) e = launch_variant<true, false>(g, stream, vec_a, vec_b, vec_c);
  else if (!a_mn && b_mn) e = launch_variant<false, true>(g, stream, vec_a, vec_b, vec_c);
  else e = launch_variant<false, false>(g, stream, vec_a, vec_b, vec_c);
  if (e == cudaSuccess) count_launch("sgemm_ffma2_128x128x16");
  return e;
}

}  // namespace b200
