/*
 * sbgemm_tcgen05.cu -- SBGEMM (bf16 x bf16 -> fp32) on the 5th-generation tensor cores:
 * TMA (cp.async.bulk.tensor) -> 128B-swizzled shared-memory ring -> tcgen05.mma kind::f16 with
 * the fp32 accumulator in TENSOR MEMORY -> tcgen05.ld epilogue applying alpha/beta -> coalesced
 * stores into column-major C.  All hand-written PTX; no CUTLASS, no cuBLAS.
 *
 * What it replaces in the reference:
 *   kernel/x86_64/sbgemm_kernel_16x16_spr.c (+_tmpl.c)   AMX TDPBF16PS 16x16x32 tiles
 *   sbgemm_{n,t}copy_16_cooperlake.c, sbgemm_o{n,t}copy_16_spr.c   packing copies
 *   driver/level3/level3.c:288-406                        GEMM_R/Q/P loop nest
 *   sgemm_beta (KERNEL.SAPPHIRERAPIDS:14)                  separate beta pass
 * Transposition is absorbed by the operand DESCRIPTORS, not by a copy: an operand stored with k
 * contiguous (A transposed / B not transposed) is fed as a K-major UMMA operand, one stored with
 * m or n contiguous (A not transposed / B transposed) as an MN-major operand; the TMA tensor map
 * just names the contiguous dimension first.  Four kernels (A major x B major), same speed.
 *
 * Two variants: single-CTA tiles (below) and the default CTA-PAIR variant (cta_group::2, 256 x 256
 * tiles, further down), which cuts L2 -> SM operand traffic by a third and measured +9 % (1.47 vs 1.35
 * PFLOP/s at 8192^3).
 *
 * CTA = 192 threads, one per SM, persistent over 128 x 256 C tiles:
 *   warp 0      TMA producer (one lane): fills a 4-stage ring of {A 128x64, B 256x64} bf16 tiles
 *   warp 1      TMEM allocator + MMA issuer (one lane): 4 x tcgen05.mma M128 N256 K16 per stage,
 *               tcgen05.commit frees the stage / publishes the accumulator
 *   warps 2..5  epilogue: tcgen05.ld 32 lanes x 32 columns at a time, alpha/beta in fp32, store
 * Two 256-column accumulators (all 512 TMEM columns) let the epilogue of tile i overlap the
 * main loop of tile i+1.  mbarrier pipelines: full/empty per stage, tmem_full/tmem_empty per
 * accumulator.  Deterministic: one CTA owns a C tile for the whole k range.
 *
 * Descriptor encodings follow the PTX ISA "tcgen05 matrix descriptor" / "instruction
 * descriptor" tables (bit positions are spelled out at make_smem_desc / make_idesc below).
 */
#include <cuda.h>
#include "gemm_common.cuh"
#include <cstdlib>

namespace b200 {
namespace {

constexpr int BLOCK_M = 128, BLOCK_N = 256, BLOCK_K = 64;      /* bf16: 64 k = one 128-byte swizzle row */
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int THREADS = 192;
constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 2;             /* 16 KB */
constexpr uint32_t B_BYTES = BLOCK_N * BLOCK_K * 2;             /* 32 KB */
constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
constexpr uint32_t ATOM_BYTES = 64 * BLOCK_K * 2;               /* one 64(mn) x 64(k) MN-major box: 8 KB */
constexpr uint32_t TMEM_COLS = 512;
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

/* ------------------------------------------------------------------------------- PTX */
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

/* Shared-memory matrix descriptor (64 bit):
 *   [0,14)  start address >> 4        [16,30) leading-dimension byte offset >> 4
 *   [32,46) stride byte offset >> 4   [46,48) version = 1 (Blackwell)
 *   [49,52) base offset = 0           [61,64) layout: 2 = SWIZZLE_128B
 * K-major, 128B swizzle: rows of 64 bf16 (128 B), 8-row atoms of 1024 B -> SBO = 1024, LBO unused.
 * MN-major, 128B swizzle: atoms of 64 (mn) x 8 (k), SBO = distance between k groups of 8 (1024 B),
 *   LBO = distance between 64-wide mn atoms (one TMA box = 8192 B). */
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

/* Instruction descriptor (32 bit) for kind::f16:
 *   [4,6) D format: 1 = f32    [7,10) A format: 1 = bf16   [10,13) B format: 1 = bf16
 *   [15] A major: 0 = K, 1 = MN    [16] B major    [17,23) N >> 3    [24,29) M >> 4 */
__host__ __device__ constexpr uint32_t make_idesc(bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}

__device__ __forceinline__ void tile_coords(int t, int tiles_m, int tiles_n, int &bm, int &bn) {
  constexpr int BAND = 8;
  int per_band = BAND * tiles_n;
  int band = t / per_band, r = t % per_band;
  int rows = tiles_m - band * BAND;
  if (rows > BAND) rows = BAND;
  bm = band * BAND + r % rows;
  bn = r / rows;
}

/* SBGEMMT (DeviceGemm::tri = 1 lower / 2 upper, M == N): the persistent loops walk only the tiles that hold an element
 * of the triangle -- the closed-form enumerations of gemm_common.cuh, WIDE for the 128 x 256 tiles of the single-CTA
 * kernel, square for the 256 x 256 tiles of the CTA pairs -- and the epilogue masks the stores of the tiles the
 * diagonal crosses.  Half the tiles, half the time; the untouched triangle of C is neither read nor written. */
template <bool WIDE>
__host__ __device__ __forceinline__ int tile_count_for(int tri, int tiles_m, int tiles_n) {
  if (!tri) return tiles_m * tiles_n;
  return (int)(WIDE ? tri12_tile_count(tri, tiles_m, tiles_n) : tri_tile_count(tiles_m));
}
template <bool WIDE>
__device__ __forceinline__ void tile_coords_for(int t, int tri, int tiles_m, int tiles_n, int &bm, int &bn) {
  if (!tri) { tile_coords(t, tiles_m, tiles_n, bm, bn); return; }
  int64_t r, c;
  if (WIDE) tri12_tile_coords(t, tri, r, c); else tri_tile_coords(t, tri, r, c);
  bm = (int)r; bn = (int)c;
}

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(THREADS, 1)
sbgemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                      float *__restrict__ C, int64_t ldc, int M, int N, int K, float alpha, float beta, int tri) {
  extern __shared__ uint8_t raw_smem[];
  const uint32_t base = (smem_u32(raw_smem) + 1023u) & ~1023u;     /* SWIZZLE_128B needs 1024-byte alignment */
  const uint32_t bars = base + STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 4);
  volatile uint32_t *tmem_slot_ptr = (volatile uint32_t *)(raw_smem + (tmem_slot - smem_u32(raw_smem)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (M + BLOCK_M - 1) / BLOCK_M, tiles_n = (N + BLOCK_N - 1) / BLOCK_N;
  const int tiles = tile_count_for<true>(tri, tiles_m, tiles_n);
  const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int s = 0; s < STAGES; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; a++) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    /* ===================================================== TMA producer */
    int stage = 0; uint32_t phase = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
      int bm, bn;
      tile_coords_for<true>(t, tri, tiles_m, tiles_n, bm, bn);
      const int m0 = bm * BLOCK_M, n0 = bn * BLOCK_N;
      for (int kb = 0; kb < num_kb; kb++) {
        if (lane == 0) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          mbar_expect_tx(full_bar(stage), STAGE_BYTES);
          const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + A_BYTES;
          const int k0 = kb * BLOCK_K;
          if (A_MN) {
#pragma unroll
            for (int i = 0; i < BLOCK_M / 64; i++) tma_load_2d(sa + i * ATOM_BYTES, &map_a, full_bar(stage), m0 + i * 64, k0);
          } else {
            tma_load_2d(sa, &map_a, full_bar(stage), k0, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int i = 0; i < BLOCK_N / 64; i++) tma_load_2d(sb + i * ATOM_BYTES, &map_b, full_bar(stage), n0 + i * 64, k0);
          } else {
            tma_load_2d(sb, &map_b, full_bar(stage), k0, n0);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    /* ===================================================== MMA issuer */
    constexpr uint32_t idesc = make_idesc(A_MN, B_MN);
    int stage = 0; uint32_t phase = 0;
    int local = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, local++) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      if (lane == 0) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);     /* epilogue has drained this accumulator */
        tc_fence_after();
      }
      __syncwarp();
      for (int kb = 0; kb < num_kb; kb++) {
        if (lane == 0) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
          for (int k4 = 0; k4 < BLOCK_K / UMMA_K; k4++) {
            /* K-major: step 16 k = 32 bytes inside the swizzled 128-byte row.
             * MN-major: step 16 k = 16 rows of 128 bytes. */
            const uint64_t adesc = A_MN ? make_smem_desc(sa + k4 * (UMMA_K * 128), ATOM_BYTES, 1024)
                                        : make_smem_desc(sa + k4 * (UMMA_K * 2), 16, 1024);
            const uint64_t bdesc = B_MN ? make_smem_desc(sb + k4 * (UMMA_K * 128), ATOM_BYTES, 1024)
                                        : make_smem_desc(sb + k4 * (UMMA_K * 2), 16, 1024);
            tc_mma_bf16(tmem_base + acc * BLOCK_N, adesc, bdesc, idesc, (kb | k4) != 0);
          }
          tc_commit(empty_bar(stage));                 /* stage reusable once these MMAs retire */
          if (kb == num_kb - 1) tc_commit(tfull_bar(acc));
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    /* ===================================================== epilogue (warps 2..5) */
    const int quarter = warp & 3;                      /* TMEM lanes 32*quarter .. +31 */
    const bool use_beta = beta != 0.f;
    int local = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, local++) {
      int bm, bn;
      tile_coords_for<true>(t, tri, tiles_m, tiles_n, bm, bn);
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const int64_t m = (int64_t)bm * BLOCK_M + quarter * 32 + lane;
      const int64_t n0 = (int64_t)bn * BLOCK_N;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * BLOCK_N + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BLOCK_N; c += 32) {
        uint32_t r[32];
        tc_ld_32x32(taddr + c, r);
        tc_wait_ld();
        if (m < M) {
          float *p = C + m + (n0 + c) * ldc;
#pragma unroll
          for (int j = 0; j < 32; j++) {
            if (n0 + c + j < N && tri_keep(tri, m, n0 + c + j)) {
              float v = alpha * __uint_as_float(r[j]);
              if (use_beta) v = fmaf(beta, p[(int64_t)j * ldc], v);
              p[(int64_t)j * ldc] = v;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

/* ==========================================================================================
 * 2-CTA variant (cta_group::2): a cluster of two CTAs on one TPC computes a 256 x 256 C tile with
 * ONE tcgen05.mma M256 N256 K16 per k16 step, issued by the leader CTA.  Each CTA stages only its
 * own 128 rows of A and HALF of the B tile (128 of the 256 columns); the tensor core reads both
 * halves from the two CTAs' shared memories, so L2 -> SM traffic per flop drops by a third
 * (32 KB instead of 48 KB per CTA per 64-wide k block; ncu on the 1-CTA kernel: tensor pipe only
 * 78 % active, the MMA issuer waiting on TMA).  Both CTAs' TMA loads report to the leader's full[]
 * barrier; tcgen05.commit multicasts "stage free" / "accumulator ready" to both CTAs; the epilogue
 * warps of both CTAs release the accumulator on the leader's barrier (remote mbarrier arrive).
 * ========================================================================================== */
namespace two {
constexpr int STAGES = 6;
constexpr uint32_t A_BYTES = 128 * BLOCK_K * 2;          /* this CTA's 128 rows of A */
constexpr uint32_t B_BYTES = 128 * BLOCK_K * 2;          /* this CTA's half of the 256-column B tile */
constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;      /* 32 KB */
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 + 256;
constexpr int TILE_M = 256, TILE_N = 256;
}  // namespace two

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
/* arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster */
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t"
      ".reg .b32 r;\n\t"
      "mapa.shared::cluster.u32 r, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [r];\n\t"
      "}" ::"r"(bar), "r"(rank) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
  /* peer bit cleared: the transaction bytes are counted on CTA 0's barrier */
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_commit_2sm_mcast(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_2sm(bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(two::TILE_N >> 3) << 17) | ((uint32_t)(two::TILE_M >> 4) << 24);
}

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(THREADS, 1)
sbgemm_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                           float *__restrict__ C, int64_t ldc, int M, int N, int K, float alpha, float beta, int tri) {
  constexpr int STAGES = two::STAGES;
  constexpr uint32_t A_BYTES = two::A_BYTES, STAGE_BYTES = two::STAGE_BYTES;
  extern __shared__ uint8_t raw_smem[];
  const uint32_t base = (smem_u32(raw_smem) + 1023u) & ~1023u;
  const uint32_t bars = base + STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 4);
  volatile uint32_t *tmem_slot_ptr = (volatile uint32_t *)(raw_smem + (tmem_slot - smem_u32(raw_smem)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int tiles_m = (M + two::TILE_M - 1) / two::TILE_M, tiles_n = (N + two::TILE_N - 1) / two::TILE_N;
  const int tiles = tile_count_for<false>(tri, tiles_m, tiles_n);
  const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int s = 0; s < STAGES; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; a++) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    /* ===================================================== TMA producer (both CTAs) */
    int stage = 0; uint32_t phase = 0;
    for (int t = cluster_id; t < tiles; t += num_clusters) {
      int bm, bn;
      tile_coords_for<false>(t, tri, tiles_m, tiles_n, bm, bn);
      const int m0 = bm * two::TILE_M + (int)rank * 128, n0 = bn * two::TILE_N + (int)rank * 128;
      for (int kb = 0; kb < num_kb; kb++) {
        if (lane == 0) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          if (leader) mbar_expect_tx(full_bar(stage), 2 * STAGE_BYTES);     /* both CTAs' bytes */
          const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + A_BYTES;
          const int k0 = kb * BLOCK_K;
          if (A_MN) {
#pragma unroll
            for (int i = 0; i < 2; i++) tma_load_2d_2sm(sa + i * ATOM_BYTES, &map_a, full_bar(stage), m0 + i * 64, k0);
          } else {
            tma_load_2d_2sm(sa, &map_a, full_bar(stage), k0, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int i = 0; i < 2; i++) tma_load_2d_2sm(sb + i * ATOM_BYTES, &map_b, full_bar(stage), n0 + i * 64, k0);
          } else {
            tma_load_2d_2sm(sb, &map_b, full_bar(stage), k0, n0);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    /* ===================================================== MMA issuer (leader CTA only) */
    if (leader) {
      constexpr uint32_t idesc = make_idesc_2sm(A_MN, B_MN);
      int stage = 0; uint32_t phase = 0;
      int local = 0;
      for (int t = cluster_id; t < tiles; t += num_clusters, local++) {
        const int acc = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1;
        if (lane == 0) {
          mbar_wait(tempty_bar(acc), acc_phase ^ 1);
          tc_fence_after();
        }
        __syncwarp();
        for (int kb = 0; kb < num_kb; kb++) {
          if (lane == 0) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
            for (int k4 = 0; k4 < BLOCK_K / UMMA_K; k4++) {
              const uint64_t adesc = A_MN ? make_smem_desc(sa + k4 * (UMMA_K * 128), ATOM_BYTES, 1024)
                                          : make_smem_desc(sa + k4 * (UMMA_K * 2), 16, 1024);
              const uint64_t bdesc = B_MN ? make_smem_desc(sb + k4 * (UMMA_K * 128), ATOM_BYTES, 1024)
                                          : make_smem_desc(sb + k4 * (UMMA_K * 2), 16, 1024);
              tc_mma_bf16_2sm(tmem_base + acc * two::TILE_N, adesc, bdesc, idesc, (kb | k4) != 0);
            }
            tc_commit_2sm_mcast(empty_bar(stage), 0x3);
            if (kb == num_kb - 1) tc_commit_2sm_mcast(tfull_bar(acc), 0x3);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    /* ===================================================== epilogue (warps 2..5 of both CTAs) */
    const int quarter = warp & 3;
    const bool use_beta = beta != 0.f;
    int local = 0;
    for (int t = cluster_id; t < tiles; t += num_clusters, local++) {
      int bm, bn;
      tile_coords_for<false>(t, tri, tiles_m, tiles_n, bm, bn);
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const int64_t m = (int64_t)bm * two::TILE_M + rank * 128 + quarter * 32 + lane;
      const int64_t n0 = (int64_t)bn * two::TILE_N;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * two::TILE_N + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < two::TILE_N; c += 32) {
        uint32_t r[32];
        tc_ld_32x32(taddr + c, r);
        tc_wait_ld();
        if (m < M) {
          float *p = C + m + (n0 + c) * ldc;
#pragma unroll
          for (int j = 0; j < 32; j++) {
            if (n0 + c + j < N && tri_keep(tri, m, n0 + c + j)) {
              float v = alpha * __uint_as_float(r[j]);
              if (use_beta) v = fmaf(beta, p[(int64_t)j * ldc], v);
              p[(int64_t)j * ldc] = v;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(tempty_bar(acc), 0);      /* the leader's MMA warp waits on its own barrier */
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

/* ------------------------------------------------------------------------ host side */
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

/* 2-D bf16 tensor map: dim0 = the contiguous dimension (extent0), dim1 strided by ld elements. */
bool make_map(CUtensorMap *map, const void *ptr, uint64_t extent0, uint64_t extent1, uint64_t ld, uint32_t box0,
              uint32_t box1) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {extent0, extent1};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <bool A_MN, bool B_MN>
cudaError_t launch_variant(const DeviceGemm &g, cudaStream_t stream) {
  static bool configured = false;
  auto kern = sbgemm_tcgen05_kernel<A_MN, B_MN>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  CUtensorMap map_a, map_b;
  /* stored A is (A_MN ? m x k : k x m), stored B is (B_MN ? n x k : k x n), both column-major */
  bool ok = A_MN ? make_map(&map_a, g.a, (uint64_t)g.m, (uint64_t)g.k, (uint64_t)g.lda, 64, BLOCK_K)
                 : make_map(&map_a, g.a, (uint64_t)g.k, (uint64_t)g.m, (uint64_t)g.lda, BLOCK_K, BLOCK_M);
  ok = ok && (B_MN ? make_map(&map_b, g.b, (uint64_t)g.n, (uint64_t)g.k, (uint64_t)g.ldb, 64, BLOCK_K)
                   : make_map(&map_b, g.b, (uint64_t)g.k, (uint64_t)g.n, (uint64_t)g.ldb, BLOCK_K, BLOCK_N));
  if (!ok) return cudaErrorNotSupported;
  int tiles = tile_count_for<true>(g.tri, (int)((g.m + BLOCK_M - 1) / BLOCK_M), (int)((g.n + BLOCK_N - 1) / BLOCK_N));
  int grid = tiles < sm_count() ? tiles : sm_count();
  kern<<<grid, THREADS, SMEM_BYTES, stream>>>(map_a, map_b, (float *)g.c, g.ldc, (int)g.m, (int)g.n, (int)g.k,
                                              (float)g.alpha_re, (float)g.beta_re, g.tri);
  return cudaGetLastError();
}

template <bool A_MN, bool B_MN>
cudaError_t launch_2cta_variant(const DeviceGemm &g, cudaStream_t stream) {
  static bool configured = false;
  auto kern = sbgemm_tcgen05_2cta_kernel<A_MN, B_MN>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)two::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  CUtensorMap map_a, map_b;
  bool ok = A_MN ? make_map(&map_a, g.a, (uint64_t)g.m, (uint64_t)g.k, (uint64_t)g.lda, 64, BLOCK_K)
                 : make_map(&map_a, g.a, (uint64_t)g.k, (uint64_t)g.m, (uint64_t)g.lda, BLOCK_K, 128);
  ok = ok && (B_MN ? make_map(&map_b, g.b, (uint64_t)g.n, (uint64_t)g.k, (uint64_t)g.ldb, 64, BLOCK_K)
                   : make_map(&map_b, g.b, (uint64_t)g.k, (uint64_t)g.n, (uint64_t)g.ldb, BLOCK_K, 128));
  if (!ok) return cudaErrorNotSupported;
  int tiles = tile_count_for<false>(g.tri, (int)((g.m + two::TILE_M - 1) / two::TILE_M), (int)((g.n + two::TILE_N - 1) / two::TILE_N));
  int clusters = sm_count() / 2;
  if (tiles < clusters) clusters = tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * clusters));
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = two::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, map_a, map_b, (float *)g.c, g.ldc, (int)g.m, (int)g.n, (int)g.k,
                            (float)g.alpha_re, (float)g.beta_re, (int)g.tri);
}

}  // namespace

/* column-major bf16 block copy into a 16-byte aligned, pitch % 8 == 0 layout (operands TMA cannot address in place) */
__global__ void __launch_bounds__(256) repack_bf16_kernel(int64_t rows, int64_t cols, const uint16_t *__restrict__ src, int64_t ld_src,
                                                          uint16_t *__restrict__ dst, int64_t ld_dst) {
  const int64_t tiles_r = (rows + 255) / 256;
  for (int64_t t = blockIdx.x; t < tiles_r * cols; t += gridDim.x) {
    const int64_t r = (t % tiles_r) * 256 + threadIdx.x, c = t / tiles_r;
    if (r < rows) dst[r + c * ld_dst] = src[r + c * ld_src];
  }
}

cudaError_t launch_sbgemm_tcgen05(const DeviceGemm &g_in, cudaStream_t stream) {
  DeviceGemm g = g_in;
  if (g.dtype != B200_SB || (g.tri && g.m != g.n)) return cudaErrorNotSupported;
  if ((uintptr_t)g.c & 3) return cudaErrorNotSupported;
  if (g.m < 1 || g.n < 1 || g.k < 1) return cudaErrorNotSupported;
  if (g.m > (1 << 30) || g.n > (1 << 30) || g.k > (1 << 30)) return cudaErrorNotSupported;
  if (((g.m + BLOCK_M - 1) / BLOCK_M) * ((g.n + BLOCK_N - 1) / BLOCK_N) > (1ll << 30)) return cudaErrorNotSupported;
  const bool a_mn = !(g.transa & 1), b_mn = (g.transb & 1);
  /* TMA needs 16-byte aligned bases and pitches that are multiples of 16 bytes.  An operand that is not (odd leading
   * dimension, a sub-matrix starting at an odd column) is copied ONCE into an aligned stream-ordered temporary by an
   * HBM-bound pass (2 bytes read + 2 written per element against 2 n or 2 m flops per element in the product) --
   * the reference's AMX kernel (sbgemm_kernel_16x16_spr_tmpl.c) repacks EVERY operand; here only misaligned ones.
   * Boxes may exceed the tensor (TMA zero-fills), so small and ragged extents need nothing special. */
  void *tmp[2] = {nullptr, nullptr};
  auto fix = [&](const void *&ptr, int64_t &ld, int64_t rows, int64_t cols, int which) -> cudaError_t {
    if ((((uintptr_t)ptr) & 15) == 0 && ld % 8 == 0) return cudaSuccess;
    const int64_t ld2 = (rows + 7) / 8 * 8;
    cudaError_t e = cudaMallocAsync(&tmp[which], (size_t)ld2 * (size_t)cols * 2, stream);
    if (e != cudaSuccess) return e;
    int64_t blocks = ((rows + 255) / 256) * cols;
    if (blocks > 8 * (int64_t)sm_count()) blocks = 8 * (int64_t)sm_count();
    repack_bf16_kernel<<<(int)blocks, 256, 0, stream>>>(rows, cols, (const uint16_t *)ptr, ld, (uint16_t *)tmp[which], ld2);
    count_launch("repack_bf16");
    ptr = tmp[which]; ld = ld2;
    return cudaGetLastError();
  };
  cudaError_t e = fix(g.a, g.lda, a_mn ? g.m : g.k, a_mn ? g.k : g.m, 0);
  if (e == cudaSuccess) e = fix(g.b, g.ldb, b_mn ? g.n : g.k, b_mn ? g.k : g.n, 1);
  if (e != cudaSuccess) { for (void *t : tmp) if (t) cudaFreeAsync(t, stream); return e; }
  static int cfg = -1;
  if (cfg < 0) { const char *ev = getenv("B200_SBGEMM_CFG"); cfg = ev ? atoi(ev) : 0; }   /* 2 = CTA pairs, 1 = single CTA, 0 = by shape */
  /* CTA pairs work on 256 x 256 tiles; with at most 128 rows half of every pair would multiply zeros: single CTAs
   * (128 x 256 tiles) take those */
  const int use = cfg ? cfg : (g.m <= 128 ? 1 : 2);
  if (use == 2) {
    if (a_mn && b_mn) e = launch_2cta_variant<true, true>(g, stream);
    else if (a_mn && !b_mn) e = launch_2cta_variant<true, false>(g, stream);
    else if (!a_mn && b_mn) e = launch_2cta_variant<false, true>(g, stream);
    else e = launch_2cta_variant<false, false>(g, stream);
    if (e == cudaSuccess) count_launch("sbgemm_tcgen05_2cta_256x256x64");
  } else {
    if (a_mn && b_mn) e = launch_variant<true, true>(g, stream);
    else if (a_mn && !b_mn) e = launch_variant<true, false>(g, stream);
    else if (!a_mn && b_mn) e = launch_variant<false, true>(g, stream);
    else e = launch_variant<false, false>(g, stream);
    if (e == cudaSuccess) count_launch("sbgemm_tcgen05_128x256x64");
  }
  for (void *t : tmp) if (t) cudaFreeAsync(t, stream);
  return e;
}

}  // namespace b200
