/* async_copy.cuh -- cp.async (LDGSTS) helpers shared by the cp.async-fed kernels. */
#pragma once
#include <cuda_runtime.h>
#include <cstring>

namespace b200 {

/* 16-byte copy global -> shared; bytes beyond src_bytes (0, 4, 8, 12 or 16) are zero-filled. */
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

/* DMMA.8x8x4: D(8x8) += A(8x4, row) * B(4x8, col); lane l holds A[l/4][l%4], B[l%4][l/4] and
 * D[l/4][2*(l%4) + {0,1}]. */
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

/* tile index -> (m-block, n-block): walk C in bands of BAND m-blocks so the tiles in flight
 * share few A row-panels and B column-panels in L2 */
template <int BAND>
__device__ __forceinline__ void banded_tile_coords(int64_t t, int64_t tiles_m, int64_t tiles_n, int64_t &bm,
                                                   int64_t &bn) {
  int64_t per_band = (int64_t)BAND * tiles_n;
  int64_t band = t / per_band, r = t % per_band;
  int64_t band_rows = tiles_m - band * BAND;
  if (band_rows > BAND) band_rows = BAND;
  bm = band * BAND + r % band_rows;
  bn = r / band_rows;
}


/* ---------------------------------------------------------------------------------------------
 * TileLoader: a thread's share of the cp.async traffic for one (ROWS mn) x (BK k) operand tile,
 * with ALL address arithmetic hoisted out of the k loop.  (First version recomputed 64-bit
 * addresses and bounds per 16-byte chunk per k tile: ncu showed ~2 integer instructions per DMMA
 * and IMADs competing with FFMA for the FMA pipe.)  Per k tile a thread now does, per chunk, one
 * pointer add and one LDGSTS; bounds in the mn direction are fixed per C tile, bounds in the k
 * direction only matter in the last k tile (issue_tail).
 *
 *   MN_CONTIG  element (mn,k) at g[mn + k*ld]  -> shared S[k][mn], row stride LD elements
 *   !MN_CONTIG element (mn,k) at g[k + mn*ld]  -> shared S[mn][k], row stride LD elements
 * ES = element bytes (4, 8, 16); chunks are 16 bytes (VE = 16/ES elements).  Requires 16-byte
 * aligned base and ld*ES % 16 == 0 (callers fall back to an element-wise loader otherwise). */
template <bool MN_CONTIG, int ES, int ROWS, int BK, int LD, int THREADS>
struct TileLoader {
  static constexpr int VE = 16 / ES;
  static constexpr int CPR = MN_CONTIG ? ROWS / VE : BK / VE;   /* chunks per contiguous run */
  static constexpr int STEP = THREADS / CPR;                    /* k rows (or mn rows) between a thread's chunks */
  static constexpr int N = (ROWS * BK / VE) / THREADS;          /* chunks per thread per tile */
  static constexpr int DST_STEP = STEP * LD * ES;               /* bytes between chunks in shared memory */
  static_assert(THREADS % CPR == 0 && (ROWS * BK / VE) % THREADS == 0, "tile does not divide among threads");

  const char *src;     /* this thread's chunk 0 of the CURRENT k tile */
  const char *safe;    /* any mapped address, used with src-size 0 */
  int64_t src_step;    /* bytes between a thread's consecutive chunks in global memory */
  int64_t k_adv;       /* bytes per k tile */
  uint32_t dst;        /* byte offset of chunk 0 inside the operand's shared tile */
  int edge;            /* MN_CONTIG: valid bytes of every chunk (mn edge). else: number of valid chunks */
  int krow;            /* MN_CONTIG: k row of chunk 0.  else: first k element of the chunks */

  __device__ __forceinline__ void init(const void *g, int64_t ld, int64_t mn0, int64_t mn_end, int tid) {
    safe = (const char *)g;
    const int run = tid % CPR, row = tid / CPR;
    if (MN_CONTIG) {
      const int64_t mn = mn0 + (int64_t)run * VE;
      int64_t left = mn_end - mn;
      edge = left >= VE ? 16 : (left > 0 ? (int)left * ES : 0);
      src = (const char *)g + (mn + (int64_t)row * ld) * ES;
      src_step = (int64_t)STEP * ld * ES;
      k_adv = (int64_t)BK * ld * ES;
      dst = (uint32_t)((row * LD + run * VE) * ES);
      krow = row;
    } else {
      const int64_t mn = mn0 + row;
      int64_t left = mn_end - mn;                                /* rows from mine to the edge */
      int64_t nv = left > 0 ? (left + STEP - 1) / STEP : 0;
      edge = nv > N ? N : (int)nv;
      src = (const char *)g + ((int64_t)run * VE + mn * ld) * ES;
      src_step = (int64_t)STEP * ld * ES;
      k_adv = (int64_t)BK * ES;
      dst = (uint32_t)((row * LD + run * VE) * ES);
      krow = run * VE;
    }
  }
  /* full k tile */
  __device__ __forceinline__ void issue(uint32_t smem_tile) const {
    const char *p = src;
#pragma unroll
    for (int i = 0; i < N; i++) {
      const int bytes = MN_CONTIG ? edge : (i < edge ? 16 : 0);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_tile + dst + i * DST_STEP),
                   "l"(bytes ? p : safe), "r"(bytes) : "memory");
      p += src_step;
    }
  }
  /* last, partial k tile: k_left (< BK) valid k */
  __device__ __forceinline__ void issue_tail(uint32_t smem_tile, int k_left) const {
    const char *p = src;
#pragma unroll
    for (int i = 0; i < N; i++) {
      int bytes;
      if (MN_CONTIG) bytes = (krow + i * STEP < k_left) ? edge : 0;
      else {
        int kl = k_left - krow;
        bytes = (i < edge && kl > 0) ? (kl >= VE ? 16 : kl * ES) : 0;
      }
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_tile + dst + i * DST_STEP),
                   "l"(bytes ? p : safe), "r"(bytes) : "memory");
      p += src_step;
    }
  }
  __device__ __forceinline__ void advance() { src += k_adv; }
  /* step back to k tile 0 of the same C tile is never needed: init() is called per C tile */
};

/* ---------------------------------------------------------------------------------------------
 * Lane -> (m position, n position) inside the 8 x 4 grid of thread tiles a warp of the FFMA kernels
 * covers.  Each lane reads its A fragment and its B fragment with LDS.128.  Measured with
 * tools/lds_probe.cu under ncu (profiles/r01_lds_wavefronts.txt): an LDS.128 is served one half-warp
 * at a time, identical addresses are merged only inside a QUAD of adjacent lanes, and a wavefront
 * delivers 128 B -- so the cost is (sum over quads of distinct bytes) / 128 B, at least 2.  The
 * obvious map (m = lane & 7, n = lane >> 3) gives every quad four distinct A chunks: 4 wavefronts
 * per A load, 2 per B load.  With 2 x 2 quads (below) a quad reads two A chunks and two B chunks,
 * a half-warp reads 8 contiguous A chunks (128 B, conflict free): 2 wavefronts for either load,
 * a third less shared-memory traffic for the same instructions. */
__device__ __forceinline__ void warp_tile_position(int lane, int &pm, int &pn) {
#ifdef B200_OLD_LANE_MAP
  pm = lane & 7; pn = lane >> 3;
#else
  pm = ((lane >> 2) & 3) * 2 + (lane & 1);
  pn = (lane >> 4) * 2 + ((lane >> 1) & 1);
#endif
}

/* ---------------------------------------------------------------------------------------------
 * KStager: the register-staged, TRANSPOSING path of the FFMA kernels for an operand stored
 * k-contiguous (element (mn,k) at g[k + mn*ld]) whose shared-memory image must be S[k][mn].
 * fetch() issues the thread's 16-byte global loads for the NEXT k tile before the FMA block,
 * store() scatters them into S[k][mn] after it.  Addresses are hoisted like in TileLoader.
 * T = float (VE = 4) or float2 (VE = 2). */
template <class T, int ROWS, int BK, int LDS, int THREADS>
struct KStager {
  static constexpr int ES = sizeof(T), VE = 16 / ES;
  static constexpr int CPR = BK / VE;
  static constexpr int STEP = THREADS / CPR;
  static constexpr int N = (ROWS * BK / VE) / THREADS;
  static_assert(THREADS % CPR == 0 && (ROWS * BK / VE) % THREADS == 0, "tile does not divide among threads");
  const char *src;
  int64_t src_step, ld_bytes;
  int nvalid, kq, row;
  bool vec;

  __device__ __forceinline__ void init(const void *g, int64_t ld, int64_t mn0, int64_t mn_end, bool aligned, int tid) {
    const int run = tid % CPR;
    row = tid / CPR;
    kq = run * VE;
    const int64_t mn = mn0 + row, left = mn_end - mn;
    int64_t nv = left > 0 ? (left + STEP - 1) / STEP : 0;
    nvalid = nv > N ? N : (int)nv;
    src = (const char *)g + ((int64_t)kq + mn * ld) * ES;
    src_step = (int64_t)STEP * ld * ES;
    vec = aligned;
  }
  __device__ __forceinline__ void fetch(T (&r)[N * VE], int k_left /* >= BK for a full tile */) const {
    const char *p = src;
#pragma unroll
    for (int i = 0; i < N; i++) {
      if (i < nvalid && vec && kq + VE <= k_left) {
        const float4 v = *reinterpret_cast<const float4 *>(p);
        const T *e = reinterpret_cast<const T *>(&v);
#pragma unroll
        for (int x = 0; x < VE; x++) r[i * VE + x] = e[x];
      } else {
#pragma unroll
        for (int x = 0; x < VE; x++) {
          T z; memset(&z, 0, sizeof z);
          r[i * VE + x] = (i < nvalid && kq + x < k_left) ? reinterpret_cast<const T *>(p)[x] : z;
        }
      }
      p += src_step;
    }
  }
  __device__ __forceinline__ void store(T *s, const T (&r)[N * VE]) const {
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
      for (int x = 0; x < VE; x++) s[(kq + x) * LDS + row + i * STEP] = r[i * VE + x];
  }
  __device__ __forceinline__ void advance() { src += (int64_t)BK * ES; }
};

/* ---------------------------------------------------------------------------------------------
 * mbarrier / bulk-copy helpers and the producer-warp tile fetch shared by the DMMA kernels. */
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
__device__ __forceinline__ void bulk_copy(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

/* Producer for one operand tile; every lane of the producer warp arrives once on `bar` per tile.
 * mn-contiguous storage, interior tile: whole 1 KB rows by bulk copy (one per lane).
 * k-contiguous storage: rows are only BK*8 = 256 bytes and the TMA engine retires roughly one bulk
 * copy per ~56 cycles per SM whatever its size (measured: TN with 256 such copies per stage ran at
 * 57 % of peak, NT with 64 copies of 1 KB at 98 %), so those tiles are fetched with 16-byte cp.async
 * from this warp instead, completion reported to the same mbarrier (arrive.noinc).
 * Edge tiles (partial in mn or in k) always use the cp.async form, whose src-size operand zero-fills
 * what lies outside the matrix. */
template <int ES, bool MN_CONTIG, int ROWS, int BK, int LD_MN, int LD_K>
__device__ __forceinline__ void produce_operand(uint32_t s_tile, const void *__restrict__ gv, int64_t ld, int64_t mn0,
                                                int64_t k0, int64_t mn_end, int64_t k_end, uint32_t bar, int lane) {
  constexpr int VE = 16 / ES;                         /* elements per 16-byte chunk (2 doubles, 1 complex double) */
  const char *g = (const char *)gv;
  const int64_t mn_left = mn_end - mn0, k_left = k_end - k0;
  const bool interior = mn_left >= ROWS && k_left >= BK;
  if (MN_CONTIG) {
    constexpr bool USE_BULK = ROWS * ES >= 1024;      /* shorter rows: the per-copy cost of the TMA engine would bound the tile */
    if (interior && !USE_BULK) {
      constexpr int CPR = ROWS / VE;                  /* 16-byte chunks per k row */
      static_assert(CPR % 32 == 0, "a k row must be a whole number of warp-wide copies");
      const char *src = g + (mn0 + lane * VE + k0 * ld) * ES;
      uint32_t dst = s_tile + (uint32_t)(lane * 16);
#pragma unroll 8
      for (int r = 0; r < BK; r++) {
#pragma unroll
        for (int j = 0; j < CPR / 32; j++)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + j * 512), "l"(src + j * 512) : "memory");
        src += ld * ES;
        dst += LD_MN * ES;
      }
      cp_async_mbar_arrive_noinc(bar);
      return;
    }
    if (interior) {
      if (lane == 0) mbar_expect_tx(bar, (uint32_t)ROWS * BK * ES); else mbar_arrive(bar);
      __syncwarp();
      for (int r = lane; r < BK; r += 32)
        bulk_copy(s_tile + (uint32_t)(r * LD_MN * ES), g + (mn0 + (k0 + r) * ld) * ES, ROWS * ES, bar);
      return;
    }
    constexpr int CPR = ROWS / VE;                    /* 16-byte chunks per k row */
    for (int c = lane; c < CPR * BK; c += 32) {
      const int k = c / CPR, mn = (c % CPR) * VE;
      const int64_t left = (k < k_left) ? mn_left - mn : 0;
      const int bytes = left >= VE ? 16 : (left > 0 ? (int)left * ES : 0);
      const char *src = bytes ? g + (mn0 + mn + (k0 + k) * ld) * ES : g;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s_tile + (uint32_t)((k * LD_MN + mn) * ES)), "l"(src), "r"(bytes) : "memory");
    }
    cp_async_mbar_arrive_noinc(bar);
  } else {
    constexpr int CPR = BK / VE;                      /* 16-byte chunks per row */
    constexpr int RSTEP = 32 / CPR;                   /* rows covered by one warp-wide copy */
    static_assert(32 % CPR == 0 && ROWS % RSTEP == 0, "k-contiguous tile does not divide over a warp");
    const int kc = (lane % CPR) * VE, r0 = lane / CPR;
    const char *src = g + (k0 + kc + (mn0 + r0) * ld) * ES;
    uint32_t dst = s_tile + (uint32_t)((r0 * LD_K + kc) * ES);
    const int64_t src_step = (int64_t)RSTEP * ld * ES;
    if (interior) {
#pragma unroll 8
      for (int i = 0; i < ROWS / RSTEP; i++) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
        src += src_step;
        dst += RSTEP * LD_K * ES;
      }
    } else {
      const int64_t kl = k_left - kc;
      const int kbytes = kl >= VE ? 16 : (kl > 0 ? (int)kl * ES : 0);
      for (int i = 0; i < ROWS / RSTEP; i++) {
        const int bytes = (r0 + i * RSTEP < mn_left) ? kbytes : 0;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(bytes ? src : g), "r"(bytes) : "memory");
        src += src_step;
        dst += RSTEP * LD_K * ES;
      }
    }
    cp_async_mbar_arrive_noinc(bar);
  }
}


}  // namespace b200
