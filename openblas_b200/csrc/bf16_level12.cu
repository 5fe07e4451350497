/*
 * bf16_level12.cu -- SBGEMV and SBDOT: bf16 operands, fp32 accumulation and result.
 *
 * Replaces kernel/x86_64/sbgemv_n.c, sbgemv_t.c (+ the AVX512-BF16 micro-kernels sbgemv_*_microk_cooperlake*.c),
 * kernel/x86_64/sbdot.c and driver/level2/sbgemv_thread.c.  Both are HBM-bound (2 flops per 2 bytes of A): the
 * kernels are plain coalesced passes, the only design points are
 *   - every warp reads contiguous 128-byte pieces of a column of A (bf16x2 per lane when the column is 4-byte
 *     aligned, scalar loads otherwise);
 *   - enough CTAs for 148 SMs whatever the shape: the reduction dimension is cut into S slices (grid.y) whose
 *     partial sums land in a workspace and are added IN SLICE ORDER by a finishing kernel, which also applies
 *     alpha and beta (beta == 0 never reads y) -- no atomics, so results are run-to-run identical, as the
 *     reference's are for a fixed thread count.
 */
#include <cuda_bf16.h>
#include "gemm_common.cuh"

namespace b200 {
namespace {

__device__ __forceinline__ float widen(uint16_t v) { return __uint_as_float((uint32_t)v << 16); }

/* N: partial[s][i] = sum over the slice's columns j of A(i, j) x(j); 128 threads x 2 rows per CTA */
__global__ void __launch_bounds__(128) sbgemv_n_kernel(int64_t m, int64_t n, const uint16_t *__restrict__ a, int64_t lda,
                                                       const uint16_t *__restrict__ x, int64_t incx, float *__restrict__ partial,
                                                       int64_t cols_per_slice, int pairs) {
  const int64_t i0 = ((int64_t)blockIdx.x * 128 + threadIdx.x) * 2;
  const int64_t j0 = (int64_t)blockIdx.y * cols_per_slice, j1 = j0 + cols_per_slice < n ? j0 + cols_per_slice : n;
  if (i0 >= m) return;
  float acc0 = 0.f, acc1 = 0.f;
  const bool two = i0 + 1 < m;
  if (pairs && two) {
    for (int64_t j = j0; j < j1; j++) {
      const uint32_t v = *(const uint32_t *)(a + i0 + j * lda);
      const float xv = widen(x[j * incx]);
      acc0 = fmaf(__uint_as_float(v << 16), xv, acc0);
      acc1 = fmaf(__uint_as_float(v & 0xffff0000u), xv, acc1);
    }
  } else {
    for (int64_t j = j0; j < j1; j++) {
      const float xv = widen(x[j * incx]);
      acc0 = fmaf(widen(a[i0 + j * lda]), xv, acc0);
      if (two) acc1 = fmaf(widen(a[i0 + 1 + j * lda]), xv, acc1);
    }
  }
  float *p = partial + (int64_t)blockIdx.y * m;
  p[i0] = acc0;
  if (two) p[i0 + 1] = acc1;
}

/* T: partial[s][j] = sum over the slice's rows i of A(i, j) x(i); one warp per column, 8 columns per CTA */
__global__ void __launch_bounds__(256) sbgemv_t_kernel(int64_t m, int64_t n, const uint16_t *__restrict__ a, int64_t lda,
                                                       const uint16_t *__restrict__ x, int64_t incx, float *__restrict__ partial,
                                                       int64_t rows_per_slice, int pairs) {
  const int lane = threadIdx.x & 31;
  const int64_t j = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (j >= n) return;
  const int64_t i0 = (int64_t)blockIdx.y * rows_per_slice, i1 = i0 + rows_per_slice < m ? i0 + rows_per_slice : m;
  const uint16_t *col = a + j * lda;
  float acc = 0.f;
  if (pairs && incx == 1) {          /* rows_per_slice is even, columns and x are 4-byte aligned */
    for (int64_t i = i0 + 2 * lane; i < i1; i += 64) {
      if (i + 1 < i1) {
        const uint32_t v = *(const uint32_t *)(col + i), xv = *(const uint32_t *)(x + i);
        acc = fmaf(__uint_as_float(v << 16), __uint_as_float(xv << 16), acc);
        acc = fmaf(__uint_as_float(v & 0xffff0000u), __uint_as_float(xv & 0xffff0000u), acc);
      } else {
        acc = fmaf(widen(col[i]), widen(x[i]), acc);
      }
    }
  } else {
    for (int64_t i = i0 + lane; i < i1; i += 32) acc = fmaf(widen(col[i]), widen(x[i * incx]), acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) partial[(int64_t)blockIdx.y * n + j] = acc;
}

/* y(i) = alpha * (partial[0][i] + partial[1][i] + ...) + beta * y(i), slices added in order */
__global__ void __launch_bounds__(256) sbgemv_finish_kernel(int64_t len, int slices, const float *__restrict__ partial, float alpha,
                                                            float beta, float *__restrict__ y, int64_t incy) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < len; i += (int64_t)gridDim.x * 256) {
    float acc = 0.f;
    for (int s = 0; s < slices; s++) acc += partial[(int64_t)s * len + i];
    float *p = y + i * incy;
    *p = beta == 0.f ? alpha * acc : alpha * acc + beta * *p;
  }
}

/* y := beta * y (alpha == 0: interface/sbgemv.c:173-176; beta == 0 writes zeros without reading y) */
__global__ void __launch_bounds__(256) scale_vector_kernel(int64_t len, float beta, float *__restrict__ y, int64_t incy) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < len; i += (int64_t)gridDim.x * 256) {
    float *p = y + i * incy;
    *p = beta == 0.f ? 0.f : beta * *p;
  }
}

/* SBDOT: per-CTA partial sums in a fixed order, then one CTA adds the partials in order */
__global__ void __launch_bounds__(256) sbdot_partial_kernel(int64_t n, const uint16_t *__restrict__ x, int64_t incx,
                                                            const uint16_t *__restrict__ y, int64_t incy, float *__restrict__ partial) {
  __shared__ float warp_sum[8];
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
    acc = fmaf(widen(x[i * incx]), widen(y[i * incy]), acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; w++) s += warp_sum[w];
    partial[blockIdx.x] = s;
  }
}
__global__ void sbdot_finish_kernel(int blocks, const float *__restrict__ partial, float *__restrict__ result) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float s = 0.f;
    for (int b = 0; b < blocks; b++) s += partial[b];
    *result = s;
  }
}

}  // namespace

/* number of slices the reduction dimension (`red` long) is cut into when `par` independent outputs give
 * `ctas` CTAs: enough for two CTAs per SM, each slice at least 256 long */
static int slices_for(int64_t ctas, int64_t red) {
  const int64_t want = 2 * (int64_t)sm_count();
  int64_t s = ctas >= want ? 1 : (want + ctas - 1) / ctas;
  const int64_t cap = (red + 255) / 256;
  if (s > cap) s = cap;
  if (s > 64) s = 64;
  return (int)(s < 1 ? 1 : s);
}

size_t sbgemv_workspace_bytes(int trans, int64_t m, int64_t n) {
  const int64_t len = trans ? n : m;
  const int64_t ctas = trans ? (n + 7) / 8 : (m + 255) / 256;
  return (size_t)slices_for(ctas, trans ? m : n) * (size_t)len * sizeof(float);
}

/* device pointers; x, y at their logical first element; workspace from sbgemv_workspace_bytes() */
cudaError_t launch_sbgemv(int trans, int64_t m, int64_t n, float alpha, const void *a, int64_t lda, const void *x, int64_t incx,
                          float beta, void *y, int64_t incy, void *workspace, cudaStream_t stream) {
  const int64_t len = trans ? n : m, red = trans ? m : n;
  if (alpha == 0.f) {
    if (beta == 1.f) return cudaSuccess;
    int64_t blocks = (len + 255) / 256;
    if (blocks > 4 * (int64_t)sm_count()) blocks = 4 * (int64_t)sm_count();
    scale_vector_kernel<<<(int)blocks, 256, 0, stream>>>(len, beta, (float *)y, incy);
    count_launch("scale_vector");
    return cudaGetLastError();
  }
  const int64_t ctas = trans ? (n + 7) / 8 : (m + 255) / 256;
  const int slices = slices_for(ctas, red);
  int64_t per = (red + slices - 1) / slices;
  per = (per + 1) & ~(int64_t)1;                 /* even: slices of a column keep the 4-byte alignment of its start */
  const int pairs = (((uintptr_t)a & 3) == 0) && (lda % 2 == 0) && (!trans || (((uintptr_t)x & 3) == 0));
  if (ctas > 2147483647LL) return cudaErrorInvalidConfiguration;
  dim3 grid((unsigned)ctas, (unsigned)slices);
  if (!trans) sbgemv_n_kernel<<<grid, 128, 0, stream>>>(m, n, (const uint16_t *)a, lda, (const uint16_t *)x, incx, (float *)workspace, per, pairs);
  else sbgemv_t_kernel<<<grid, 256, 0, stream>>>(m, n, (const uint16_t *)a, lda, (const uint16_t *)x, incx, (float *)workspace, per, pairs);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  count_launch(trans ? "sbgemv_t" : "sbgemv_n");
  int64_t blocks = (len + 255) / 256;
  if (blocks > 4 * (int64_t)sm_count()) blocks = 4 * (int64_t)sm_count();
  sbgemv_finish_kernel<<<(int)blocks, 256, 0, stream>>>(len, slices, (const float *)workspace, alpha, beta, (float *)y, incy);
  count_launch("sbgemv_finish");
  return cudaGetLastError();
}

size_t sbdot_workspace_bytes() { return (size_t)(2 * sm_count() + 1) * sizeof(float); }

/* result: device float (inside the workspace or anywhere else on the device) */
cudaError_t launch_sbdot(int64_t n, const void *x, int64_t incx, const void *y, int64_t incy, void *workspace, float *result,
                         cudaStream_t stream) {
  int64_t blocks = (n + 255) / 256;
  if (blocks > 2 * (int64_t)sm_count()) blocks = 2 * (int64_t)sm_count();
  if (blocks < 1) blocks = 1;
  sbdot_partial_kernel<<<(int)blocks, 256, 0, stream>>>(n, (const uint16_t *)x, incx, (const uint16_t *)y, incy, (float *)workspace);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  count_launch("sbdot_partial");
  sbdot_finish_kernel<<<1, 32, 0, stream>>>((int)blocks, (const float *)workspace, result);
  count_launch("sbdot_finish");
  return cudaGetLastError();
}

}  // namespace b200
