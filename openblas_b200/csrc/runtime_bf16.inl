/*
 * runtime_bf16.inl -- host side of SBGEMV / SBDOT (included by runtime.cu inside extern "C"): staging of
 * host operands around the kernels of bf16_level12.cu.  A goes up with the strided-copy machinery of the
 * GEMM path (h2d_any: pinned slot ring for pageable memory); the vectors are gathered from / scattered to
 * the caller's increments through the pinned block, so the device sees unit strides; device-resident
 * operands are used in place with their own increments.  The reference does the same compress / expand
 * of strided vectors around its kernels (kernel/x86_64/sbgemv_n.c:100-132).
 */
static int run_sbgemv_on_context(Context *ctx, int trans, int64_t m, int64_t n, float alpha, const void *a, int64_t lda,
                                 const void *x, int64_t incx, float beta, void *y, int64_t incy) {
  const int64_t lenx = trans ? m : n, leny = trans ? n : m;
  const bool product = alpha != 0.f;
  if (!product && beta == 1.f) return 0;
  const PtrKind ka = product ? classify(a) : PTR_DEVICE, kx = product ? classify(x) : PTR_DEVICE, ky = classify(y);
  if (t_foreign) return (int)cudaErrorInvalidDevice;
  if ((product && (ka == PTR_DEVICE || kx == PTR_DEVICE)) || ky == PTR_DEVICE) { int e = order_after_caller(ctx); if (e) return e; }
  cudaStream_t s = ctx->stream;

  const int64_t lda_dev = ka == PTR_DEVICE ? lda : (int64_t)(round_up((size_t)m * 2, 128) / 2);
  const size_t a_bytes = (product && ka != PTR_DEVICE) ? round_up((size_t)lda_dev * (size_t)n * 2, 256) : 0;
  const size_t x_bytes = (product && kx != PTR_DEVICE) ? round_up((size_t)lenx * 2, 256) : 0;
  const size_t y_bytes = ky != PTR_DEVICE ? round_up((size_t)leny * 4, 256) : 0;
  const size_t ws_bytes = product ? round_up(sbgemv_workspace_bytes(trans, m, n), 256) : 0;
  int err = reserve_device(ctx, a_bytes + x_bytes + y_bytes + ws_bytes);
  if (err) return err;
  if (x_bytes + y_bytes) { if ((err = reserve_pinned(ctx, x_bytes + y_bytes))) return err; }
  char *d_a = ctx->dws, *d_x = d_a + a_bytes, *d_y = d_x + x_bytes, *d_ws = d_y + y_bytes;
  char *h_x = ctx->hws, *h_y = ctx->hws + x_bytes;

  if (a_bytes) {
    if ((err = h2d_any(ctx, s, ka, d_a, (size_t)lda_dev * 2, (const char *)a, (size_t)lda * 2, (size_t)m * 2, (size_t)n))) return err;
  }
  if (x_bytes) {
    const uint16_t *xs = (const uint16_t *)x;
    uint16_t *xd = (uint16_t *)h_x;
    for (int64_t i = 0; i < lenx; i++) xd[i] = xs[i * incx];
    CK(cudaMemcpyAsync(d_x, h_x, (size_t)lenx * 2, cudaMemcpyHostToDevice, s));
  }
  if (y_bytes && beta != 0.f) {
    const float *ys = (const float *)y;
    float *yd = (float *)h_y;
    for (int64_t i = 0; i < leny; i++) yd[i] = ys[i * incy];
    CK(cudaMemcpyAsync(d_y, h_y, (size_t)leny * 4, cudaMemcpyHostToDevice, s));
  }
  CK(launch_sbgemv(trans, m, n, alpha, a_bytes ? d_a : a, lda_dev, x_bytes ? d_x : x, x_bytes ? 1 : incx, beta,
                   y_bytes ? (void *)d_y : y, y_bytes ? 1 : incy, d_ws, s));
  if (y_bytes) {
    CK(cudaMemcpyAsync(h_y, d_y, (size_t)leny * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    float *ys = (float *)y;
    const float *yd = (const float *)h_y;
    for (int64_t i = 0; i < leny; i++) ys[i * incy] = yd[i];
    return 0;
  }
  CK(cudaStreamSynchronize(s));
  return 0;
}

B200_HIDDEN int b200_run_sbgemv(int trans, int64_t m, int64_t n, float alpha, const void *a, int64_t lda, const void *x,
                                int64_t incx, float beta, void *y, int64_t incy) {
  if (alpha == 0.f && beta == 1.f) return 0;             /* nothing to do: return before CUDA is touched */
  ContextLease lease;
  int err = acquire(&lease);
  if (err) return err;
  t_error[0] = 0;
  err = run_sbgemv_on_context(lease.c, trans, m, n, alpha, a, lda, x, incx, beta, y, incy);
  lease.failed = err != 0;
  return err;
}

static int run_sbdot_on_context(Context *ctx, int64_t n, const void *x, int64_t incx, const void *y, int64_t incy, float *result) {
  const PtrKind kx = classify(x), ky = classify(y);
  if (t_foreign) return (int)cudaErrorInvalidDevice;
  if (kx == PTR_DEVICE || ky == PTR_DEVICE) { int e = order_after_caller(ctx); if (e) return e; }
  cudaStream_t s = ctx->stream;
  const size_t v_bytes = round_up((size_t)n * 2, 256);
  const size_t x_bytes = kx != PTR_DEVICE ? v_bytes : 0, y_bytes = ky != PTR_DEVICE ? v_bytes : 0;
  const size_t ws_bytes = round_up(sbdot_workspace_bytes(), 256);
  int err = reserve_device(ctx, x_bytes + y_bytes + ws_bytes + 256);
  if (err) return err;
  if ((err = reserve_pinned(ctx, x_bytes + y_bytes + 256))) return err;
  char *d_x = ctx->dws, *d_y = d_x + x_bytes, *d_ws = d_y + y_bytes, *d_res = d_ws + ws_bytes;
  char *h_x = ctx->hws, *h_y = h_x + x_bytes, *h_res = h_y + y_bytes;
  if (x_bytes) {
    const uint16_t *src = (const uint16_t *)x; uint16_t *dst = (uint16_t *)h_x;
    for (int64_t i = 0; i < n; i++) dst[i] = src[i * incx];
  }
  if (y_bytes) {
    const uint16_t *src = (const uint16_t *)y; uint16_t *dst = (uint16_t *)h_y;
    for (int64_t i = 0; i < n; i++) dst[i] = src[i * incy];
  }
  if (x_bytes + y_bytes) CK(cudaMemcpyAsync(d_x, h_x, x_bytes + y_bytes, cudaMemcpyHostToDevice, s));   /* adjacent in both blocks */
  CK(launch_sbdot(n, x_bytes ? d_x : x, x_bytes ? 1 : incx, y_bytes ? d_y : y, y_bytes ? 1 : incy, d_ws, (float *)d_res, s));
  CK(cudaMemcpyAsync(h_res, d_res, sizeof(float), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  *result = *(const float *)h_res;
  return 0;
}

B200_HIDDEN int b200_run_sbdot(int64_t n, const void *x, int64_t incx, const void *y, int64_t incy, float *result) {
  ContextLease lease;
  int err = acquire(&lease);
  if (err) return err;
  t_error[0] = 0;
  err = run_sbdot_on_context(lease.c, n, x, incx, y, incy, result);
  lease.failed = err != 0;
  return err;
}
