/*
 * runtime_level3.inl -- device side of SYMM/HEMM, SYRK/HERK, SYR2K/HER2K, TRMM/TRSM (included by
 * runtime.cu inside namespace b200).  These routines share the GEMM kernels in the reference too
 * (driver/level3/symm_k.c runs the GEMM loop nest over a symmetric packer; level3_syrk.c and
 * level3_syr2k.c run it over the triangle with syrk_kernel.c masking the diagonal blocks; trmm_L.c /
 * trsm_L.c block the triangular matrix around the same micro-kernel); here:
 *
 *   SYMM / HEMM   the referenced triangle is expanded ONCE into a full matrix in workspace
 *                 (O(ka^2) bytes of HBM traffic against O(ka^2 * n) flops) and the product is one
 *                 GEMM at the GEMM kernels' speed;
 *   SYRK family   preferred: ONE GEMM launch per product whose kernel walks only the tiles of the
 *                 triangle and masks the stores of the tiles the diagonal crosses (DeviceGemm::tri).
 *                 Fallback when the eligible kernel cannot mask (tiny or unaligned problems): C is cut
 *                 into block columns of width NB; the rectangle strictly inside the triangle is a
 *                 plain GEMM straight into C, the NB x NB diagonal block is computed in full into a
 *                 scratch tile and merged under the triangle mask (tri_merge), which also applies
 *                 beta and, for HERK/HER2K, zeroes the diagonal's imaginary part;
 *   TRMM / TRSM   recursive halving of op(A) around one large GEMM per level, 128 x 128 (real) or 64 x 64 (complex) diagonal
 *                 blocks in tri_block_kernel (see tri_recurse below).
 *
 * Host operands are staged whole (A, B, and the full m x n rectangle of C -- the part of C
 * outside the triangle travels up and comes back bit-identical; rows beyond m never move).
 */

static cudaError_t gemm_on_device(int dtype, int ta, int tb, int64_t m, int64_t n, int64_t k, double ar, double ai,
                                  const void *a, int64_t lda, const void *b, int64_t ldb, double br, double bi, void *c,
                                  int64_t ldc, cudaStream_t s, int tri = 0) {
  if (m <= 0 || n <= 0) return cudaSuccess;
  DeviceGemm g;
  g.tri = tri;
  g.dtype = dtype; g.transa = ta; g.transb = tb; g.m = m; g.n = n; g.k = k;
  g.lda = lda; g.ldb = ldb; g.ldc = ldc; g.a = a; g.b = b; g.c = c;
  g.alpha_re = ar; g.alpha_im = ai; g.beta_re = br; g.beta_im = bi;
  return dispatch(g, s);
}

static int64_t rankk_block(int64_t n) { return n >= 4096 ? 512 : n >= 1024 ? 256 : 128; }

/* SYMM / HEMM: how many columns (side left) or rows (side right) of the symmetric operand are expanded at a time.
 * The whole ka x ka matrix when that fits B200_SYMM_FULL_BYTES (one GEMM); otherwise panels of about
 * B200_SYMM_PANEL_BYTES, i.e. ka / width GEMMs with inner dimension `width` accumulating into C -- a 65536^2 ZHEMM
 * then needs 0.5 GiB of workspace instead of 64 GiB (round 1 expanded in full, always). */
#ifdef B200_HOSTSIM
#define B200_SYMM_FULL_BYTES  ((size_t)16 << 10)
#define B200_SYMM_PANEL_BYTES ((size_t)8 << 10)
#define B200_SYMM_PANEL_ALIGN 8
#else
#define B200_SYMM_FULL_BYTES  ((size_t)1 << 30)
#define B200_SYMM_PANEL_BYTES ((size_t)512 << 20)
#define B200_SYMM_PANEL_ALIGN 256
#endif
static int64_t symm_panel_width(int64_t ka, size_t es) {
  /* B200_SYMM_FULL_MB / B200_SYMM_PANEL_MB override the limits (tests force the panel scheme at small sizes) */
  static const size_t full_bytes = getenv("B200_SYMM_FULL_MB") ? (size_t)atol(getenv("B200_SYMM_FULL_MB")) << 20 : B200_SYMM_FULL_BYTES;
  static const size_t panel_bytes = getenv("B200_SYMM_PANEL_MB") ? (size_t)atol(getenv("B200_SYMM_PANEL_MB")) << 20 : B200_SYMM_PANEL_BYTES;
  if (round_up((size_t)ka * es, 128) * (size_t)ka <= full_bytes) return ka;
  int64_t w = (int64_t)(panel_bytes / (round_up((size_t)ka * es, 128))) / B200_SYMM_PANEL_ALIGN * B200_SYMM_PANEL_ALIGN;
  if (w < B200_SYMM_PANEL_ALIGN) w = B200_SYMM_PANEL_ALIGN;
  return w < ka ? w : ka;
}
static size_t symm_scratch_bytes(int64_t ka, size_t es) {
  const int64_t w = symm_panel_width(ka, es);
  const size_t left = round_up((size_t)ka * es, 128) * (size_t)w, right = round_up((size_t)w * es, 128) * (size_t)ka;
  return round_up(left > right ? left : right, 256);
}

/* ---- TRMM / TRSM: recursive splitting of the triangular matrix E = op(A) ------------------------
 * (the reference blocks the same routines over its GEMM kernel: driver/level3/trmm_L.c, trmm_R.c,
 * trsm_L.c, trsm_R.c).  E is cut at a multiple of 64 near the middle; the off-diagonal block is ONE
 * GEMM with a large inner dimension (that is where the flops are), the two diagonal halves recurse,
 * diagonal blocks of at most tri_block_max() (128 real, 64 complex) go to the block kernels of level3_aux.cu.  The order of the three steps is what makes the
 * update valid in place:
 *   TRMM  left,  E lower:  B2 := a E22 B2;  B2 += a E21 B1;  B1 := a E11 B1      (upper: mirrored)
 *   TRMM  right, E lower:  B1 := a B1 E11;  B1 += a B2 E21;  B2 := a B2 E22
 *   TRSM  left,  E lower:  X1 = a E11^-1 B1;  B2 := a B2 - E21 X1;  X2 = E22^-1 B2
 *   TRSM  right, E lower:  X2 = a B2 E22^-1;  B1 := a B1 - X2 E21;  X1 = B1 E11^-1 */
struct TriWork {
  int dtype, op; bool solve, left, eff_lower, unit;
  const char *a; int64_t lda; char *b; int64_t ldb; int64_t m, n; size_t es; cudaStream_t s;
  /* rows [r0, ..) x cols [c0, ..) of E as a GEMM operand that still needs `op` applied */
  const char *eblk(int64_t r0, int64_t c0) const { return a + ((op & 1) ? ((size_t)c0 + (size_t)r0 * lda) : ((size_t)r0 + (size_t)c0 * lda)) * es; }
  char *bpart(int64_t off) const { return b + (left ? (size_t)off : (size_t)off * ldb) * es; }   /* rows (left) or columns (right) from off */
};

/* B(dst part) := beta * B(dst part) + alpha * [E(r0.., c0..) applied to B(src part)] */
static cudaError_t tri_gemm(const TriWork &w, int64_t r0, int64_t nr, int64_t c0, int64_t nc, double ar, double ai, double br, double bi) {
  if (w.left)    /* E block (nr x nc) times rows c0.. of B into rows r0.. */
    return gemm_on_device(w.dtype, w.op, B200_N, nr, w.n, nc, ar, ai, w.eblk(r0, c0), w.lda, w.bpart(c0), w.ldb, br, bi, w.bpart(r0), w.ldb, w.s);
  /* columns r0.. of B times E block (nr x nc) into columns c0.. */
  return gemm_on_device(w.dtype, B200_N, w.op, w.m, nc, nr, ar, ai, w.bpart(r0), w.ldb, w.eblk(r0, c0), w.lda, br, bi, w.bpart(c0), w.ldb, w.s);
}

static int tri_recurse(const TriWork &w, int64_t off, int64_t size, double ar, double ai) {
  const int64_t base = tri_block_max(w.dtype);       /* 128 (real types) or 64 */
  if (size <= base) {
    const bool tr = (w.op & 1) != 0;
    /* left: E(i,k) = op(F)(i,k); right: the kernel works on E^T */
    const int64_t fs_i = (w.left ? tr : !tr) ? w.lda : 1, fs_k = (w.left ? tr : !tr) ? 1 : w.lda;
    CK(launch_tri_block(w.dtype, w.solve, (int)size, w.left ? w.n : w.m, w.left ? w.eff_lower : !w.eff_lower, w.unit, w.op >= 2,
                        w.a + ((size_t)off + (size_t)off * w.lda) * w.es, fs_i, fs_k, ar, ai, w.bpart(off), w.left ? 1 : w.ldb,
                        w.left ? w.ldb : 1, w.s));
    return 0;
  }
  const int64_t s1 = ((size / 2 + base - 1) / base) * base, s2 = size - s1, o2 = off + s1;
  int err;
  /* which half must be finished first, and which off-diagonal block links them */
  const bool lower_first_is_2 = w.solve ? !w.left : w.left;     /* for E lower: does part 2 go first? */
  const bool part2_first = w.eff_lower ? lower_first_is_2 : !lower_first_is_2;
  const int64_t fo = part2_first ? o2 : off, fs = part2_first ? s2 : s1;      /* first part */
  const int64_t lo = part2_first ? off : o2, ls = part2_first ? s1 : s2;      /* last part */
  if (!w.solve) {
    /* the first part's product reads the OTHER part's original data in the GEMM, so: first part's own
     * diagonal product, then the coupling GEMM into it, then the other part's diagonal product */
    if ((err = tri_recurse(w, fo, fs, ar, ai))) return err;
    if (w.left) CK(tri_gemm(w, fo, fs, lo, ls, ar, ai, 1.0, 0.0));     /* B_first += a E(first, last) B_last */
    else        CK(tri_gemm(w, lo, ls, fo, fs, ar, ai, 1.0, 0.0));     /* B_first += a B_last E(last, first) */
    return tri_recurse(w, lo, ls, ar, ai);
  }
  if ((err = tri_recurse(w, fo, fs, ar, ai))) return err;               /* X_first */
  if (w.left) CK(tri_gemm(w, lo, ls, fo, fs, -1.0, 0.0, ar, ai));        /* B_last := a B_last - E(last, first) X_first */
  else        CK(tri_gemm(w, fo, fs, lo, ls, -1.0, 0.0, ar, ai));        /* B_last := a B_last - X_first E(first, last) */
  return tri_recurse(w, lo, ls, 1.0, 0.0);
}

/* everything on the device, pointers are device pointers; scratch holds the expanded operand
 * (SYMM/HEMM) or one NB x NB tile (the others) */
static int level3_on_device(const b200_l3_problem *p, const char *a, int64_t lda, const char *b, int64_t ldb, char *c,
                            int64_t ldc, char *scratch, cudaStream_t s) {
  const size_t es = b200_in_size(p->dtype);
  const double ar = p->alpha[0], ai = p->alpha[1], br = p->beta[0], bi = p->beta[1];
  const bool alpha_zero = ar == 0.0 && ai == 0.0;

  if (p->routine == B200_SYMM || p->routine == B200_HEMM) {
    if (alpha_zero) {                  /* C := beta * C (beta == 0: exact zeros, C not read) */
      CK(gemm_on_device(p->dtype, 0, 0, p->m, p->n, 0, 0.0, 0.0, c, ldc, c, ldc, br, bi, c, ldc, s));
      return 0;
    }
    const int64_t ka = p->side ? p->n : p->m;
    const int64_t w = symm_panel_width(ka, es);
    if (w >= ka) {                     /* the whole operand at once: ONE GEMM */
      const int64_t ldf = (int64_t)(round_up((size_t)ka * es, 128) / es);
      CK(launch_expand_symmetric(p->dtype, p->uplo, p->routine == B200_HEMM, ka, a, lda, scratch, ldf, s));
      if (!p->side) CK(gemm_on_device(p->dtype, 0, 0, p->m, p->n, p->m, ar, ai, scratch, ldf, b, ldb, br, bi, c, ldc, s));
      else CK(gemm_on_device(p->dtype, 0, 0, p->m, p->n, p->n, ar, ai, b, ldb, scratch, ldf, br, bi, c, ldc, s));
      return 0;
    }
    /* panel by panel along the inner dimension: C := alpha A(:, K) B(K, :) + beta' C  /  alpha B(:, K) A(K, :) + beta' C */
    for (int64_t k0 = 0; k0 < ka; k0 += w) {
      const int64_t kw = ka - k0 < w ? ka - k0 : w;
      const double pbr = k0 ? 1.0 : br, pbi = k0 ? 0.0 : bi;
      if (!p->side) {
        const int64_t ldf = (int64_t)(round_up((size_t)ka * es, 128) / es);
        CK(launch_expand_symmetric(p->dtype, p->uplo, p->routine == B200_HEMM, ka, a, lda, scratch, ldf, s, 0, ka, k0, kw));
        CK(gemm_on_device(p->dtype, 0, 0, p->m, p->n, kw, ar, ai, scratch, ldf, b + (size_t)k0 * es, ldb, pbr, pbi, c, ldc, s));
      } else {
        const int64_t ldw = (int64_t)(round_up((size_t)kw * es, 128) / es);
        CK(launch_expand_symmetric(p->dtype, p->uplo, p->routine == B200_HEMM, ka, a, lda, scratch, ldw, s, k0, kw, 0, ka));
        CK(gemm_on_device(p->dtype, 0, 0, p->m, p->n, kw, ar, ai, b + (size_t)k0 * (size_t)ldb * es, ldb, scratch, ldw, pbr, pbi, c, ldc, s));
      }
    }
    return 0;
  }

  if (p->routine == B200_TRMM || p->routine == B200_TRSM) {
    if (alpha_zero) {                  /* B := 0, exact zeros, B not read (the reference's GEMM_BETA pass, trsm_L.c:101-106) */
      CK(gemm_on_device(p->dtype, 0, 0, p->m, p->n, 0, 0.0, 0.0, c, ldc, c, ldc, 0.0, 0.0, c, ldc, s));
      return 0;
    }
    TriWork w;
    w.dtype = p->dtype; w.op = p->trans; w.solve = p->routine == B200_TRSM; w.left = !p->side;
    w.eff_lower = (p->uplo != 0) != ((p->trans & 1) != 0); w.unit = p->unit != 0;
    w.a = a; w.lda = lda; w.b = c; w.ldb = ldc; w.m = p->m; w.n = p->n; w.es = es; w.s = s;
    return tri_recurse(w, 0, p->side ? p->n : p->m, ar, ai);
  }

  /* SYRK family and GEMMT: the uplo triangle of the n x n matrix C := alpha op1(X) op2(Y) [+ second product] + beta C */
  const bool herm = p->routine == B200_HERK || p->routine == B200_HER2K;
  const bool two = p->routine == B200_SYR2K || p->routine == B200_HER2K;
  const bool gemmt = p->routine == B200_GEMMT;
  const int64_t n = p->n, k = p->k;
  /* SBGEMMT: bf16 factors (es = 2), fp32 C: C addresses, the scratch tile and the merge are fp32's */
  const size_t esc = b200_out_size(p->dtype);
  const int merge_dtype = p->dtype == B200_SB ? B200_S : p->dtype;
  const bool product = k > 0 && !alpha_zero;
  if (!product) {                      /* only the triangle is scaled; beta == 1 leaves C alone */
    if (br == 1.0 && bi == 0.0) return 0;
    CK(launch_tri_merge(merge_dtype, p->uplo, herm, n, nullptr, 0, br, bi, c, ldc, s));
    return 0;
  }
  /* first factor op(X) (rows of C), second factor op(Y)^T or op(Y)^H (columns of C):
   *   trans == 0: X is n x k, rows i0.. start at X + i0;        first N, second T (or C)
   *   trans == 1: X is k x n, rows i0.. start at X + i0 * ldx;  first T (or C), second N */
  /* GEMMT (interface/gemmt.c): op1 = op(A), op2 = op(B) as given, Y = B */
  const int op_first = gemmt ? p->trans : p->trans ? (herm ? B200_C_ : B200_T) : B200_N;
  const int op_second = gemmt ? p->transb : p->trans ? B200_N : (herm ? B200_C_ : B200_T);
  /* rows i0.. of op1(X) / columns j0.. of op2(Y) as GEMM operands */
  auto at = [&](const char *x, int64_t ldx, int64_t i0) { return x + ((op_first & 1) ? (size_t)i0 * (size_t)ldx : (size_t)i0) * es; };
  auto at2 = [&](const char *y, int64_t ldy, int64_t j0) { return y + ((op_second & 1) ? (size_t)j0 : (size_t)j0 * (size_t)ldy) * es; };
  const double ai2 = herm ? -ai : ai;  /* HER2K: the second product carries conj(alpha) */
  const char *y1 = (two || gemmt) ? b : a;            /* second factor of the first product */
  const int64_t ldy1 = (two || gemmt) ? ldb : lda;

  /* Preferred: ONE launch per product with the triangle handled inside the GEMM kernel (tiles
   * outside the triangle skipped, stores of the diagonal tiles masked): no small diagonal-block
   * GEMMs, no scratch, the whole triangle's tiles share the SMs.  Kernels that cannot mask (generic
   * kernel for tiny problems, cp.async variants for unaligned operands) answer "not supported"
   * BEFORE launching anything, and the block-column scheme below takes over. */
  const char *tri_env = getenv("B200_RANKK_TRI");          /* =0 forces the block-column scheme (tests cover both) */
  const bool tri_gemm_ok = !(tri_env && atoi(tri_env) == 0);
  if (tri_gemm_ok) {
    const int tri = p->uplo ? 1 : 2;
    cudaError_t e = gemm_on_device(p->dtype, op_first, op_second, n, n, k, ar, ai, a, lda, y1, ldy1, br, bi, c, ldc, s, tri);
    if (e == cudaSuccess) {
      if (two) CK(gemm_on_device(p->dtype, op_first, op_second, n, n, k, ar, ai2, b, ldb, a, lda, 1.0, 0.0, c, ldc, s, tri));
      if (herm) CK(launch_real_diagonal(p->dtype, n, c, ldc, s));
      return 0;
    }
    if (e != cudaErrorNotSupported) return set_error(e, "triangular GEMM");
  }

  const int64_t nb = rankk_block(n);
  const int64_t ldt = (int64_t)(round_up((size_t)nb * esc, 128) / esc);
  for (int64_t j0 = 0; j0 < n; j0 += nb) {
    const int64_t jb = n - j0 < nb ? n - j0 : nb;
    /* rectangle inside the triangle: rows below the diagonal block (lower) or above it (upper) */
    const int64_t i0 = p->uplo ? j0 + jb : 0, mr = p->uplo ? n - j0 - jb : j0;
    char *c_rect = c + ((size_t)i0 + (size_t)j0 * (size_t)ldc) * esc;
    if (mr > 0) {
      CK(gemm_on_device(p->dtype, op_first, op_second, mr, jb, k, ar, ai, at(a, lda, i0), lda, at2(y1, ldy1, j0), ldy1, br, bi, c_rect, ldc, s));
      if (two) CK(gemm_on_device(p->dtype, op_first, op_second, mr, jb, k, ar, ai2, at(b, ldb, i0), ldb, at2(a, lda, j0), lda, 1.0, 0.0,
                                 c_rect, ldc, s));
    }
    /* diagonal block through the scratch tile */
    CK(gemm_on_device(p->dtype, op_first, op_second, jb, jb, k, ar, ai, at(a, lda, j0), lda, at2(y1, ldy1, j0), ldy1, 0.0, 0.0, scratch, ldt, s));
    if (two) CK(gemm_on_device(p->dtype, op_first, op_second, jb, jb, k, ar, ai2, at(b, ldb, j0), ldb, at2(a, lda, j0), lda, 1.0, 0.0,
                               scratch, ldt, s));
    CK(launch_tri_merge(merge_dtype, p->uplo, herm, jb, scratch, ldt, br, bi, c + ((size_t)j0 + (size_t)j0 * (size_t)ldc) * esc, ldc, s));
  }
  return 0;
}

static int run_level3_on_context(Context *ctx, const b200_l3_problem *p) {
  const size_t es = b200_in_size(p->dtype);
  const bool trxm = p->routine == B200_TRMM || p->routine == B200_TRSM;
  const bool symm = p->routine == B200_SYMM || p->routine == B200_HEMM;
  const bool two = p->routine == B200_SYR2K || p->routine == B200_HER2K;
  const bool gemmt = p->routine == B200_GEMMT;
  const bool alpha_zero = p->alpha[0] == 0.0 && p->alpha[1] == 0.0;
  const bool beta_one = p->beta[0] == 1.0 && p->beta[1] == 0.0, beta_zero = p->beta[0] == 0.0 && p->beta[1] == 0.0;
  const bool product = !alpha_zero && (symm || trxm || p->k > 0);
  if (!product && beta_one && !trxm) return 0;

  Operand A, B, C;
  A.es = B.es = es;
  C.es = b200_out_size(p->dtype);      /* differs from es only for SBGEMMT (bf16 in, fp32 out) */
  if (symm || trxm) {
    const int64_t ka = p->side ? p->n : p->m;
    A.rows = A.cols = ka; B.rows = p->m; B.cols = p->n;
  } else {
    A.rows = (p->trans & 1) ? p->k : p->n; A.cols = (p->trans & 1) ? p->n : p->k;
    B.rows = A.rows; B.cols = A.cols;
    if (gemmt) { B.rows = (p->transb & 1) ? p->n : p->k; B.cols = (p->transb & 1) ? p->k : p->n; }
  }
  C.rows = p->m; C.cols = p->n;
  A.ld_user = p->lda; B.ld_user = p->ldb; C.ld_user = p->ldc;
  const bool use_b = product && (symm || two || gemmt);
  A.kind = product ? classify(p->a) : PTR_DEVICE;
  B.kind = use_b ? classify(p->b) : PTR_DEVICE;
  C.kind = classify(p->c);
  if (t_foreign) return (int)cudaErrorInvalidDevice;
  if ((product && A.kind == PTR_DEVICE) || (use_b && B.kind == PTR_DEVICE) || C.kind == PTR_DEVICE) { int oe = order_after_caller(ctx); if (oe) return oe; }

  Operand *ops[3] = {&A, &B, &C};
  const void *user[3] = {p->a, p->b, p->c};
  const bool needed[3] = {product, use_b, true};
  size_t need = 0;
  for (int i = 0; i < 3; i++) {
    Operand &o = *ops[i];
    if (!needed[i]) continue;
    if (o.kind == PTR_DEVICE) { o.dev = (char *)user[i]; o.ld_dev = o.ld_user; continue; }
    o.host = (const char *)user[i];
    o.ld_dev = (int64_t)(round_up((size_t)(o.rows > 0 ? o.rows : 1) * o.es, 128) / o.es);
    need += round_up(o.bytes_dev(), 256);
  }
  size_t scratch_bytes = 0;
  if (product && !trxm) {
    const int64_t edge = rankk_block(p->n);
    scratch_bytes = symm ? symm_scratch_bytes(A.rows, es) : round_up(round_up((size_t)edge * C.es, 128) * (size_t)edge, 256);
  }
  int err = reserve_device(ctx, need + scratch_bytes);
  if (err) return err;
  size_t off = 0;
  for (int i = 0; i < 3; i++) {
    Operand &o = *ops[i];
    if (!needed[i] || o.kind == PTR_DEVICE) continue;
    o.dev = ctx->dws + off;
    off += round_up(o.bytes_dev(), 256);
  }
  char *scratch = ctx->dws + need;

  cudaStream_t s = ctx->stream;
  /* C goes up unless it is written in full without being read (SYMM/HEMM with beta == 0): the
   * triangular routines bring the whole rectangle back, so the untouched triangle must be there */
  const bool c_up = trxm ? product : !(symm && beta_zero);
  const bool small = need > 0 && need <= kSmallBytes;
  if (small) {
    if ((err = reserve_pinned(ctx, need))) return err;
    size_t up_begin = (size_t)-1, up_end = 0;
    for (int i = 0; i < 3; i++) {
      Operand &o = *ops[i];
      if (!needed[i] || o.kind == PTR_DEVICE || (i == 2 && !c_up)) continue;
      const size_t o_off = (size_t)(o.dev - ctx->dws);
      pack_to(ctx->hws + o_off, o);
      if (o_off < up_begin) up_begin = o_off;
      if (o_off + o.bytes_dev() > up_end) up_end = o_off + o.bytes_dev();
    }
    if (up_end > up_begin) CK(cudaMemcpyAsync(ctx->dws + up_begin, ctx->hws + up_begin, up_end - up_begin, cudaMemcpyHostToDevice, s));
  } else {
    for (int i = 0; i < 3; i++) {
      Operand &o = *ops[i];
      if (!needed[i] || o.kind == PTR_DEVICE || (i == 2 && !c_up)) continue;
      if ((err = h2d_any(ctx, s, o.kind, o.dev, (size_t)o.ld_dev * o.es, o.host, (size_t)o.ld_user * o.es, (size_t)o.rows * o.es,
                         (size_t)o.cols))) return err;
    }
  }
  if ((err = level3_on_device(p, A.dev, A.ld_dev, B.dev, B.ld_dev, C.dev, C.ld_dev, scratch, s))) return err;
  if (C.kind != PTR_DEVICE) {
    if (small) {
      const size_t c_off = (size_t)(C.dev - ctx->dws);
      CK(cudaMemcpyAsync(ctx->hws + c_off, C.dev, C.bytes_dev(), cudaMemcpyDeviceToHost, s));
      CK(cudaStreamSynchronize(s));
      const size_t row_bytes = (size_t)C.rows * C.es;
      char *uc = (char *)p->c;
      for (int64_t j = 0; j < C.cols; j++)
        memcpy(uc + (size_t)j * (size_t)C.ld_user * C.es, ctx->hws + c_off + (size_t)j * (size_t)C.ld_dev * C.es, row_bytes);
      return 0;
    }
    if ((err = d2h_any(ctx, s, C.kind, (char *)p->c, (size_t)C.ld_user * C.es, C.dev, (size_t)C.ld_dev * C.es, (size_t)C.rows * C.es,
                       (size_t)C.cols))) return err;
    if ((err = drain_all_out(ctx))) return err;
  }
  CK(cudaStreamSynchronize(s));
  return 0;
}
