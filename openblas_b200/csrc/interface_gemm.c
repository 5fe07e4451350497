/*
 * interface_gemm.c -- BLAS / CBLAS entry points of the GEMM path, host side, plain C.
 *
 * Replaces the reference's interface/gemm.c (compiled there once per precision and ABI
 * with -DDOUBLE/-DCOMPLEX/-DBFLOAT16/-DCBLAS/-DGEMM3M, interface/Makefile:1303-1326).
 * Here one table-driven implementation serves all of them: an entry point only packs its
 * arguments into a `call` record; normalise() restates the reference's argument handling
 *
 *   - op decoding, real types fold R->N and C->T        (interface/gemm.c:245-269, 387-406)
 *   - row-major = the transposed column-major problem   (interface/gemm.c:423-471)
 *   - argument checks, "last failing check wins" order  (interface/gemm.c:276-285, 411-420)
 *   - xerbla_(NAME, &info, sizeof(NAME)) then return    (interface/gemm.c:287-290, 473-476)
 *   - quick return when m == 0 or n == 0                (interface/gemm.c:494)
 *
 * and then hands the normalised column-major problem to the CUDA side (b200_run_problem).
 * What the reference does after this point on the CPU -- small-matrix kernels, packing
 * buffer, thread-count choice, the gemm[] table of 32 drivers (interface/gemm.c:551-642) --
 * has no counterpart here: the kernel choice is made on the device side by shape.
 * There is no CPU compute path in this file or behind it.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "shim.h"

/* Resolved at link/load time so that a program's own xerbla_ wins over the weak default in
 * xerbla.c (ctest/c_xerbla.c:131-135 relies on this). */
extern int xerbla_(char *name, blasint *info, blasint len);

static const char *const error_name[] = {
    /* interface/gemm.c:47-79 */
    "SGEMM ", "DGEMM ", "CGEMM ", "ZGEMM ", "SBGEMM ", "CGEMM3M ", "ZGEMM3M "};
enum { NAME_3M_OFFSET = 3 }; /* B200_C + 3 -> "CGEMM3M ", B200_Z + 3 -> "ZGEMM3M " */

typedef struct call {
  int dtype;
  int name_idx;
  int cblas;          /* 1: CBLAS conventions (order argument, info -1 == ok) */
  int order;          /* CBLAS only */
  int transa, transb; /* already decoded to 0..3, or -1 when illegal */
  int64_t m, n, k, lda, ldb, ldc;
  const void *alpha, *beta, *a, *b;
  void *c;
} call;

static int decode_cblas_op(int t, int complex_type) {
  /* interface/gemm.c:387-406 */
  switch (t) {
    case CblasNoTrans:     return B200_N;
    case CblasTrans:       return B200_T;
    case CblasConjNoTrans: return complex_type ? B200_R : B200_N;
    case CblasConjTrans:   return complex_type ? B200_C_ : B200_T;
    default:               return -1;
  }
}

static int decode_char_op(char ch, int complex_type) {
  /* interface/gemm.c:239-269 (TOUPPER then N/T/R/C) */
  if (ch >= 'a' && ch <= 'z') ch = (char)(ch - 'a' + 'A');
  switch (ch) {
    case 'N': return B200_N;
    case 'T': return B200_T;
    case 'R': return complex_type ? B200_R : B200_N;
    case 'C': return complex_type ? B200_C_ : B200_T;
    default:  return -1;
  }
}

/* The eight checks of interface/gemm.c:276-285 in their original order: later assignments
 * overwrite earlier ones, so the LOWEST-numbered failing argument is what gets reported. */
static blasint check_problem(const b200_problem *p, blasint ok_value) {
  int64_t nrowa = (p->transa & 1) ? p->k : p->m;
  int64_t nrowb = (p->transb & 1) ? p->n : p->k;
  blasint info = ok_value;
  if (p->ldc < p->m)   info = 13;
  if (p->ldb < nrowb)  info = 10;
  if (p->lda < nrowa)  info = 8;
  if (p->k < 0)        info = 5;
  if (p->n < 0)        info = 4;
  if (p->m < 0)        info = 3;
  if (p->transb < 0)   info = 2;
  if (p->transa < 0)   info = 1;
  return info;
}

/* Returns 1 when `out` holds a problem to run, 0 when the call is finished (error reported
 * or quick return). */
B200_HIDDEN int b200_normalise(const call *in, b200_problem *out) {
  blasint info;
  out->dtype = in->dtype;
  out->algo3m = in->name_idx != in->dtype;      /* the "CGEMM3M " / "ZGEMM3M " entry points */
  out->alpha = in->alpha;
  out->beta = in->beta;
  out->c = in->c;
  out->ldc = in->ldc;
  out->k = in->k;

  if (!in->cblas || in->order == CblasColMajor) {
    out->m = in->m; out->n = in->n;
    out->a = in->a; out->b = in->b;
    out->lda = in->lda; out->ldb = in->ldb;
    out->transa = in->transa; out->transb = in->transb;
    info = check_problem(out, in->cblas ? -1 : 0);
  } else if (in->order == CblasRowMajor) {
    /* C^T = op(B)^T op(A)^T : swap the operands and the two extents (gemm.c:423-447) */
    out->m = in->n; out->n = in->m;
    out->a = in->b; out->b = in->a;
    out->lda = in->ldb; out->ldb = in->lda;
    out->transa = in->transb; out->transb = in->transa;
    info = check_problem(out, -1);
  } else {
    info = 0; /* illegal order: the reference leaves info at 0 and still fires (gemm.c:370,473) */
  }

  if (in->cblas ? (info >= 0) : (info != 0)) {
    char name[16];
    size_t len = strlen(error_name[in->name_idx]) + 1; /* sizeof(ERROR_NAME) */
    memcpy(name, error_name[in->name_idx], len);
    xerbla_(name, &info, (blasint)len);
    return 0;
  }
  if (out->m == 0 || out->n == 0) return 0; /* gemm.c:494 */
  return 1;
}

static void run_call(const call *cl) {
  b200_problem p;
  if (!b200_normalise(cl, &p)) return;
  int err = b200_run_problem(&p);
  if (err) b200_fatal(error_name[cl->name_idx], err);
}

/* ---------------------------------------------------------------- CBLAS entry points */
#define CBLAS_REAL(NAME, DT, TIN, TOUT)                                                       \
  B200_EXPORT void NAME(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE ta,                      \
                        enum CBLAS_TRANSPOSE tb, blasint m, blasint n, blasint k, TOUT alpha, \
                        const TIN *a, blasint lda, const TIN *b, blasint ldb, TOUT beta,      \
                        TOUT *c, blasint ldc) {                                               \
    call cl = {DT, DT, 1, (int)order, decode_cblas_op((int)ta, 0), decode_cblas_op((int)tb, 0), \
               m, n, k, lda, ldb, ldc, &alpha, &beta, a, b, c};                               \
    run_call(&cl);                                                                            \
  }

#define CBLAS_CPLX(NAME, DT, NAMEIDX)                                                         \
  B200_EXPORT void NAME(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE ta,                      \
                        enum CBLAS_TRANSPOSE tb, blasint m, blasint n, blasint k,             \
                        const void *alpha, const void *a, blasint lda, const void *b,         \
                        blasint ldb, const void *beta, void *c, blasint ldc) {                \
    call cl = {DT, NAMEIDX, 1, (int)order, decode_cblas_op((int)ta, 1),                       \
               decode_cblas_op((int)tb, 1), m, n, k, lda, ldb, ldc, alpha, beta, a, b, c};    \
    run_call(&cl);                                                                            \
  }

CBLAS_REAL(cblas_sgemm, B200_S, float, float)
CBLAS_REAL(cblas_dgemm, B200_D, double, double)
CBLAS_REAL(cblas_sbgemm, B200_SB, bfloat16, float)
CBLAS_CPLX(cblas_cgemm, B200_C, B200_C)
CBLAS_CPLX(cblas_zgemm, B200_Z, B200_Z)
/* GEMM3M: same contract as GEMM.  Large products are computed from THREE real GEMMs (runtime.cu: gemm3m_on_device, the
 * scheme of driver/level3/gemm3m_level3.c: 25 % fewer multiplications, error bound in terms of (|re| + |im|) sums);
 * small ones with the regular 4-multiply complex kernel (the reference's generic targets forward GEMM3M to GEMM
 * altogether, Changelog.txt:9-10). */
CBLAS_CPLX(cblas_cgemm3m, B200_C, B200_C + NAME_3M_OFFSET)
CBLAS_CPLX(cblas_zgemm3m, B200_Z, B200_Z + NAME_3M_OFFSET)

/* -------------------------------------------------------------- Fortran entry points */
#define F77_GEMM(NAME, DT, NAMEIDX, TIN, TOUT, CPLX)                                          \
  B200_EXPORT void NAME(char *TRANSA, char *TRANSB, blasint *M, blasint *N, blasint *K,       \
                        TOUT *alpha, TIN *a, blasint *ldA, TIN *b, blasint *ldB, TOUT *beta,  \
                        TOUT *c, blasint *ldC) {                                              \
    call cl = {DT, NAMEIDX, 0, 0, decode_char_op(*TRANSA, CPLX), decode_char_op(*TRANSB, CPLX), \
               *M, *N, *K, *ldA, *ldB, *ldC, alpha, beta, a, b, c};                           \
    run_call(&cl);                                                                            \
  }

F77_GEMM(sgemm_, B200_S, B200_S, float, float, 0)
F77_GEMM(dgemm_, B200_D, B200_D, double, double, 0)
F77_GEMM(sbgemm_, B200_SB, B200_SB, bfloat16, float, 0)
F77_GEMM(cgemm_, B200_C, B200_C, float, float, 1)
F77_GEMM(zgemm_, B200_Z, B200_Z, double, double, 1)
F77_GEMM(cgemm3m_, B200_C, B200_C + NAME_3M_OFFSET, float, float, 1)
F77_GEMM(zgemm3m_, B200_Z, B200_Z + NAME_3M_OFFSET, double, double, 1)

/* ------------------------------------------------------------------- batched GEMM ----
 * interface/gemm_batch.c:118-372: per group one (trans, m, n, k, alpha, ld*, beta) tuple and
 * group_size[i] matrices; every group is validated like a single GEMM (error name
 * "?GEMM_BATCH ", first bad group aborts the whole call before anything is computed for
 * it or later groups -- earlier groups are not run either, since the reference only
 * launches after the validation loop, gemm_batch.c:283-287,366-368). */
static const char *const batch_error_name[] = {"SGEMM_BATCH ", "DGEMM_BATCH ", "CGEMM_BATCH ",
                                               "ZGEMM_BATCH ", "SBGEMM_BATCH "};

static void run_batch(int dtype, int order, const enum CBLAS_TRANSPOSE *ta_arr,
                      const enum CBLAS_TRANSPOSE *tb_arr, const blasint *m_arr,
                      const blasint *n_arr, const blasint *k_arr, const void *alpha_arr,
                      const void *const *a_arr, const blasint *lda_arr,
                      const void *const *b_arr, const blasint *ldb_arr, const void *beta_arr,
                      void *const *c_arr, const blasint *ldc_arr, blasint group_count,
                      const blasint *group_size) {
  int64_t total = 0, count = 0, idx = 0;
  int cplx = b200_is_complex(dtype);
  size_t scalar = b200_out_size(dtype);
  for (blasint g = 0; g < group_count; g++) total += group_size[g];
  /* no early return for an empty batch: the reference still validates every group of it
   * (gemm_batch.c:152-158 allocates, :186-287 checks group by group whatever the group sizes are) */
  b200_problem *probs = (b200_problem *)malloc((size_t)(total > 0 ? total : 1) * sizeof(b200_problem));
  if (!probs) { fprintf(stderr, "openblas_b200: gemm_batch: out of host memory\n"); return; }

  for (blasint g = 0; g < group_count; idx += group_size[g], g++) {
    b200_problem p;
    blasint info;
    memset(&p, 0, sizeof p);
    p.dtype = dtype;
    p.alpha = (const char *)alpha_arr + (size_t)g * scalar;
    p.beta = (const char *)beta_arr + (size_t)g * scalar;
    p.k = k_arr[g];
    p.ldc = ldc_arr[g];
    p.transa = p.transb = -1;
    if (order == CblasColMajor) {
      p.m = m_arr[g]; p.n = n_arr[g];
      p.lda = lda_arr[g]; p.ldb = ldb_arr[g];
      p.transa = decode_cblas_op((int)ta_arr[g], cplx);
      p.transb = decode_cblas_op((int)tb_arr[g], cplx);
      info = check_problem(&p, -1);
    } else if (order == CblasRowMajor) {
      p.m = n_arr[g]; p.n = m_arr[g];
      p.lda = ldb_arr[g]; p.ldb = lda_arr[g];
      p.transa = decode_cblas_op((int)tb_arr[g], cplx);
      p.transb = decode_cblas_op((int)ta_arr[g], cplx);
      info = check_problem(&p, -1);
    } else {
      info = 0;
    }
    if (info >= 0) {
      char name[16];
      size_t len = strlen(batch_error_name[dtype]) + 1;
      memcpy(name, batch_error_name[dtype], len);
      xerbla_(name, &info, (blasint)len);
      free(probs);
      return;
    }
    if (p.m == 0 || p.n == 0) continue;
    for (blasint j = 0; j < group_size[g]; j++) {
      b200_problem *q = &probs[count++];
      *q = p;
      if (order == CblasColMajor) { q->a = a_arr[idx + j]; q->b = b_arr[idx + j]; }
      else                        { q->a = b_arr[idx + j]; q->b = a_arr[idx + j]; }
      q->c = c_arr[idx + j];
    }
  }
  if (count > 0) {
    int err = b200_run_batch(probs, count);
    if (err) b200_fatal(batch_error_name[dtype], err);
  }
  free(probs);
}

#define CBLAS_BATCH(NAME, DT, TSCAL, TIN, TOUT)                                               \
  B200_EXPORT void NAME(enum CBLAS_ORDER order, const enum CBLAS_TRANSPOSE *ta,               \
                        const enum CBLAS_TRANSPOSE *tb, const blasint *m, const blasint *n,   \
                        const blasint *k, const TSCAL *alpha, const TIN **a,                  \
                        const blasint *lda, const TIN **b, const blasint *ldb,                \
                        const TSCAL *beta, TOUT **c, const blasint *ldc, blasint group_count, \
                        const blasint *group_size) {                                          \
    run_batch(DT, (int)order, ta, tb, m, n, k, alpha, (const void *const *)a, lda,            \
              (const void *const *)b, ldb, beta, (void *const *)c, ldc, group_count,          \
              group_size);                                                                    \
  }

CBLAS_BATCH(cblas_sgemm_batch, B200_S, float, float, float)
CBLAS_BATCH(cblas_dgemm_batch, B200_D, double, double, double)
CBLAS_BATCH(cblas_cgemm_batch, B200_C, void, void, void)
CBLAS_BATCH(cblas_zgemm_batch, B200_Z, void, void, void)
CBLAS_BATCH(cblas_sbgemm_batch, B200_SB, float, bfloat16, float)

/* ------------------------------------------------------------ bf16 conversion helpers
 * interface/tobf16.c, interface/bf16to.c: n <= 0 returns; negative increments walk the
 * array backwards from its far end. */
static void convert(int dir, int64_t n, const void *in, int64_t inc_in, size_t in_sz, void *out,
                    int64_t inc_out, size_t out_sz) {
  if (n <= 0) return;
  if (inc_in < 0) in = (const char *)in - (n - 1) * inc_in * (int64_t)in_sz;
  if (inc_out < 0) out = (char *)out - (n - 1) * inc_out * (int64_t)out_sz;
  int err = b200_run_convert(dir, n, in, inc_in, out, inc_out);
  if (err) b200_fatal("bf16 convert", err);
}

B200_EXPORT void cblas_sbstobf16(blasint n, const float *in, blasint incin, bfloat16 *out,
                                 blasint incout) { convert(0, n, in, incin, 4, out, incout, 2); }
B200_EXPORT void cblas_sbdtobf16(blasint n, const double *in, blasint incin, bfloat16 *out,
                                 blasint incout) { convert(1, n, in, incin, 8, out, incout, 2); }
B200_EXPORT void cblas_sbf16tos(blasint n, const bfloat16 *in, blasint incin, float *out,
                                blasint incout) { convert(2, n, in, incin, 2, out, incout, 4); }
B200_EXPORT void cblas_dbf16tod(blasint n, const bfloat16 *in, blasint incin, double *out,
                                blasint incout) { convert(3, n, in, incin, 2, out, incout, 8); }
B200_EXPORT void sbstobf16_(blasint *n, float *in, blasint *incin, bfloat16 *out,
                            blasint *incout) { convert(0, *n, in, *incin, 4, out, *incout, 2); }
B200_EXPORT void sbdtobf16_(blasint *n, double *in, blasint *incin, bfloat16 *out,
                            blasint *incout) { convert(1, *n, in, *incin, 8, out, *incout, 2); }
B200_EXPORT void sbf16tos_(blasint *n, bfloat16 *in, blasint *incin, float *out,
                           blasint *incout) { convert(2, *n, in, *incin, 2, out, *incout, 4); }
B200_EXPORT void dbf16tod_(blasint *n, bfloat16 *in, blasint *incin, double *out,
                           blasint *incout) { convert(3, *n, in, *incin, 2, out, *incout, 8); }

/* ------------------------------------------------------------------- SBGEMV / SBDOT
 * interface/sbgemv.c: flag decoding :76-79 (N, T, R = N, C = T) / :121-136 (row-major flips trans and swaps m, n);
 * checks :81-86 / :138-143 in this order, xerbla_("SBGEMV ", &info, 8); m == 0 or n == 0 returns (:164); alpha == 0
 * only scales y (:171-174); negative increments start from the far end (:179-180).
 * interface/bf16dot.c: n <= 0 returns 0 (:14, :34). */
static void sbgemv_entry(int cblas, int trans, int64_t m, int64_t n, float alpha, const bfloat16 *a, int64_t lda,
                         const bfloat16 *x, int64_t incx, float beta, float *y, int64_t incy) {
  blasint info = cblas ? -1 : 0;
  char nm[9] = "SBGEMV ";
  if (incy == 0) info = 11;
  if (incx == 0) info = 8;
  if (lda < (m > 1 ? m : 1)) info = 6;
  if (n < 0) info = 3;
  if (m < 0) info = 2;
  if (trans < 0) info = 1;
  if (cblas ? info >= 0 : info != 0) { xerbla_(nm, &info, 8); return; }
  if (m == 0 || n == 0) return;
  const int64_t lenx = trans ? m : n, leny = trans ? n : m;
  if (incx < 0) x -= (lenx - 1) * incx;
  if (incy < 0) y -= (leny - 1) * incy;
  int err = b200_run_sbgemv(trans, m, n, alpha, a, lda, x, incx, beta, y, incy);
  if (err) b200_fatal("SBGEMV", err);
}

B200_EXPORT void sbgemv_(char *TRANS, blasint *M, blasint *N, float *ALPHA, bfloat16 *a, blasint *LDA, bfloat16 *x,
                         blasint *INCX, float *BETA, float *y, blasint *INCY) {
  char t = *TRANS;
  if (t >= 'a' && t <= 'z') t = (char)(t - 'a' + 'A');
  const int trans = (t == 'N' || t == 'R') ? 0 : (t == 'T' || t == 'C') ? 1 : -1;
  sbgemv_entry(0, trans, *M, *N, *ALPHA, a, *LDA, x, *INCX, *BETA, y, *INCY);
}

B200_EXPORT void cblas_sbgemv(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE TransA, blasint m, blasint n, float alpha,
                              const bfloat16 *a, blasint lda, const bfloat16 *x, blasint incx, float beta, float *y,
                              blasint incy) {
  int trans = -1;
  const int nt = TransA == CblasNoTrans || TransA == CblasConjNoTrans, tr = TransA == CblasTrans || TransA == CblasConjTrans;
  if (order == CblasColMajor) {
    trans = nt ? 0 : tr ? 1 : -1;
  } else {                      /* the reference treats every other order value as row-major (sbgemv.c:127) */
    trans = nt ? 1 : tr ? 0 : -1;
    blasint t = n; n = m; m = t;
  }
  sbgemv_entry(1, trans, m, n, alpha, a, lda, x, incx, beta, y, incy);
}

static float sbdot_entry(int64_t n, const bfloat16 *x, int64_t incx, const bfloat16 *y, int64_t incy) {
  if (n <= 0) return 0.f;
  if (incx < 0) x -= (n - 1) * incx;
  if (incy < 0) y -= (n - 1) * incy;
  float r = 0.f;
  int err = b200_run_sbdot(n, x, incx, y, incy, &r);
  if (err) b200_fatal("SBDOT", err);
  return r;
}
B200_EXPORT float sbdot_(blasint *N, bfloat16 *x, blasint *INCX, bfloat16 *y, blasint *INCY) { return sbdot_entry(*N, x, *INCX, y, *INCY); }
B200_EXPORT float cblas_sbdot(blasint n, const bfloat16 *x, blasint incx, const bfloat16 *y, blasint incy) { return sbdot_entry(n, x, incx, y, incy); }
