/*
 * interface_level3.c -- BLAS / CBLAS entry points of the rest of level 3, host side, plain C:
 * SYMM / HEMM, SYRK / HERK, SYR2K / HER2K and TRMM / TRSM for s, d, c, z.
 *
 * Replaces the reference's interface/symm.c, interface/syrk.c, interface/syr2k.c and interface/trsm.c (each
 * compiled there once per precision, ABI and -DHEMM).  One table-driven implementation serves
 * all of them; an entry point only decodes its flag arguments and packs the rest.  What is
 * restated from the reference:
 *
 *   - flag decoding: side L/R, uplo U/L, trans per routine (real: N, T, C=T; complex SYRK/SYR2K:
 *     N, T; HERK/HER2K: N, C)                     (symm.c:184-190, syrk.c:131-152, 254-268)
 *   - row-major = the transposed column-major problem: side and uplo flip and m <-> n for
 *     SYMM/HEMM (symm.c:328-367); uplo and trans flip for the SYRK family (syrk.c:288-312),
 *     HER2K additionally conjugates alpha (syr2k.c:305-311)
 *   - argument checks in the reference's order, lowest-numbered failing argument reported
 *     (symm.c:203-227, syrk.c:158-165, syr2k.c:280-288); an unset side takes the "right"
 *     branch of the leading-dimension checks exactly as `if (!side) ... else ...` does
 *   - xerbla_(NAME, &info, sizeof(NAME)) then return; CBLAS fires on info >= 0 with -1 = ok,
 *     an illegal order leaves info = 0 (symm.c:369-372)
 *   - quick returns: m == 0 or n == 0 (symm.c:376); n == 0 (syrk.c:344)
 *
 * The normalised problem goes to the CUDA side (b200_run_level3): the symmetric operand is
 * expanded once on the device and multiplied by the GEMM kernels; the rank-k updates are cut
 * into block columns whose rectangles are plain GEMMs into C and whose diagonal blocks are
 * merged under the triangle mask.  There is no CPU compute path in this file or behind it.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "shim.h"

extern int xerbla_(char *name, blasint *info, blasint len);

static void report(const char *name6, blasint info) {
  char name[8];
  memcpy(name, name6, 7);                 /* six characters + NUL: sizeof(ERROR_NAME) == 7 */
  xerbla_(name, &info, 7);
}

static int64_t max1(int64_t x) { return x > 1 ? x : 1; }

static char upper(char ch) { return (ch >= 'a' && ch <= 'z') ? (char)(ch - 'a' + 'A') : ch; }
static int side_of_char(char ch) { ch = upper(ch); return ch == 'L' ? 0 : ch == 'R' ? 1 : -1; }
static int uplo_of_char(char ch) { ch = upper(ch); return ch == 'U' ? 0 : ch == 'L' ? 1 : -1; }
static int side_of_cblas(int s) { return s == CblasLeft ? 0 : s == CblasRight ? 1 : -1; }
static int uplo_of_cblas(int u) { return u == CblasUpper ? 0 : u == CblasLower ? 1 : -1; }

/* kind: 0 = real type, 1 = complex symmetric routine (SYRK/SYR2K), 2 = Hermitian routine */
static int trans_of_char(char ch, int kind) {
  ch = upper(ch);
  if (ch == 'N') return 0;
  if (kind == 0) return (ch == 'T' || ch == 'C') ? 1 : -1;
  if (kind == 1) return ch == 'T' ? 1 : -1;
  return ch == 'C' ? 1 : -1;
}
static int trans_of_cblas(int t, int kind) {
  if (t == CblasNoTrans) return 0;
  if (kind == 0) return t == CblasTrans ? 1 : t == CblasConjNoTrans ? 0 : t == CblasConjTrans ? 1 : -1;
  if (kind == 1) return t == CblasTrans ? 1 : -1;
  return t == CblasConjTrans ? 1 : -1;
}
static int flip(int v) { return v < 0 ? v : !v; }

static void run(const b200_l3_problem *p, const char *where) {
  int err = b200_run_level3(p);
  if (err) b200_fatal(where, err);
}

/* ------------------------------------------------------------------ SYMM / HEMM ---- */
static void symm_entry(const char *name, int routine, int dtype, int cblas, int order, int side, int uplo,
                       int64_t m, int64_t n, const double alpha[2], const void *a, int64_t lda, const void *b,
                       int64_t ldb, const double beta[2], void *c, int64_t ldc) {
  blasint info = cblas ? -1 : 0;
  if (cblas && order == CblasRowMajor) {
    side = flip(side); uplo = flip(uplo);
    int64_t t = m; m = n; n = t;
  } else if (cblas && order != CblasColMajor) {
    report(name, 0);
    return;
  }
  if (ldc < max1(m)) info = 12;
  if (side == 0) {
    if (ldb < max1(m)) info = 9;
    if (lda < max1(m)) info = 7;
  } else {
    if (ldb < max1(m)) info = 9;
    if (lda < max1(n)) info = 7;
  }
  if (n < 0) info = 4;
  if (m < 0) info = 3;
  if (uplo < 0) info = 2;
  if (side < 0) info = 1;
  if (cblas ? info >= 0 : info != 0) { report(name, info); return; }
  if (m == 0 || n == 0) return;

  b200_l3_problem p;
  memset(&p, 0, sizeof p);
  p.routine = routine; p.dtype = dtype; p.side = side; p.uplo = uplo;
  p.m = m; p.n = n; p.lda = lda; p.ldb = ldb; p.ldc = ldc;
  p.alpha[0] = alpha[0]; p.alpha[1] = alpha[1]; p.beta[0] = beta[0]; p.beta[1] = beta[1];
  p.a = a; p.b = b; p.c = c;
  run(&p, name);
}

/* ------------------------------------------------- SYRK / HERK / SYR2K / HER2K ---- */
static void rankk_entry(const char *name, int routine, int dtype, int cblas, int order, int uplo, int trans,
                        int64_t n, int64_t k, const double alpha_in[2], const void *a, int64_t lda, const void *b,
                        int64_t ldb, const double beta[2], void *c, int64_t ldc) {
  const int two = routine == B200_SYR2K || routine == B200_HER2K;
  double alpha[2] = {alpha_in[0], alpha_in[1]};
  blasint info = cblas ? -1 : 0;
  if (cblas && order == CblasRowMajor) {
    uplo = flip(uplo); trans = flip(trans);
    if (routine == B200_HER2K) alpha[1] = -alpha[1];
  } else if (cblas && order != CblasColMajor) {
    report(name, 0);
    return;
  }
  const int64_t nrowa = (trans & 1) ? k : n;
  if (ldc < max1(n)) info = two ? 12 : 10;
  if (two && ldb < max1(nrowa)) info = 9;
  if (lda < max1(nrowa)) info = 7;
  if (k < 0) info = 4;
  if (n < 0) info = 3;
  if (trans < 0) info = 2;
  if (uplo < 0) info = 1;
  if (cblas ? info >= 0 : info != 0) { report(name, info); return; }
  if (n == 0) return;

  b200_l3_problem p;
  memset(&p, 0, sizeof p);
  p.routine = routine; p.dtype = dtype; p.uplo = uplo; p.trans = trans;
  p.n = n; p.m = n; p.k = k; p.lda = lda; p.ldb = two ? ldb : lda; p.ldc = ldc;
  p.alpha[0] = alpha[0]; p.alpha[1] = alpha[1]; p.beta[0] = beta[0]; p.beta[1] = beta[1];
  p.a = a; p.b = two ? b : a; p.c = c;
  run(&p, name);
}

/* --------------------------------------------------------------------- GEMMT ---- */
/* interface/gemmt.c: the uplo triangle of the m x m matrix C := alpha op(A) op(B) + beta C.  Flag decoding :124-161
 * (Fortran) / :232-262 (CBLAS); row-major = the column-major problem with A and B swapped and uplo flipped (:305-345);
 * checks :163-186 column-major, :365-389 row-major -- the row-major branch reports positions 10 / 8 for the leading
 * dimensions it calls lda / ldb (the caller's LDB / LDA; the 10 test comes last) and 3 / 2 for the trans flags;
 * xerbla_ gets sizeof("?GEMMT ") = 8; quick return m == 0 (:464).  SBGEMMT (interface/sbgemmt.c) comes through here too.  The reference then walks the triangle column by
 * column with GEMV (and conjugates its "b" operand IN PLACE for transb = R / C, :466-476, never undoing it -- not
 * reproduced: inputs are const here); here it is ONE triangle-masked launch of the GEMM kernel. */
static void gemmt_entry(const char *name, int dtype, int cblas, int order, int uplo, int transa, int transb, int64_t m,
                        int64_t k, const double alpha[2], const void *a, int64_t lda, const void *b, int64_t ldb,
                        const double beta[2], void *c, int64_t ldc) {
  blasint info = cblas ? -1 : 0;
  const int sb = dtype == B200_SB;
  const blasint name_len = (blasint)strlen(name) + 1;      /* sizeof(ERROR_NAME): 8, or 9 for "SBGEMMT " */
  char nm[10];
  memcpy(nm, name, (size_t)name_len);
  if (cblas && order == CblasRowMajor) {
    const void *t = a; a = b; b = t;
    int64_t tl = lda; lda = ldb; ldb = tl;
    int tt = transa; transa = transb; transb = tt;
    /* interface/sbgemmt.c:239-240 keeps uplo as given (gemmt.c:322-323 flips it), so a row-major SBGEMMT updates the
     * other triangle of the caller's C than a row-major ?GEMMT would: reproduced, callers of the reference see this */
    if (!sb) uplo = flip(uplo);
    const int64_t ncola = (transa >= 0 && (transa & 1)) ? k : m, ncolb = (transb >= 0 && (transb & 1)) ? m : k;
    if (ldc < max1(m)) info = 13;
    if (ldb < max1(ncolb)) info = 8;
    if (lda < max1(ncola)) info = 10;
    if (k < 0) info = 5;
    if (m < 0) info = 4;
    if (transb < 0) info = 2;
    if (transa < 0) info = 3;
    if (uplo < 0) info = 1;
  } else if (cblas && order != CblasColMajor) {
    info = 0;
  } else {
    const int64_t nrowa = (transa >= 0 && (transa & 1)) ? k : m, nrowb = (transb >= 0 && (transb & 1)) ? m : k;
    if (ldc < max1(m)) info = 13;
    if (ldb < max1(nrowb)) info = 10;
    if (lda < max1(nrowa)) info = 8;
    if (k < 0) info = 5;
    if (m < 0) info = 4;
    if (transb < 0) info = 3;
    if (transa < 0) info = 2;
    if (uplo < 0) info = 1;
  }
  if (cblas ? info >= 0 : info != 0) { xerbla_(nm, &info, name_len); return; }
  if (m == 0) return;
  /* SBGEMMT hands every column to the SBGEMV kernel, which returns at once for an empty dimension
   * (kernel/x86_64/sbgemv_n.c:114): k == 0 leaves C as it is, whatever beta says */
  if (sb && k == 0) return;

  b200_l3_problem p;
  memset(&p, 0, sizeof p);
  p.routine = B200_GEMMT; p.dtype = dtype; p.uplo = uplo; p.trans = transa; p.transb = transb;
  p.n = m; p.m = m; p.k = k; p.lda = lda; p.ldb = ldb; p.ldc = ldc;
  p.alpha[0] = alpha[0]; p.alpha[1] = alpha[1]; p.beta[0] = beta[0]; p.beta[1] = beta[1];
  p.a = a; p.b = b; p.c = c;
  run(&p, name);
}

/* ---------------------------------------------------------------- TRMM / TRSM ---- */
static int diag_of_char(char ch) { ch = upper(ch); return ch == 'U' ? 1 : ch == 'N' ? 0 : -1; }   /* 1 = unit */
static int diag_of_cblas(int d) { return d == CblasUnit ? 1 : d == CblasNonUnit ? 0 : -1; }
static int op_of_char(char ch, int cplx) {        /* trsm.c:172-175: N, T, R, C for every type; real types fold R->N, C->T */
  ch = upper(ch);
  return ch == 'N' ? B200_N : ch == 'T' ? B200_T : ch == 'R' ? (cplx ? B200_R : B200_N) : ch == 'C' ? (cplx ? B200_C_ : B200_T) : -1;
}
static int op_of_cblas(int t, int cplx) {         /* trsm.c:283-291 */
  return t == CblasNoTrans ? B200_N : t == CblasTrans ? B200_T : t == CblasConjNoTrans ? (cplx ? B200_R : B200_N)
       : t == CblasConjTrans ? (cplx ? B200_C_ : B200_T) : -1;
}

/* interface/trsm.c: the checks (:188-197, :296-305); row-major swaps m and n and flips side and uplo
 * but not trans (:308-333); the Fortran entry hands xerbla_ sizeof(ERROR_NAME)-1 = 6 (:207), the CBLAS
 * one sizeof(ERROR_NAME) = 7 (:354); quick return m == 0 or n == 0 (:359). */
static void trxm_entry(const char *name, int routine, int dtype, int cblas, int order, int side, int uplo, int trans, int unit,
                       int64_t m, int64_t n, const double alpha[2], const void *a, int64_t lda, void *b, int64_t ldb) {
  blasint info = cblas ? -1 : 0;
  if (cblas && order == CblasRowMajor) {
    side = flip(side); uplo = flip(uplo);
    int64_t t = m; m = n; n = t;
  } else if (cblas && order != CblasColMajor) {
    report(name, 0);
    return;
  }
  const int64_t nrowa = (side & 1) ? n : m;
  if (ldb < max1(m)) info = 11;
  if (lda < max1(nrowa)) info = 9;
  if (n < 0) info = 6;
  if (m < 0) info = 5;
  if (unit < 0) info = 4;
  if (trans < 0) info = 3;
  if (uplo < 0) info = 2;
  if (side < 0) info = 1;
  if (cblas ? info >= 0 : info != 0) {
    char nm[8];
    memcpy(nm, name, 7);
    xerbla_(nm, &info, cblas ? 7 : 6);
    return;
  }
  if (m == 0 || n == 0) return;

  b200_l3_problem p;
  memset(&p, 0, sizeof p);
  p.routine = routine; p.dtype = dtype; p.side = side; p.uplo = uplo; p.trans = trans; p.unit = unit;
  p.m = m; p.n = n; p.lda = lda; p.ldb = ldb; p.ldc = ldb;
  p.alpha[0] = alpha[0]; p.alpha[1] = alpha[1]; p.beta[0] = 0.0; p.beta[1] = 0.0;
  p.a = a; p.b = NULL; p.c = b;
  run(&p, name);
}

/* ---------------------------------------------------------------- entry points ---- */
#define REAL2(x)  {(double)(x), 0.0}
#define CPLX2(T, p) {(double)((const T *)(p))[0], (double)((const T *)(p))[1]}

/* SYMM / HEMM.  SCALAR_C / SCALAR_F turn the scalar arguments into double[2]. */
#define DEF_SYMM(P, NAME, ROUTINE, DTYPE, T, CS, CSCAL, FSCAL)                                                     \
  B200_EXPORT void cblas_##P(enum CBLAS_ORDER Order, enum CBLAS_SIDE Side, enum CBLAS_UPLO Uplo, blasint M, blasint N, \
                             CS alpha, const T *A, blasint lda, const T *B, blasint ldb, CS beta, T *C, blasint ldc) { \
    const double al[2] = CSCAL(alpha), be[2] = CSCAL(beta);                                                         \
    symm_entry(NAME, ROUTINE, DTYPE, 1, (int)Order, side_of_cblas((int)Side), uplo_of_cblas((int)Uplo), M, N, al, A, \
               lda, B, ldb, be, C, ldc);                                                                            \
  }                                                                                                                 \
  B200_EXPORT void P##_(char *SIDE, char *UPLO, blasint *M, blasint *N, FSCAL *alpha, FSCAL *a, blasint *ldA,        \
                        FSCAL *b, blasint *ldB, FSCAL *beta, FSCAL *c, blasint *ldC) {                               \
    const double al[2] = F_##CSCAL(alpha), be[2] = F_##CSCAL(beta);                                                 \
    symm_entry(NAME, ROUTINE, DTYPE, 0, 0, side_of_char(*SIDE), uplo_of_char(*UPLO), *M, *N, al, a, *ldA, b, *ldB,  \
               be, c, *ldC);                                                                                        \
  }
#define SC_REAL(x) REAL2(x)
#define F_SC_REAL(p) REAL2(*(p))
#define SC_CF(p) CPLX2(float, p)
#define F_SC_CF(p) CPLX2(float, p)
#define SC_CD(p) CPLX2(double, p)
#define F_SC_CD(p) CPLX2(double, p)

DEF_SYMM(ssymm, "SSYMM ", B200_SYMM, B200_S, float, float, SC_REAL, float)
DEF_SYMM(dsymm, "DSYMM ", B200_SYMM, B200_D, double, double, SC_REAL, double)
DEF_SYMM(csymm, "CSYMM ", B200_SYMM, B200_C, void, const void *, SC_CF, float)
DEF_SYMM(zsymm, "ZSYMM ", B200_SYMM, B200_Z, void, const void *, SC_CD, double)
DEF_SYMM(chemm, "CHEMM ", B200_HEMM, B200_C, void, const void *, SC_CF, float)
DEF_SYMM(zhemm, "ZHEMM ", B200_HEMM, B200_Z, void, const void *, SC_CD, double)

/* SYRK / HERK: KIND selects the legal trans values; ASCAL/BSCAL the scalar conventions */
#define DEF_SYRK(P, NAME, ROUTINE, DTYPE, KIND, T, CSA, CSB, ASCAL, BSCAL, FSCAL)                                    \
  B200_EXPORT void cblas_##P(enum CBLAS_ORDER Order, enum CBLAS_UPLO Uplo, enum CBLAS_TRANSPOSE Trans, blasint N,    \
                             blasint K, CSA alpha, const T *A, blasint lda, CSB beta, T *C, blasint ldc) {           \
    const double al[2] = ASCAL(alpha), be[2] = BSCAL(beta);                                                         \
    rankk_entry(NAME, ROUTINE, DTYPE, 1, (int)Order, uplo_of_cblas((int)Uplo), trans_of_cblas((int)Trans, KIND), N,  \
                K, al, A, lda, NULL, 0, be, C, ldc);                                                                \
  }                                                                                                                 \
  B200_EXPORT void P##_(char *UPLO, char *TRANS, blasint *N, blasint *K, FSCAL *alpha, FSCAL *a, blasint *ldA,       \
                        FSCAL *beta, FSCAL *c, blasint *ldC) {                                                       \
    const double al[2] = F_##ASCAL(alpha), be[2] = F_##BSCAL(beta);                                                 \
    rankk_entry(NAME, ROUTINE, DTYPE, 0, 0, uplo_of_char(*UPLO), trans_of_char(*TRANS, KIND), *N, *K, al, a, *ldA,  \
                NULL, 0, be, c, *ldC);                                                                              \
  }
DEF_SYRK(ssyrk, "SSYRK ", B200_SYRK, B200_S, 0, float, float, float, SC_REAL, SC_REAL, float)
DEF_SYRK(dsyrk, "DSYRK ", B200_SYRK, B200_D, 0, double, double, double, SC_REAL, SC_REAL, double)
DEF_SYRK(csyrk, "CSYRK ", B200_SYRK, B200_C, 1, void, const void *, const void *, SC_CF, SC_CF, float)
DEF_SYRK(zsyrk, "ZSYRK ", B200_SYRK, B200_Z, 1, void, const void *, const void *, SC_CD, SC_CD, double)
DEF_SYRK(cherk, "CHERK ", B200_HERK, B200_C, 2, void, float, float, SC_REAL, SC_REAL, float)
DEF_SYRK(zherk, "ZHERK ", B200_HERK, B200_Z, 2, void, double, double, SC_REAL, SC_REAL, double)

#define DEF_SYR2K(P, NAME, ROUTINE, DTYPE, KIND, T, CSA, CSB, ASCAL, BSCAL, FSCAL)                                   \
  B200_EXPORT void cblas_##P(enum CBLAS_ORDER Order, enum CBLAS_UPLO Uplo, enum CBLAS_TRANSPOSE Trans, blasint N,    \
                             blasint K, CSA alpha, const T *A, blasint lda, const T *B, blasint ldb, CSB beta, T *C, \
                             blasint ldc) {                                                                         \
    const double al[2] = ASCAL(alpha), be[2] = BSCAL(beta);                                                         \
    rankk_entry(NAME, ROUTINE, DTYPE, 1, (int)Order, uplo_of_cblas((int)Uplo), trans_of_cblas((int)Trans, KIND), N,  \
                K, al, A, lda, B, ldb, be, C, ldc);                                                                 \
  }                                                                                                                 \
  B200_EXPORT void P##_(char *UPLO, char *TRANS, blasint *N, blasint *K, FSCAL *alpha, FSCAL *a, blasint *ldA,       \
                        FSCAL *b, blasint *ldB, FSCAL *beta, FSCAL *c, blasint *ldC) {                               \
    const double al[2] = F_##ASCAL(alpha), be[2] = F_##BSCAL(beta);                                                 \
    rankk_entry(NAME, ROUTINE, DTYPE, 0, 0, uplo_of_char(*UPLO), trans_of_char(*TRANS, KIND), *N, *K, al, a, *ldA,  \
                b, *ldB, be, c, *ldC);                                                                              \
  }
DEF_SYR2K(ssyr2k, "SSYR2K", B200_SYR2K, B200_S, 0, float, float, float, SC_REAL, SC_REAL, float)
DEF_SYR2K(dsyr2k, "DSYR2K", B200_SYR2K, B200_D, 0, double, double, double, SC_REAL, SC_REAL, double)
DEF_SYR2K(csyr2k, "CSYR2K", B200_SYR2K, B200_C, 1, void, const void *, const void *, SC_CF, SC_CF, float)
DEF_SYR2K(zsyr2k, "ZSYR2K", B200_SYR2K, B200_Z, 1, void, const void *, const void *, SC_CD, SC_CD, double)
DEF_SYR2K(cher2k, "CHER2K", B200_HER2K, B200_C, 2, void, const void *, float, SC_CF, SC_REAL, float)
DEF_SYR2K(zher2k, "ZHER2K", B200_HER2K, B200_Z, 2, void, const void *, double, SC_CD, SC_REAL, double)

#define DEF_TRXM(P, NAME, ROUTINE, DTYPE, CPLX, T, CS, CSCAL, FSCAL)                                                  \
  B200_EXPORT void cblas_##P(enum CBLAS_ORDER Order, enum CBLAS_SIDE Side, enum CBLAS_UPLO Uplo,                      \
                             enum CBLAS_TRANSPOSE TransA, enum CBLAS_DIAG Diag, blasint M, blasint N, CS alpha,       \
                             const T *A, blasint lda, T *B, blasint ldb) {                                            \
    const double al[2] = CSCAL(alpha);                                                                               \
    trxm_entry(NAME, ROUTINE, DTYPE, 1, (int)Order, side_of_cblas((int)Side), uplo_of_cblas((int)Uplo),              \
               op_of_cblas((int)TransA, CPLX), diag_of_cblas((int)Diag), M, N, al, A, lda, B, ldb);                  \
  }                                                                                                                  \
  B200_EXPORT void P##_(char *SIDE, char *UPLO, char *TRANSA, char *DIAG, blasint *M, blasint *N, FSCAL *alpha,       \
                        FSCAL *a, blasint *ldA, FSCAL *b, blasint *ldB) {                                             \
    const double al[2] = F_##CSCAL(alpha);                                                                           \
    trxm_entry(NAME, ROUTINE, DTYPE, 0, 0, side_of_char(*SIDE), uplo_of_char(*UPLO), op_of_char(*TRANSA, CPLX),      \
               diag_of_char(*DIAG), *M, *N, al, a, *ldA, b, *ldB);                                                   \
  }
DEF_TRXM(strmm, "STRMM ", B200_TRMM, B200_S, 0, float, float, SC_REAL, float)
DEF_TRXM(dtrmm, "DTRMM ", B200_TRMM, B200_D, 0, double, double, SC_REAL, double)
DEF_TRXM(ctrmm, "CTRMM ", B200_TRMM, B200_C, 1, void, const void *, SC_CF, float)
DEF_TRXM(ztrmm, "ZTRMM ", B200_TRMM, B200_Z, 1, void, const void *, SC_CD, double)
DEF_TRXM(strsm, "STRSM ", B200_TRSM, B200_S, 0, float, float, SC_REAL, float)
DEF_TRXM(dtrsm, "DTRSM ", B200_TRSM, B200_D, 0, double, double, SC_REAL, double)
DEF_TRXM(ctrsm, "CTRSM ", B200_TRSM, B200_C, 1, void, const void *, SC_CF, float)
DEF_TRXM(ztrsm, "ZTRSM ", B200_TRSM, B200_Z, 1, void, const void *, SC_CD, double)

/* GEMMT: op decoding as GEMM's (N, T, R, C; real types fold R -> N, C -> T) */
#define DEF_GEMMT(P, NAME, DTYPE, CPLX, T, CS, CSCAL, FSCAL)                                                          \
  B200_EXPORT void cblas_##P(enum CBLAS_ORDER Order, enum CBLAS_UPLO Uplo, enum CBLAS_TRANSPOSE TransA,                \
                             enum CBLAS_TRANSPOSE TransB, blasint M, blasint K, CS alpha, const T *A, blasint lda,     \
                             const T *B, blasint ldb, CS beta, T *C, blasint ldc) {                                    \
    const double al[2] = CSCAL(alpha), be[2] = CSCAL(beta);                                                           \
    gemmt_entry(NAME, DTYPE, 1, (int)Order, uplo_of_cblas((int)Uplo), op_of_cblas((int)TransA, CPLX),                  \
                op_of_cblas((int)TransB, CPLX), M, K, al, A, lda, B, ldb, be, C, ldc);                                 \
  }                                                                                                                   \
  B200_EXPORT void P##_(char *UPLO, char *TRANSA, char *TRANSB, blasint *M, blasint *K, FSCAL *alpha, FSCAL *a,        \
                        blasint *ldA, FSCAL *b, blasint *ldB, FSCAL *beta, FSCAL *c, blasint *ldC) {                   \
    const double al[2] = F_##CSCAL(alpha), be[2] = F_##CSCAL(beta);                                                   \
    gemmt_entry(NAME, DTYPE, 0, 0, uplo_of_char(*UPLO), op_of_char(*TRANSA, CPLX), op_of_char(*TRANSB, CPLX), *M, *K,  \
                al, a, *ldA, b, *ldB, be, c, *ldC);                                                                    \
  }
DEF_GEMMT(sgemmt, "SGEMMT ", B200_S, 0, float, float, SC_REAL, float)
DEF_GEMMT(dgemmt, "DGEMMT ", B200_D, 0, double, double, SC_REAL, double)
DEF_GEMMT(cgemmt, "CGEMMT ", B200_C, 1, void, const void *, SC_CF, float)
DEF_GEMMT(zgemmt, "ZGEMMT ", B200_Z, 1, void, const void *, SC_CD, double)

/* SBGEMMT (interface/sbgemmt.c; built with BUILD_BFLOAT16, interface/Makefile:52,290; declared in no public header
 * of the reference): bf16 A and B, fp32 alpha, beta and C; checks and info numbers as ?GEMMT (:121-137, :286-301),
 * xerbla_ gets sizeof("SBGEMMT ") = 9 */
B200_EXPORT void cblas_sbgemmt(enum CBLAS_ORDER Order, enum CBLAS_UPLO Uplo, enum CBLAS_TRANSPOSE TransA,
                               enum CBLAS_TRANSPOSE TransB, blasint M, blasint K, float alpha, const bfloat16 *A, blasint lda,
                               const bfloat16 *B, blasint ldb, float beta, float *C, blasint ldc) {
  const double al[2] = {alpha, 0.0}, be[2] = {beta, 0.0};
  gemmt_entry("SBGEMMT ", B200_SB, 1, (int)Order, uplo_of_cblas((int)Uplo), op_of_cblas((int)TransA, 0), op_of_cblas((int)TransB, 0),
              M, K, al, A, lda, B, ldb, be, C, ldc);
}
B200_EXPORT void sbgemmt_(char *UPLO, char *TRANSA, char *TRANSB, blasint *M, blasint *K, float *alpha, bfloat16 *a, blasint *ldA,
                          bfloat16 *b, blasint *ldB, float *beta, float *c, blasint *ldC) {
  const double al[2] = {*alpha, 0.0}, be[2] = {*beta, 0.0};
  gemmt_entry("SBGEMMT ", B200_SB, 0, 0, uplo_of_char(*UPLO), op_of_char(*TRANSA, 0), op_of_char(*TRANSB, 0), *M, *K, al, a, *ldA,
              b, *ldB, be, c, *ldC);
}
