/*
 * errexit.c -- C-ABI error-exit harness (no GPU needed: every call here must return before
 * the library touches CUDA).  It is this repository's analogue of the reference's
 * ctest/c_d3chke.c + ctest/c_xerbla.c: the program supplies its OWN xerbla_, which must
 * override the library's weak default, and checks for each illegal call
 *   - that xerbla_ was called exactly once, with the right routine name and info
 *   - that C was not written
 * for the CBLAS symbols in both orders (info is evaluated after the row-major swap,
 * interface/gemm.c:423-471) and for the Fortran symbols.  Also checks the quick returns
 * (m == 0 or n == 0: nothing touched, no xerbla, interface/gemm.c:494).
 */
#include <stdio.h>
#include <string.h>
#include "openblas_b200.h"

static int calls, last_info, failures;
static char last_name[16];

int xerbla_(char *name, blasint *info, blasint len) {
  calls++;
  last_info = *info;
  memset(last_name, 0, sizeof last_name);
  strncpy(last_name, name, len < 15 ? len : 15);
  return 0;
}

static void expect(const char *what, const char *name, int info) {
  if (calls != 1 || last_info != info || strncmp(last_name, name, strlen(name)) != 0) {
    printf("FAIL %-40s calls=%d name='%s' info=%d (want '%s' %d)\n", what, calls, last_name, last_info,
           name, info);
    failures++;
  }
  calls = 0;
}
static void expect_none(const char *what) {
  if (calls != 0) { printf("FAIL %-40s unexpected xerbla '%s' %d\n", what, last_name, last_info); failures++; }
  calls = 0;
}

#define N CblasNoTrans
#define T CblasTrans
#define COL CblasColMajor
#define ROW CblasRowMajor
#define BAD ((enum CBLAS_TRANSPOSE)0)

int main(void) {
  double a[8] = {0}, b[8] = {0}, c[8];
  float fa[8] = {0}, fb[8] = {0}, fc[8];
  double zal[2] = {1, 0}, zbe[2] = {0, 0};
  bfloat16 ha[8] = {0}, hb[8] = {0};
  for (int i = 0; i < 8; i++) { c[i] = 42.0; fc[i] = 42.0f; }

  /* invalid order: info 0 (interface/gemm.c:370-373,473) */
  cblas_dgemm((enum CBLAS_ORDER)0, N, N, 0, 0, 0, 1.0, a, 1, b, 1, 0.0, c, 1); expect("bad order", "DGEMM ", 0);

  /* column-major: the table of c_d3chke.c:58-160 */
  cblas_dgemm(COL, BAD, N, 0, 0, 0, 1.0, a, 1, b, 1, 0.0, c, 1); expect("col transa", "DGEMM ", 1);
  cblas_dgemm(COL, N, BAD, 0, 0, 0, 1.0, a, 1, b, 1, 0.0, c, 1); expect("col transb", "DGEMM ", 2);
  cblas_dgemm(COL, N, N, -1, 0, 0, 1.0, a, 1, b, 1, 0.0, c, 1); expect("col m<0", "DGEMM ", 3);
  cblas_dgemm(COL, T, T, 0, -1, 0, 1.0, a, 1, b, 1, 0.0, c, 1); expect("col n<0", "DGEMM ", 4);
  cblas_dgemm(COL, N, T, 0, 0, -1, 1.0, a, 1, b, 1, 0.0, c, 1); expect("col k<0", "DGEMM ", 5);
  cblas_dgemm(COL, N, N, 2, 0, 0, 1.0, a, 1, b, 1, 0.0, c, 2); expect("col lda NN", "DGEMM ", 8);
  cblas_dgemm(COL, T, N, 0, 0, 2, 1.0, a, 1, b, 2, 0.0, c, 1); expect("col lda TN", "DGEMM ", 8);
  cblas_dgemm(COL, N, N, 0, 0, 2, 1.0, a, 1, b, 1, 0.0, c, 1); expect("col ldb NN", "DGEMM ", 10);
  cblas_dgemm(COL, N, T, 0, 2, 0, 1.0, a, 1, b, 1, 0.0, c, 1); expect("col ldb NT", "DGEMM ", 10);
  cblas_dgemm(COL, N, N, 2, 0, 0, 1.0, a, 2, b, 1, 0.0, c, 1); expect("col ldc", "DGEMM ", 13);
  cblas_dgemm(COL, T, T, 2, 0, 0, 1.0, a, 1, b, 1, 0.0, c, 1); expect("col ldc TT", "DGEMM ", 13);

  /* row-major: checks run on the swapped problem, so m<0 reports 4, n<0 reports 3, lda is
   * reported as 10 and ldb as 8 (c_xerbla.c:36-42 maps them back for the ctest harness) */
  cblas_dgemm(ROW, BAD, N, 0, 0, 0, 1.0, a, 1, b, 1, 0.0, c, 1); expect("row transa", "DGEMM ", 2);
  cblas_dgemm(ROW, N, BAD, 0, 0, 0, 1.0, a, 1, b, 1, 0.0, c, 1); expect("row transb", "DGEMM ", 1);
  cblas_dgemm(ROW, N, N, -1, 0, 0, 1.0, a, 1, b, 1, 0.0, c, 1); expect("row m<0", "DGEMM ", 4);
  cblas_dgemm(ROW, N, N, 0, -1, 0, 1.0, a, 1, b, 1, 0.0, c, 1); expect("row n<0", "DGEMM ", 3);
  cblas_dgemm(ROW, N, N, 0, 0, -1, 1.0, a, 1, b, 1, 0.0, c, 1); expect("row k<0", "DGEMM ", 5);
  cblas_dgemm(ROW, N, N, 0, 0, 2, 1.0, a, 1, b, 1, 0.0, c, 1); expect("row lda NN", "DGEMM ", 10);
  cblas_dgemm(ROW, N, N, 0, 2, 0, 1.0, a, 1, b, 1, 0.0, c, 2); expect("row ldb NN", "DGEMM ", 8);
  cblas_dgemm(ROW, N, N, 0, 2, 0, 1.0, a, 1, b, 2, 0.0, c, 1); expect("row ldc", "DGEMM ", 13);

  /* the other precisions share the code; one probe each incl. the 7/8-char names */
  cblas_sgemm(COL, N, N, -1, 0, 0, 1.f, fa, 1, fb, 1, 0.f, fc, 1); expect("sgemm", "SGEMM ", 3);
  cblas_sbgemm(COL, N, N, 0, -1, 0, 1.f, ha, 1, hb, 1, 0.f, fc, 1); expect("sbgemm", "SBGEMM ", 4);
  cblas_cgemm(COL, CblasConjTrans, BAD, 0, 0, 0, zal, fa, 1, fb, 1, zbe, fc, 1); expect("cgemm", "CGEMM ", 2);
  cblas_zgemm(COL, CblasConjNoTrans, N, 2, 0, 0, zal, a, 1, b, 1, zbe, c, 2); expect("zgemm", "ZGEMM ", 8);
  cblas_zgemm3m(COL, N, N, 0, 0, -1, zal, a, 1, b, 1, zbe, c, 1); expect("zgemm3m", "ZGEMM3M ", 5);
  cblas_cgemm3m(COL, N, N, 0, 0, -1, zal, fa, 1, fb, 1, zbe, fc, 1); expect("cgemm3m", "CGEMM3M ", 5);

  /* Fortran ABI: 1-based info, trans by character, case-insensitive, R/C accepted */
  { blasint m = -1, z = 0, one = 1, two = 2; double al = 1, be = 0; char n = 'n', t = 'T', x = 'X', r = 'r', cc = 'c';
    dgemm_(&n, &n, &m, &z, &z, &al, a, &one, b, &one, &be, c, &one); expect("f77 m<0", "DGEMM ", 3);
    dgemm_(&x, &n, &z, &z, &z, &al, a, &one, b, &one, &be, c, &one); expect("f77 transa", "DGEMM ", 1);
    dgemm_(&t, &x, &z, &z, &z, &al, a, &one, b, &one, &be, c, &one); expect("f77 transb", "DGEMM ", 2);
    dgemm_(&r, &cc, &z, &z, &z, &al, a, &one, b, &one, &be, c, &one); expect_none("f77 R/C accepted for real");
    dgemm_(&n, &n, &two, &z, &z, &al, a, &one, b, &one, &be, c, &two); expect("f77 lda", "DGEMM ", 8);
    dgemm_(&n, &n, &two, &z, &z, &al, a, &two, b, &one, &be, c, &one); expect("f77 ldc", "DGEMM ", 13);
    float fal = 1, fbe = 0;
    sbgemm_(&n, &n, &z, &m, &z, &fal, ha, &one, hb, &one, &fbe, fc, &one); expect("f77 sbgemm", "SBGEMM ", 4);
    zgemm_(&r, &x, &z, &z, &z, zal, a, &one, b, &one, zbe, c, &one); expect("f77 zgemm", "ZGEMM ", 2);
  }

  /* batch: a bad group aborts the whole call (interface/gemm_batch.c:283-287) */
  { enum CBLAS_TRANSPOSE ta[1] = {N}, tb[1] = {N}; blasint mm[1] = {-1}, zz[1] = {0}, ld[1] = {1}, gs[1] = {1};
    double al[1] = {1}, be[1] = {0}; const double *ap[1] = {a}, *bp[1] = {b}; double *cp[1] = {c};
    cblas_dgemm_batch(COL, ta, tb, mm, zz, zz, al, ap, ld, bp, ld, be, cp, ld, 1, gs); expect("batch m<0", "DGEMM_BATCH ", 3);
    cblas_dgemm_batch((enum CBLAS_ORDER)7, ta, tb, zz, zz, zz, al, ap, ld, bp, ld, be, cp, ld, 1, gs); expect("batch order", "DGEMM_BATCH ", 0);
    cblas_dgemm_batch(COL, ta, tb, zz, zz, zz, al, ap, ld, bp, ld, be, cp, ld, 1, gs); expect_none("batch empty group");
  }

  /* quick returns: legal, m == 0 or n == 0 -> no xerbla, nothing touched, no CUDA */
  cblas_dgemm(COL, N, N, 0, 3, 3, 1.0, a, 1, b, 3, 0.0, c, 1); expect_none("m == 0");
  cblas_dgemm(COL, N, N, 3, 0, 3, 1.0, a, 3, b, 3, 0.0, c, 3); expect_none("n == 0");
  cblas_dgemm(ROW, T, N, 0, 2, 2, 1.0, a, 1, b, 2, 0.0, c, 2); expect_none("row m == 0");
  { blasint z = 0, one = 1; cblas_sbstobf16(z, fa, one, ha, one); expect_none("tobf16 n == 0"); }

  for (int i = 0; i < 8; i++)
    if (c[i] != 42.0 || fc[i] != 42.0f) { printf("FAIL C was written at %d\n", i); failures++; break; }

  printf(failures ? "ERROR EXITS: %d FAILURES\n" : "ERROR EXITS PASSED (%d failures)\n", failures);
  return failures != 0;
}
