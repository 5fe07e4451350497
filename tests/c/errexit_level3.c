/*
 * errexit_level3.c -- prints what xerbla_ receives for a table of illegal SYMM/HEMM, SYRK/HERK,
 * SYR2K/HER2K and TRMM/TRSM calls (CBLAS in both orders, Fortran), plus the quick returns.  No GPU needed: every
 * call must return before the library touches CUDA.  The SAME program is linked once against the
 * reference (oracle/_ref/generic/libopenblas_ref.so -> tests/golden/errexit_level3_reference.txt,
 * written by tests/golden/make_golden.py) and once against libopenblas_b200.so; the two outputs
 * must be identical line for line (tests/test_abi.py).  It supplies its own xerbla_, which must
 * override the library's, like ctest/c_xerbla.c.
 */
#include <stdio.h>
#include <string.h>
#include "openblas_b200.h"

static int calls, last_info;
static char last_name[16];

int xerbla_(char *name, blasint *info, blasint len) {
  calls++;
  last_info = *info;
  memset(last_name, 0, sizeof last_name);
  strncpy(last_name, name, len < 15 ? len : 15);
  return 0;
}
static void show(const char *what) {
  if (calls) printf("%-34s calls=%d name='%s' info=%d\n", what, calls, last_name, last_info);
  else printf("%-34s no error\n", what);
  calls = 0; last_info = -99; last_name[0] = 0;
}

#define COL CblasColMajor
#define ROW CblasRowMajor
#define L_ CblasLeft
#define R_ CblasRight
#define U_ CblasUpper
#define LO CblasLower
#define N_ CblasNoTrans
#define T_ CblasTrans
#define C_ CblasConjTrans
#define BADS ((enum CBLAS_SIDE)0)
#define BADU ((enum CBLAS_UPLO)0)
#define BADT ((enum CBLAS_TRANSPOSE)0)
#define NU CblasNonUnit
#define UN CblasUnit
#define BADD ((enum CBLAS_DIAG)0)

int main(void) {
  double a[16] = {0}, b[16] = {0}, c[16];
  float fa[16] = {0}, fb[16] = {0}, fc[16];
  double zal[2] = {1, 0}, zbe[2] = {1, 0};   /* beta == 1 everywhere: a call that is legal in one order is then a no-op */
  float cal[2] = {1, 0}, cbe[2] = {1, 0};
  for (int i = 0; i < 16; i++) { c[i] = 42.0; fc[i] = 42.0f; }
  enum CBLAS_ORDER orders[3] = {COL, ROW, (enum CBLAS_ORDER)0};
  const char *oname[3] = {"col", "row", "bad-order"};
  char what[64];

  for (int o = 0; o < 3; o++) {
    enum CBLAS_ORDER ord = orders[o];
#define W(s) (snprintf(what, sizeof what, "%s %s", oname[o], s), what)
    /* SYMM: side, uplo, m, n, lda (left: vs m; right: vs n), ldb, ldc */
    cblas_dsymm(ord, BADS, U_, 0, 0, 1.0, a, 1, b, 1, 1.0, c, 1); show(W("dsymm side"));
    cblas_dsymm(ord, L_, BADU, 0, 0, 1.0, a, 1, b, 1, 1.0, c, 1); show(W("dsymm uplo"));
    cblas_dsymm(ord, L_, U_, -1, 0, 1.0, a, 1, b, 1, 1.0, c, 1); show(W("dsymm m<0"));
    cblas_dsymm(ord, R_, LO, 0, -1, 1.0, a, 1, b, 1, 1.0, c, 1); show(W("dsymm n<0"));
    cblas_dsymm(ord, L_, U_, 2, 0, 1.0, a, 1, b, 2, 1.0, c, 2); show(W("dsymm lda left"));
    cblas_dsymm(ord, R_, U_, 0, 2, 1.0, a, 1, b, 2, 1.0, c, 2); show(W("dsymm lda right"));
    cblas_dsymm(ord, L_, LO, 2, 0, 1.0, a, 2, b, 1, 1.0, c, 2); show(W("dsymm ldb left m=2"));
    cblas_dsymm(ord, R_, LO, 0, 2, 1.0, a, 2, b, 1, 1.0, c, 2); show(W("dsymm ldb right n=2"));
    cblas_dsymm(ord, L_, U_, 2, 0, 1.0, a, 2, b, 2, 1.0, c, 1); show(W("dsymm ldc m=2"));
    cblas_dsymm(ord, R_, U_, 0, 2, 1.0, a, 2, b, 2, 1.0, c, 1); show(W("dsymm ldc n=2"));
    cblas_dsymm(ord, BADS, U_, 2, 3, 1.0, a, 2, b, 2, 1.0, c, 1); show(W("dsymm side + others"));
    cblas_ssymm(ord, L_, U_, 0, -1, 1.f, fa, 1, fb, 1, 1.f, fc, 1); show(W("ssymm n<0"));
    cblas_csymm(ord, L_, BADU, 0, 0, cal, fa, 1, fb, 1, cbe, fc, 1); show(W("csymm uplo"));
    cblas_zsymm(ord, R_, U_, 0, 2, zal, a, 1, b, 2, zbe, c, 2); show(W("zsymm lda right"));
    cblas_chemm(ord, L_, U_, 2, 0, cal, fa, 1, fb, 2, cbe, fc, 2); show(W("chemm lda left"));
    cblas_zhemm(ord, R_, LO, -1, 0, zal, a, 1, b, 1, zbe, c, 1); show(W("zhemm m<0"));
    /* SYRK / HERK: uplo, trans, n, k, lda (N: vs n, T: vs k), ldc */
    cblas_dsyrk(ord, BADU, N_, 0, 0, 1.0, a, 1, 1.0, c, 1); show(W("dsyrk uplo"));
    cblas_dsyrk(ord, U_, BADT, 0, 0, 1.0, a, 1, 1.0, c, 1); show(W("dsyrk trans"));
    cblas_dsyrk(ord, U_, C_, 0, 0, 1.0, a, 1, 1.0, c, 1); show(W("dsyrk trans C (real: legal)"));
    cblas_dsyrk(ord, U_, N_, -1, 0, 1.0, a, 1, 1.0, c, 1); show(W("dsyrk n<0"));
    cblas_dsyrk(ord, LO, T_, 0, -1, 1.0, a, 1, 1.0, c, 1); show(W("dsyrk k<0"));
    cblas_dsyrk(ord, U_, N_, 2, 0, 1.0, a, 1, 1.0, c, 2); show(W("dsyrk lda N n=2"));
    cblas_dsyrk(ord, U_, T_, 0, 2, 1.0, a, 1, 1.0, c, 1); show(W("dsyrk lda T k=2"));
    cblas_dsyrk(ord, LO, N_, 2, 0, 1.0, a, 2, 1.0, c, 1); show(W("dsyrk ldc"));
    cblas_ssyrk(ord, U_, N_, 0, -1, 1.f, fa, 1, 1.f, fc, 1); show(W("ssyrk k<0"));
    cblas_csyrk(ord, U_, C_, 0, 0, cal, fa, 1, cbe, fc, 1); show(W("csyrk trans C (illegal)"));
    cblas_zsyrk(ord, U_, T_, 0, 2, zal, a, 1, zbe, c, 1); show(W("zsyrk lda T"));
    cblas_cherk(ord, U_, T_, 0, 0, 1.f, fa, 1, 1.f, fc, 1); show(W("cherk trans T (illegal)"));
    cblas_zherk(ord, LO, C_, 0, 2, 1.0, a, 1, 1.0, c, 1); show(W("zherk lda C"));
    cblas_zherk(ord, LO, N_, 2, 0, 1.0, a, 2, 1.0, c, 1); show(W("zherk ldc"));
    /* SYR2K / HER2K: + ldb, ldc moves to 12 */
    cblas_dsyr2k(ord, BADU, N_, 0, 0, 1.0, a, 1, b, 1, 1.0, c, 1); show(W("dsyr2k uplo"));
    cblas_dsyr2k(ord, U_, BADT, 0, 0, 1.0, a, 1, b, 1, 1.0, c, 1); show(W("dsyr2k trans"));
    cblas_dsyr2k(ord, U_, N_, -1, 0, 1.0, a, 1, b, 1, 1.0, c, 1); show(W("dsyr2k n<0"));
    cblas_dsyr2k(ord, U_, N_, 0, -1, 1.0, a, 1, b, 1, 1.0, c, 1); show(W("dsyr2k k<0"));
    cblas_dsyr2k(ord, U_, N_, 2, 0, 1.0, a, 1, b, 2, 1.0, c, 2); show(W("dsyr2k lda"));
    cblas_dsyr2k(ord, U_, T_, 0, 2, 1.0, a, 2, b, 1, 1.0, c, 1); show(W("dsyr2k ldb T"));
    cblas_dsyr2k(ord, LO, N_, 2, 0, 1.0, a, 2, b, 2, 1.0, c, 1); show(W("dsyr2k ldc"));
    cblas_ssyr2k(ord, LO, T_, 0, 2, 1.f, fa, 1, fb, 2, 1.f, fc, 1); show(W("ssyr2k lda T"));
    cblas_csyr2k(ord, LO, C_, 0, 0, cal, fa, 1, fb, 1, cbe, fc, 1); show(W("csyr2k trans C (illegal)"));
    cblas_zsyr2k(ord, LO, N_, 2, 0, zal, a, 2, b, 1, zbe, c, 2); show(W("zsyr2k ldb N"));
    cblas_cher2k(ord, LO, T_, 0, 0, cal, fa, 1, fb, 1, 1.f, fc, 1); show(W("cher2k trans T (illegal)"));
    cblas_zher2k(ord, U_, C_, 0, 2, zal, a, 1, b, 2, 1.0, c, 1); show(W("zher2k lda C"));
    /* TRMM / TRSM: side, uplo, trans, diag, m, n, lda (left: vs m; right: vs n), ldb */
    cblas_dtrmm(ord, BADS, U_, N_, NU, 0, 0, 1.0, a, 1, c, 1); show(W("dtrmm side"));
    cblas_dtrmm(ord, L_, BADU, N_, NU, 0, 0, 1.0, a, 1, c, 1); show(W("dtrmm uplo"));
    cblas_dtrmm(ord, L_, U_, BADT, NU, 0, 0, 1.0, a, 1, c, 1); show(W("dtrmm trans"));
    cblas_dtrmm(ord, L_, U_, N_, BADD, 0, 0, 1.0, a, 1, c, 1); show(W("dtrmm diag"));
    cblas_dtrmm(ord, L_, U_, N_, NU, -1, 0, 1.0, a, 1, c, 1); show(W("dtrmm m<0"));
    cblas_dtrmm(ord, R_, LO, T_, UN, 0, -1, 1.0, a, 1, c, 1); show(W("dtrmm n<0"));
    cblas_dtrmm(ord, L_, U_, N_, NU, 2, 0, 1.0, a, 1, c, 2); show(W("dtrmm lda left m=2"));
    cblas_dtrmm(ord, R_, U_, N_, NU, 0, 2, 1.0, a, 1, c, 2); show(W("dtrmm lda right n=2"));
    cblas_dtrmm(ord, L_, LO, C_, NU, 2, 0, 1.0, a, 2, c, 1); show(W("dtrmm ldb m=2"));
    cblas_dtrmm(ord, R_, LO, C_, NU, 0, 2, 1.0, a, 2, c, 1); show(W("dtrmm ldb n=2"));
    cblas_dtrmm(ord, BADS, U_, N_, NU, 2, 3, 1.0, a, 1, c, 1); show(W("dtrmm side + others"));
    cblas_strmm(ord, L_, U_, N_, BADD, 0, 0, 1.f, fa, 1, fc, 1); show(W("strmm diag"));
    cblas_ctrmm(ord, L_, U_, CblasConjNoTrans, NU, 2, 0, cal, fa, 1, fc, 2); show(W("ctrmm lda (R legal)"));
    cblas_ztrmm(ord, R_, U_, C_, UN, 0, -1, zal, a, 1, c, 1); show(W("ztrmm n<0"));
    cblas_dtrsm(ord, BADS, U_, N_, NU, 0, 0, 1.0, a, 1, c, 1); show(W("dtrsm side"));
    cblas_dtrsm(ord, L_, U_, BADT, NU, 0, 0, 1.0, a, 1, c, 1); show(W("dtrsm trans"));
    cblas_dtrsm(ord, L_, U_, N_, NU, 2, 0, 1.0, a, 1, c, 2); show(W("dtrsm lda left m=2"));
    cblas_dtrsm(ord, R_, U_, N_, NU, 0, 2, 1.0, a, 1, c, 2); show(W("dtrsm lda right n=2"));
    cblas_dtrsm(ord, L_, LO, T_, UN, 2, 0, 1.0, a, 2, c, 1); show(W("dtrsm ldb m=2"));
    cblas_strsm(ord, L_, BADU, N_, NU, 0, 0, 1.f, fa, 1, fc, 1); show(W("strsm uplo"));
    cblas_ctrsm(ord, R_, U_, N_, NU, -1, 0, cal, fa, 1, fc, 1); show(W("ctrsm m<0"));
    cblas_ztrsm(ord, L_, U_, N_, BADD, 0, 0, zal, a, 1, c, 1); show(W("ztrsm diag"));
    cblas_dtrsm(ord, L_, U_, N_, NU, 0, 3, 1.0, a, 1, c, 1); show(W("dtrsm m == 0"));
    cblas_dtrmm(ord, R_, U_, N_, NU, 3, 0, 1.0, a, 1, c, 3); show(W("dtrmm n == 0"));
    /* quick returns */
    cblas_dsymm(ord, L_, U_, 0, 3, 1.0, a, 1, b, 1, 1.0, c, 1); show(W("dsymm m == 0"));
    cblas_dsymm(ord, R_, U_, 3, 0, 1.0, a, 1, b, 3, 1.0, c, 3); show(W("dsymm n == 0"));
    cblas_dsyrk(ord, U_, N_, 0, 3, 1.0, a, 1, 1.0, c, 1); show(W("dsyrk n == 0"));
    cblas_zher2k(ord, U_, N_, 0, 3, zal, a, 1, b, 1, 1.0, c, 1); show(W("zher2k n == 0"));
  }

  /* Fortran ABI: flags by character, case-insensitive */
  { blasint z = 0, m1 = -1, one = 1, two = 2; double al = 1, be = 1; float fal = 1, fbe = 1;
    char l = 'l', r = 'R', u = 'u', lo = 'L', n = 'n', t = 'T', cc = 'c', x = 'X';
    dsymm_(&x, &u, &z, &z, &al, a, &one, b, &one, &be, c, &one); show("f77 dsymm side");
    dsymm_(&l, &x, &z, &z, &al, a, &one, b, &one, &be, c, &one); show("f77 dsymm uplo");
    dsymm_(&l, &u, &m1, &z, &al, a, &one, b, &one, &be, c, &one); show("f77 dsymm m<0");
    dsymm_(&r, &lo, &z, &two, &al, a, &one, b, &one, &be, c, &one); show("f77 dsymm lda right");
    dsymm_(&l, &lo, &two, &z, &al, a, &two, b, &one, &be, c, &two); show("f77 dsymm ldb");
    dsymm_(&l, &lo, &two, &z, &al, a, &two, b, &two, &be, c, &one); show("f77 dsymm ldc");
    zhemm_(&r, &u, &z, &m1, zal, a, &one, b, &one, zbe, c, &one); show("f77 zhemm n<0");
    dsyrk_(&u, &cc, &z, &z, &al, a, &one, &be, c, &one); show("f77 dsyrk trans c (real: legal)");
    dsyrk_(&u, &x, &z, &z, &al, a, &one, &be, c, &one); show("f77 dsyrk trans");
    dsyrk_(&lo, &t, &z, &two, &al, a, &one, &be, c, &one); show("f77 dsyrk lda T");
    ssyrk_(&x, &n, &z, &z, &fal, fa, &one, &fbe, fc, &one); show("f77 ssyrk uplo");
    zsyrk_(&u, &cc, &z, &z, zal, a, &one, zbe, c, &one); show("f77 zsyrk trans c (illegal)");
    cherk_(&u, &t, &z, &z, &fal, fa, &one, &fbe, fc, &one); show("f77 cherk trans t (illegal)");
    zherk_(&u, &cc, &two, &z, &al, a, &one, &be, c, &one); show("f77 zherk ldc");
    dsyr2k_(&u, &n, &two, &z, &al, a, &two, b, &one, &be, c, &two); show("f77 dsyr2k ldb");
    dsyr2k_(&u, &n, &two, &z, &al, a, &two, b, &two, &be, c, &one); show("f77 dsyr2k ldc");
    zher2k_(&lo, &t, &z, &z, zal, a, &one, b, &one, &be, c, &one); show("f77 zher2k trans t (illegal)");
    cher2k_(&lo, &cc, &z, &m1, cal, fa, &one, fb, &one, &fbe, fc, &one); show("f77 cher2k k<0");
    dsyrk_(&u, &n, &z, &two, &al, a, &one, &be, c, &one); show("f77 dsyrk n == 0");
    /* trsm.c:207 hands xerbla_ sizeof(ERROR_NAME)-1: the name arrives without its NUL */
    dtrmm_(&x, &u, &n, &n, &z, &z, &al, a, &one, c, &one); show("f77 dtrmm side");
    dtrmm_(&l, &u, &x, &n, &z, &z, &al, a, &one, c, &one); show("f77 dtrmm trans");
    dtrmm_(&l, &u, &r, &u, &z, &z, &al, a, &one, c, &one); show("f77 dtrmm trans r (legal), diag u");
    dtrmm_(&l, &u, &n, &x, &z, &z, &al, a, &one, c, &one); show("f77 dtrmm diag");
    dtrsm_(&r, &lo, &cc, &n, &z, &two, &al, a, &one, c, &one); show("f77 dtrsm lda right");
    dtrsm_(&l, &lo, &t, &u, &two, &z, &al, a, &two, c, &one); show("f77 dtrsm ldb");
    ztrsm_(&l, &u, &n, &n, &m1, &z, zal, a, &one, c, &one); show("f77 ztrsm m<0");
    ctrmm_(&l, &x, &n, &n, &z, &z, cal, fa, &one, fc, &one); show("f77 ctrmm uplo");
  }
  for (int i = 0; i < 16; i++)
    if (c[i] != 42.0 || fc[i] != 42.0f) { printf("C was written at %d\n", i); break; }
  return 0;
}
