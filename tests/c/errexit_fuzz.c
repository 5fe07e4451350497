/*
 * errexit_fuzz.c -- randomized companion of errexit.c / errexit_level3.c.  A fixed-seed generator
 * draws 4000 calls over every exported level-3 entry point (CBLAS column-major, row-major, illegal
 * order, Fortran ABI) with flags, extents and leading dimensions taken from small sets that include
 * illegal values, and prints what xerbla_ received.  Every call is a no-op when it is legal (alpha = 0
 * and beta = 1 for GEMM / SYMM / rank-k updates; an empty matrix for TRMM / TRSM), so neither the
 * reference nor this library computes anything and no GPU is needed.  Linked against the reference it
 * wrote tests/golden/errexit_fuzz_reference.txt (tests/golden/make_level3_golden.py); linked against
 * libopenblas_b200.so it must print the same bytes (tests/test_abi.py).
 */
#include <stdio.h>
#include <stdlib.h>
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
static unsigned long long state = 0x9E3779B97F4A7C15ull;
static unsigned rnd(unsigned n) {
  state = state * 6364136223846793005ull + 1442695040888963407ull;
  return (unsigned)((state >> 33) % n);
}
static int pick(const int *set, int n) { return set[rnd((unsigned)n)]; }

int main(int argc, char **argv) {
  static double a[64], b[64], c[64];
  const int iterations = argc > 1 ? atoi(argv[1]) : 4000;          /* defaults = the committed fixture */
  if (argc > 2) state ^= strtoull(argv[2], NULL, 0) * 0xD6E8FEB86659FD93ull;
  static const int orders[] = {CblasColMajor, CblasRowMajor, CblasColMajor, CblasRowMajor, 0};
  static const int transes[] = {CblasNoTrans, CblasTrans, CblasConjTrans, CblasConjNoTrans, 0, 7};
  static const int sides[] = {CblasLeft, CblasRight, CblasLeft, CblasRight, 0};
  static const int uplos[] = {CblasUpper, CblasLower, CblasUpper, CblasLower, 3};
  static const int diags[] = {CblasNonUnit, CblasUnit, CblasNonUnit, CblasUnit, 9};
  static const int dims[] = {-1, 0, 1, 2, 3, 5};
  static const int lds[] = {0, 1, 2, 3, 5, 6};
  static const char tch[] = {'N', 'T', 'C', 'R', 'n', 'c', 'X', 'q'};
  static const char sch[] = {'L', 'R', 'l', 'r', 'X'};
  static const char uch[] = {'U', 'L', 'u', 'l', 'Z'};
  static const char dch[] = {'N', 'U', 'n', 'u', 'Y'};
  double z0[2] = {0, 0}, z1[2] = {1, 0};
  float c0[2] = {0, 0}, c1[2] = {1, 0};
  double d0 = 0, d1 = 1; float s0 = 0, s1 = 1;
  for (int i = 0; i < 64; i++) c[i] = 42.0;

  for (int it = 0; it < iterations; it++) {
    const int fam = (int)rnd(10), prec = (int)rnd(4), f77 = rnd(4) == 0;
    const enum CBLAS_ORDER o = (enum CBLAS_ORDER)pick(orders, 5);
    const enum CBLAS_TRANSPOSE ta = (enum CBLAS_TRANSPOSE)pick(transes, 6), tb = (enum CBLAS_TRANSPOSE)pick(transes, 6);
    const enum CBLAS_SIDE sd = (enum CBLAS_SIDE)pick(sides, 5);
    const enum CBLAS_UPLO up = (enum CBLAS_UPLO)pick(uplos, 5);
    const enum CBLAS_DIAG dg = (enum CBLAS_DIAG)pick(diags, 5);
    blasint m = pick(dims, 6), n = pick(dims, 6), k = pick(dims, 6), lda = pick(lds, 6), ldb = pick(lds, 6), ldc = pick(lds, 6);
    char cta = tch[rnd(8)], ctb = tch[rnd(8)], cs = sch[rnd(5)], cu = uch[rnd(5)], cd = dch[rnd(5)];
    const char *what = "?";
    if ((fam == 7 || fam == 8) && m > 0 && n > 0) { if (rnd(2)) m = 0; else n = 0; }     /* TRMM / TRSM: legal calls must be empty */
    /* scalars: alpha = 0 (real or complex), beta = 1 */
#define AL (prec == 0 ? (void *)&s0 : prec == 1 ? (void *)&d0 : prec == 2 ? (void *)c0 : (void *)z0)
#define BE (prec == 0 ? (void *)&s1 : prec == 1 ? (void *)&d1 : prec == 2 ? (void *)c1 : (void *)z1)
    switch (fam) {
      case 0:   /* GEMM */
        what = "gemm";
        if (f77) {
          if (prec == 0) sgemm_(&cta, &ctb, &m, &n, &k, &s0, (float *)a, &lda, (float *)b, &ldb, &s1, (float *)c, &ldc);
          else if (prec == 1) dgemm_(&cta, &ctb, &m, &n, &k, &d0, a, &lda, b, &ldb, &d1, c, &ldc);
          else if (prec == 2) cgemm_(&cta, &ctb, &m, &n, &k, c0, (float *)a, &lda, (float *)b, &ldb, c1, (float *)c, &ldc);
          else zgemm_(&cta, &ctb, &m, &n, &k, z0, a, &lda, b, &ldb, z1, c, &ldc);
        } else {
          if (prec == 0) cblas_sgemm(o, ta, tb, m, n, k, 0.f, (float *)a, lda, (float *)b, ldb, 1.f, (float *)c, ldc);
          else if (prec == 1) cblas_dgemm(o, ta, tb, m, n, k, 0.0, a, lda, b, ldb, 1.0, c, ldc);
          else if (prec == 2) cblas_cgemm(o, ta, tb, m, n, k, c0, a, lda, b, ldb, c1, c, ldc);
          else cblas_zgemm(o, ta, tb, m, n, k, z0, a, lda, b, ldb, z1, c, ldc);
        }
        break;
      case 1:   /* SBGEMM / GEMM3M */
        what = "sbgemm/gemm3m";
        if (prec < 2) cblas_sbgemm(o, ta, tb, m, n, k, 0.f, (bfloat16 *)a, lda, (bfloat16 *)b, ldb, 1.f, (float *)c, ldc);
        else if (prec == 2) cblas_cgemm3m(o, ta, tb, m, n, k, c0, a, lda, b, ldb, c1, c, ldc);
        else cblas_zgemm3m(o, ta, tb, m, n, k, z0, a, lda, b, ldb, z1, c, ldc);
        break;
      case 2:   /* SYMM */
        what = "symm";
        if (f77) {
          if (prec == 0) ssymm_(&cs, &cu, &m, &n, &s0, (float *)a, &lda, (float *)b, &ldb, &s1, (float *)c, &ldc);
          else if (prec == 1) dsymm_(&cs, &cu, &m, &n, &d0, a, &lda, b, &ldb, &d1, c, &ldc);
          else if (prec == 2) csymm_(&cs, &cu, &m, &n, c0, (float *)a, &lda, (float *)b, &ldb, c1, (float *)c, &ldc);
          else zsymm_(&cs, &cu, &m, &n, z0, a, &lda, b, &ldb, z1, c, &ldc);
        } else {
          if (prec == 0) cblas_ssymm(o, sd, up, m, n, 0.f, (float *)a, lda, (float *)b, ldb, 1.f, (float *)c, ldc);
          else if (prec == 1) cblas_dsymm(o, sd, up, m, n, 0.0, a, lda, b, ldb, 1.0, c, ldc);
          else if (prec == 2) cblas_csymm(o, sd, up, m, n, c0, a, lda, b, ldb, c1, c, ldc);
          else cblas_zsymm(o, sd, up, m, n, z0, a, lda, b, ldb, z1, c, ldc);
        }
        break;
      case 3:   /* HEMM */
        what = "hemm";
        if (f77) { if (prec & 1) zhemm_(&cs, &cu, &m, &n, z0, a, &lda, b, &ldb, z1, c, &ldc); else chemm_(&cs, &cu, &m, &n, c0, (float *)a, &lda, (float *)b, &ldb, c1, (float *)c, &ldc); }
        else { if (prec & 1) cblas_zhemm(o, sd, up, m, n, z0, a, lda, b, ldb, z1, c, ldc); else cblas_chemm(o, sd, up, m, n, c0, a, lda, b, ldb, c1, c, ldc); }
        break;
      case 4:   /* SYRK / HERK */
        what = "syrk/herk";
        if (f77) {
          if (prec == 0) ssyrk_(&cu, &cta, &n, &k, &s0, (float *)a, &lda, &s1, (float *)c, &ldc);
          else if (prec == 1) dsyrk_(&cu, &cta, &n, &k, &d0, a, &lda, &d1, c, &ldc);
          else if (prec == 2) { if (rnd(2)) csyrk_(&cu, &cta, &n, &k, c0, (float *)a, &lda, c1, (float *)c, &ldc); else cherk_(&cu, &cta, &n, &k, &s0, (float *)a, &lda, &s1, (float *)c, &ldc); }
          else { if (rnd(2)) zsyrk_(&cu, &cta, &n, &k, z0, a, &lda, z1, c, &ldc); else zherk_(&cu, &cta, &n, &k, &d0, a, &lda, &d1, c, &ldc); }
        } else {
          if (prec == 0) cblas_ssyrk(o, up, ta, n, k, 0.f, (float *)a, lda, 1.f, (float *)c, ldc);
          else if (prec == 1) cblas_dsyrk(o, up, ta, n, k, 0.0, a, lda, 1.0, c, ldc);
          else if (prec == 2) { if (rnd(2)) cblas_csyrk(o, up, ta, n, k, c0, a, lda, c1, c, ldc); else cblas_cherk(o, up, ta, n, k, 0.f, a, lda, 1.f, c, ldc); }
          else { if (rnd(2)) cblas_zsyrk(o, up, ta, n, k, z0, a, lda, z1, c, ldc); else cblas_zherk(o, up, ta, n, k, 0.0, a, lda, 1.0, c, ldc); }
        }
        break;
      case 5: case 6:   /* SYR2K / HER2K */
        what = "syr2k/her2k";
        if (f77) {
          if (prec == 0) ssyr2k_(&cu, &cta, &n, &k, &s0, (float *)a, &lda, (float *)b, &ldb, &s1, (float *)c, &ldc);
          else if (prec == 1) dsyr2k_(&cu, &cta, &n, &k, &d0, a, &lda, b, &ldb, &d1, c, &ldc);
          else if (prec == 2) { if (fam == 5) csyr2k_(&cu, &cta, &n, &k, c0, (float *)a, &lda, (float *)b, &ldb, c1, (float *)c, &ldc); else cher2k_(&cu, &cta, &n, &k, c0, (float *)a, &lda, (float *)b, &ldb, &s1, (float *)c, &ldc); }
          else { if (fam == 5) zsyr2k_(&cu, &cta, &n, &k, z0, a, &lda, b, &ldb, z1, c, &ldc); else zher2k_(&cu, &cta, &n, &k, z0, a, &lda, b, &ldb, &d1, c, &ldc); }
        } else {
          if (prec == 0) cblas_ssyr2k(o, up, ta, n, k, 0.f, (float *)a, lda, (float *)b, ldb, 1.f, (float *)c, ldc);
          else if (prec == 1) cblas_dsyr2k(o, up, ta, n, k, 0.0, a, lda, b, ldb, 1.0, c, ldc);
          else if (prec == 2) { if (fam == 5) cblas_csyr2k(o, up, ta, n, k, c0, a, lda, b, ldb, c1, c, ldc); else cblas_cher2k(o, up, ta, n, k, c0, a, lda, b, ldb, 1.f, c, ldc); }
          else { if (fam == 5) cblas_zsyr2k(o, up, ta, n, k, z0, a, lda, b, ldb, z1, c, ldc); else cblas_zher2k(o, up, ta, n, k, z0, a, lda, b, ldb, 1.0, c, ldc); }
        }
        break;
      case 9: { /* GEMM_BATCH: up to two groups, the second one drawn independently; a bad group aborts the call */
        what = "gemm_batch";
        enum CBLAS_TRANSPOSE tas[2] = {ta, (enum CBLAS_TRANSPOSE)pick(transes, 6)}, tbs[2] = {tb, (enum CBLAS_TRANSPOSE)pick(transes, 6)};
        blasint ms[2] = {m, pick(dims, 6)}, ns[2] = {n, pick(dims, 6)}, ks[2] = {k, pick(dims, 6)};
        blasint las[2] = {lda, pick(lds, 6)}, lbs[2] = {ldb, pick(lds, 6)}, lcs[2] = {ldc, pick(lds, 6)};
        blasint gs[2] = {(blasint)rnd(3), (blasint)rnd(3)}, gc = (blasint)rnd(3);
        const void *ap[4] = {a, a, a, a}, *bp[4] = {b, b, b, b}; void *cp[4] = {c, c, c, c};
        if (prec == 0) { float al[2] = {0, 0}, be[2] = {1, 1}; cblas_sgemm_batch(o, tas, tbs, ms, ns, ks, al, (const float **)ap, las, (const float **)bp, lbs, be, (float **)cp, lcs, gc, gs); }
        else if (prec == 1) { double al[2] = {0, 0}, be[2] = {1, 1}; cblas_dgemm_batch(o, tas, tbs, ms, ns, ks, al, (const double **)ap, las, (const double **)bp, lbs, be, (double **)cp, lcs, gc, gs); }
        else if (prec == 2) { float al[4] = {0, 0, 0, 0}, be[4] = {1, 0, 1, 0}; cblas_cgemm_batch(o, tas, tbs, ms, ns, ks, al, ap, las, bp, lbs, be, cp, lcs, gc, gs); }
        else { double al[4] = {0, 0, 0, 0}, be[4] = {1, 0, 1, 0}; cblas_zgemm_batch(o, tas, tbs, ms, ns, ks, al, ap, las, bp, lbs, be, cp, lcs, gc, gs); }
        break;
      }
      default:  /* 7: TRMM, 8: TRSM */
        what = fam == 7 ? "trmm" : "trsm";
        if (f77) {
          if (fam == 7) { if (prec == 0) strmm_(&cs, &cu, &cta, &cd, &m, &n, &s1, (float *)a, &lda, (float *)c, &ldb);
                          else if (prec == 1) dtrmm_(&cs, &cu, &cta, &cd, &m, &n, &d1, a, &lda, c, &ldb);
                          else if (prec == 2) ctrmm_(&cs, &cu, &cta, &cd, &m, &n, c1, (float *)a, &lda, (float *)c, &ldb);
                          else ztrmm_(&cs, &cu, &cta, &cd, &m, &n, z1, a, &lda, c, &ldb); }
          else          { if (prec == 0) strsm_(&cs, &cu, &cta, &cd, &m, &n, &s1, (float *)a, &lda, (float *)c, &ldb);
                          else if (prec == 1) dtrsm_(&cs, &cu, &cta, &cd, &m, &n, &d1, a, &lda, c, &ldb);
                          else if (prec == 2) ctrsm_(&cs, &cu, &cta, &cd, &m, &n, c1, (float *)a, &lda, (float *)c, &ldb);
                          else ztrsm_(&cs, &cu, &cta, &cd, &m, &n, z1, a, &lda, c, &ldb); }
        } else {
          if (fam == 7) { if (prec == 0) cblas_strmm(o, sd, up, ta, dg, m, n, 1.f, (float *)a, lda, (float *)c, ldb);
                          else if (prec == 1) cblas_dtrmm(o, sd, up, ta, dg, m, n, 1.0, a, lda, c, ldb);
                          else if (prec == 2) cblas_ctrmm(o, sd, up, ta, dg, m, n, c1, a, lda, c, ldb);
                          else cblas_ztrmm(o, sd, up, ta, dg, m, n, z1, a, lda, c, ldb); }
          else          { if (prec == 0) cblas_strsm(o, sd, up, ta, dg, m, n, 1.f, (float *)a, lda, (float *)c, ldb);
                          else if (prec == 1) cblas_dtrsm(o, sd, up, ta, dg, m, n, 1.0, a, lda, c, ldb);
                          else if (prec == 2) cblas_ctrsm(o, sd, up, ta, dg, m, n, c1, a, lda, c, ldb);
                          else cblas_ztrsm(o, sd, up, ta, dg, m, n, z1, a, lda, c, ldb); }
        }
    }
    (void)AL; (void)BE;
    if (calls) printf("%4d %-13s p%d %s calls=%d name='%s' info=%d\n", it, what, prec, f77 ? "f77" : "cblas", calls, last_name, last_info);
    else printf("%4d %-13s p%d %s ok\n", it, what, prec, f77 ? "f77" : "cblas");
    calls = 0; last_info = -99; last_name[0] = 0;
  }
  for (int i = 0; i < 64; i++) if (c[i] != 42.0) { printf("C was written at %d\n", i); break; }
  return 0;
}
