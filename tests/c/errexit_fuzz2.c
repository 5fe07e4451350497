/*
 * errexit_fuzz2.c -- the randomized argument probes of errexit_fuzz.c for the entry points added in round 2:
 * ?gemmt (interface/gemmt.c) for s, d, c, z and sbgemv (interface/sbgemv.c), CBLAS in both orders and an illegal
 * one, and the Fortran ABI; flags, extents, leading dimensions and increments from small sets that include illegal
 * values.  Every legal call is a no-op (alpha = 0, beta = 1), so neither the reference nor this library computes
 * anything and no GPU is needed.  Linked against the reference (oracle/_ref/generic) it wrote
 * tests/golden/errexit_fuzz2_reference.txt; linked against libopenblas_b200.so it must print the same bytes.
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
  const int iterations = argc > 1 ? atoi(argv[1]) : 3000;
  if (argc > 2) state ^= strtoull(argv[2], NULL, 0) * 0xD6E8FEB86659FD93ull;
  static const int orders[] = {CblasColMajor, CblasRowMajor, CblasColMajor, CblasRowMajor, 0};
  static const int transes[] = {CblasNoTrans, CblasTrans, CblasConjTrans, CblasConjNoTrans, 0, 7};
  static const int uplos[] = {CblasUpper, CblasLower, CblasUpper, CblasLower, 3};
  static const int dims[] = {-1, 0, 1, 2, 3, 5};
  static const int lds[] = {0, 1, 2, 3, 5, 6};
  static const int incs[] = {-2, -1, 0, 1, 2, 3};
  static const char tch[] = {'N', 'T', 'C', 'R', 'n', 'c', 'X', 'q'};
  static const char uch[] = {'U', 'L', 'u', 'l', 'Z'};
  double z0[2] = {0, 0}, z1[2] = {1, 0};
  float c0[2] = {0, 0}, c1[2] = {1, 0};
  double d0 = 0, d1 = 1; float s0 = 0, s1 = 1;
  for (int i = 0; i < 64; i++) c[i] = 42.0;

  for (int it = 0; it < iterations; it++) {
    const int fam = (int)rnd(5), prec = (int)rnd(4), f77 = rnd(3) == 0;
    const enum CBLAS_ORDER o = (enum CBLAS_ORDER)pick(orders, 5);
    const enum CBLAS_TRANSPOSE ta = (enum CBLAS_TRANSPOSE)pick(transes, 6), tb = (enum CBLAS_TRANSPOSE)pick(transes, 6);
    const enum CBLAS_UPLO up = (enum CBLAS_UPLO)pick(uplos, 5);
    blasint m = pick(dims, 6), n = pick(dims, 6), k = pick(dims, 6), lda = pick(lds, 6), ldb = pick(lds, 6), ldc = pick(lds, 6);
    blasint incx = pick(incs, 6), incy = pick(incs, 6);
    char cta = tch[rnd(8)], ctb = tch[rnd(8)], cu = uch[rnd(5)];
    const char *what;
    if (fam < 4) {        /* GEMMT */
      what = "gemmt";
      if (f77) {
        if (prec == 0) sgemmt_(&cu, &cta, &ctb, &m, &k, &s0, (float *)a, &lda, (float *)b, &ldb, &s1, (float *)c, &ldc);
        else if (prec == 1) dgemmt_(&cu, &cta, &ctb, &m, &k, &d0, a, &lda, b, &ldb, &d1, c, &ldc);
        else if (prec == 2) cgemmt_(&cu, &cta, &ctb, &m, &k, c0, (float *)a, &lda, (float *)b, &ldb, c1, (float *)c, &ldc);
        else zgemmt_(&cu, &cta, &ctb, &m, &k, z0, a, &lda, b, &ldb, z1, c, &ldc);
      } else {
        if (prec == 0) cblas_sgemmt(o, up, ta, tb, m, k, 0.f, (float *)a, lda, (float *)b, ldb, 1.f, (float *)c, ldc);
        else if (prec == 1) cblas_dgemmt(o, up, ta, tb, m, k, 0.0, a, lda, b, ldb, 1.0, c, ldc);
        else if (prec == 2) cblas_cgemmt(o, up, ta, tb, m, k, c0, a, lda, b, ldb, c1, c, ldc);
        else cblas_zgemmt(o, up, ta, tb, m, k, z0, a, lda, b, ldb, z1, c, ldc);
      }
    } else {              /* SBGEMV */
      what = "sbgemv";
      if (f77) sbgemv_(&cta, &m, &n, &s0, (bfloat16 *)a, &lda, (bfloat16 *)b, &incx, &s1, (float *)c, &incy);
      else cblas_sbgemv(o, ta, m, n, 0.f, (bfloat16 *)a, lda, (bfloat16 *)b, incx, 1.f, (float *)c, incy);
    }
    if (calls) printf("%4d %-7s p%d %s calls=%d name='%s' info=%d\n", it, what, prec, f77 ? "f77" : "cblas", calls, last_name, last_info);
    else printf("%4d %-7s p%d %s ok\n", it, what, prec, f77 ? "f77" : "cblas");
    calls = 0; last_info = -99; last_name[0] = 0;
  }
  /* sbdot: n <= 0 returns 0 without reading anything */
  printf("sbdot n=0 %g n=-3 %g\n", (double)cblas_sbdot(0, (bfloat16 *)a, 1, (bfloat16 *)b, 1), (double)cblas_sbdot(-3, (bfloat16 *)a, 0, (bfloat16 *)b, 0));
  for (int i = 0; i < 64; i++) if (c[i] != 42.0) { printf("C was written at %d\n", i); break; }
  return 0;
}
