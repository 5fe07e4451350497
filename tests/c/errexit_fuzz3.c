/*
 * errexit_fuzz3.c -- randomized argument probes of SBGEMMT (interface/sbgemmt.c: sbgemmt_ and cblas_sbgemmt in both
 * orders and an illegal one): flags, extents and leading dimensions from small sets that include illegal values.
 * Legal calls run with alpha = 0, beta = 1 on zeroed operands, which changes nothing in either library and needs
 * no GPU here (nothing to multiply, beta == 1).  Linked against the reference (oracle/_ref/generic) it wrote
 * tests/golden/errexit_fuzz3_reference.txt; linked against libopenblas_b200.so it must print the same bytes:
 * routine name as handed to xerbla_ ("SBGEMMT ", length 9), info, and how often xerbla_ fired.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "openblas_b200.h"

static int fired, info_seen, len_seen;
static char name_seen[16];
int xerbla_(char *name, blasint *info, blasint len) {
  fired++;
  info_seen = *info;
  len_seen = len;
  memset(name_seen, 0, sizeof name_seen);
  memcpy(name_seen, name, len < 15 ? (size_t)len : 15u);
  return 0;
}

static unsigned long long lcg = 0x2545F4914F6CDD1Dull;
static unsigned draw(unsigned n) {
  lcg = lcg * 2862933555777941757ull + 3037000493ull;
  return (unsigned)((lcg >> 35) % n);
}

int main(int argc, char **argv) {
  static bfloat16 a[64], b[64];
  static float c[64];
  const int rounds = argc > 1 ? atoi(argv[1]) : 2000;
  if (argc > 2) lcg += strtoull(argv[2], NULL, 0) * 0x9FB21C651E98DF25ull;
  static const int order_set[] = {CblasColMajor, CblasRowMajor, CblasRowMajor, CblasColMajor, 100, 0};
  static const int trans_set[] = {CblasNoTrans, CblasTrans, CblasConjTrans, CblasConjNoTrans, 110, -1};
  static const int uplo_set[] = {CblasUpper, CblasLower, CblasLower, CblasUpper, 120};
  static const int extent_set[] = {-2, 0, 1, 2, 4, 6};
  static const int ld_set[] = {-1, 0, 1, 2, 4, 6, 7};
  static const char trans_ch[] = {'N', 'T', 'C', 'R', 't', 'r', '?', 'U'};
  static const char uplo_ch[] = {'U', 'L', 'l', 'u', 'N', ' '};
  for (int i = 0; i < 64; i++) c[i] = 17.5f;

  for (int r = 0; r < rounds; r++) {
    const int fortran = draw(2) == 0;
    const int order = order_set[draw(6)], ta = trans_set[draw(6)], tb = trans_set[draw(6)], up = uplo_set[draw(5)];
    blasint m = extent_set[draw(6)], k = extent_set[draw(6)], lda = ld_set[draw(7)], ldb = ld_set[draw(7)], ldc = ld_set[draw(7)];
    char cta = trans_ch[draw(8)], ctb = trans_ch[draw(8)], cup = uplo_ch[draw(6)];
    float zero = 0.f, one = 1.f;
    if (fortran) {
      sbgemmt_(&cup, &cta, &ctb, &m, &k, &zero, a, &lda, b, &ldb, &one, c, &ldc);
      printf("%4d f77 %c%c%c m=%d k=%d ld=%d,%d,%d :", r, cup, cta, ctb, (int)m, (int)k, (int)lda, (int)ldb, (int)ldc);
    } else {
      cblas_sbgemmt((enum CBLAS_ORDER)order, (enum CBLAS_UPLO)up, (enum CBLAS_TRANSPOSE)ta, (enum CBLAS_TRANSPOSE)tb, m, k, 0.f, a, lda, b,
                    ldb, 1.f, c, ldc);
      printf("%4d cblas o=%d u=%d t=%d,%d m=%d k=%d ld=%d,%d,%d :", r, order, up, ta, tb, (int)m, (int)k, (int)lda, (int)ldb, (int)ldc);
    }
    if (fired) printf(" xerbla x%d '%s' len=%d info=%d\n", fired, name_seen, len_seen, info_seen);
    else printf(" accepted\n");
    fired = 0; info_seen = -77; len_seen = 0; name_seen[0] = 0;
  }
  for (int i = 0; i < 64; i++) if (c[i] != 17.5f) { printf("C changed at %d\n", i); return 1; }
  printf("C untouched\n");
  return 0;
}
