/*
 * level3_oracle.c -- TEST INFRASTRUCTURE.  CPU restatement of the rest of level 3 (SYMM/HEMM,
 * SYRK/HERK, SYR2K/HER2K, TRMM/TRSM) used to check the CUDA path; never linked into the product.
 *
 * It follows the SEMANTICS the reference implements -- the netlib definitions the reference
 * ships under reference/ and its ctest drivers check against:
 *   reference/dsymmf.f:206-296, zhemmf.f:211-304   which triangle of A is read, the Hermitian
 *                                                  diagonal taken as real, beta == 0 never reads C
 *   reference/dsyrkf.f:170-262, zherkf.f:180-330   only the named triangle of C is written;
 *                                                  alpha == 0 or k == 0 scales that triangle only,
 *                                                  beta == 1 then leaves C untouched; HERK/HER2K
 *                                                  force the diagonal's imaginary part to zero
 *   reference/dsyr2kf.f:176-326, zher2kf.f:187-369 the two products, conj(alpha) on the second
 * The SUMMATION ORDER is its own (one dot product per element of C, in the working precision of
 * the type), so it is compared with the reference's optimised drivers (driver/level3/symm_k.c,
 * level3_syrk.c, level3_syr2k.c, compiled into oracle/_ref) through the componentwise bound
 * c*k*eps*gauge, not bit for bit: PARITY PINNED BY BOUND against oracle/_ref in
 * tests/test_oracle_pin.py, not bit-exact (the GEMM oracle in gemm_oracle.c is the bit-exact one).
 *
 * `gauge` (optional, m x n doubles, leading dimension m): |alpha| * sum |x||y| + |beta| * |c0|
 * per element with |z| = |re| + |im| (ABS1 of ctest/c_zblat3.f:2478); 0 for untouched elements.
 */
#include <math.h>
#include <stddef.h>
#include <string.h>

enum { OR_S = 0, OR_D = 1, OR_C = 2, OR_Z = 3 };

#define GEN(T, SFX)                                                                                         \
  typedef struct { T re, im; } cx_##SFX;                                                                    \
  static cx_##SFX ld_##SFX(const T *p, long idx, int cplx) {                                                \
    cx_##SFX v;                                                                                             \
    if (cplx) { v.re = p[2 * idx]; v.im = p[2 * idx + 1]; } else { v.re = p[idx]; v.im = 0; }               \
    return v;                                                                                               \
  }                                                                                                         \
  static void st_##SFX(T *p, long idx, int cplx, cx_##SFX v) {                                              \
    if (cplx) { p[2 * idx] = v.re; p[2 * idx + 1] = v.im; } else p[idx] = v.re;                             \
  }                                                                                                         \
  static cx_##SFX mul_##SFX(cx_##SFX x, cx_##SFX y) {                                                       \
    cx_##SFX r; r.re = x.re * y.re - x.im * y.im; r.im = x.re * y.im + x.im * y.re; return r;               \
  }                                                                                                         \
  static cx_##SFX add_##SFX(cx_##SFX x, cx_##SFX y) { cx_##SFX r; r.re = x.re + y.re; r.im = x.im + y.im; return r; } \
  static cx_##SFX cj_##SFX(cx_##SFX x, int on) { if (on) x.im = -x.im; return x; }                          \
  static double a1_##SFX(cx_##SFX x) { return fabs((double)x.re) + fabs((double)x.im); }                    \
  /* element (i, l) of the symmetric / Hermitian matrix stored in one triangle */                          \
  static cx_##SFX sym_##SFX(const T *a, long lda, int uplo, int herm, int cplx, long i, long l) {           \
    int stored = uplo ? (i >= l) : (i <= l);                                                                \
    cx_##SFX v = stored ? ld_##SFX(a, i + l * lda, cplx) : cj_##SFX(ld_##SFX(a, l + i * lda, cplx), herm);  \
    if (herm && i == l) v.im = 0;                                                                           \
    return v;                                                                                               \
  }                                                                                                         \
  static void symm_##SFX(int cplx, int herm, int side, int uplo, long m, long n, const T *alpha, const T *a, \
                         long lda, const T *b, long ldb, const T *beta, T *c, long ldc, double *gauge) {     \
    cx_##SFX al = ld_##SFX(alpha, 0, cplx), be = ld_##SFX(beta, 0, cplx);                                   \
    int beta_zero = be.re == 0 && be.im == 0;                                                               \
    long ka = side ? n : m;                                                                                 \
    for (long j = 0; j < n; j++)                                                                            \
      for (long i = 0; i < m; i++) {                                                                        \
        cx_##SFX acc = {0, 0};                                                                              \
        double g = 0;                                                                                       \
        for (long l = 0; l < ka; l++) {                                                                     \
          cx_##SFX x = side ? ld_##SFX(b, i + l * ldb, cplx) : sym_##SFX(a, lda, uplo, herm, cplx, i, l);   \
          cx_##SFX y = side ? sym_##SFX(a, lda, uplo, herm, cplx, l, j) : ld_##SFX(b, l + j * ldb, cplx);   \
          acc = add_##SFX(acc, mul_##SFX(x, y));                                                            \
          g += a1_##SFX(x) * a1_##SFX(y);                                                                   \
        }                                                                                                   \
        cx_##SFX r = mul_##SFX(al, acc);                                                                    \
        g *= a1_##SFX(al);                                                                                  \
        if (!beta_zero) {                                                                                   \
          cx_##SFX c0 = ld_##SFX(c, i + j * ldc, cplx);                                                     \
          r = add_##SFX(r, mul_##SFX(be, c0));                                                              \
          g += a1_##SFX(be) * a1_##SFX(c0);                                                                 \
        }                                                                                                   \
        st_##SFX(c, i + j * ldc, cplx, r);                                                                  \
        if (gauge) gauge[i + j * m] = g;                                                                    \
      }                                                                                                     \
  }                                                                                                         \
  static void rankk_##SFX(int cplx, int herm, int two, int uplo, int trans, long n, long k, const T *alpha,   \
                          const T *a, long lda, const T *b, long ldb, const T *beta, T *c, long ldc,          \
                          double *gauge) {                                                                   \
    cx_##SFX al = ld_##SFX(alpha, 0, cplx), be = ld_##SFX(beta, 0, cplx);                                   \
    if (herm && !two) al.im = 0;                                                                            \
    if (herm) be.im = 0;                                                                                    \
    int product = k > 0 && !(al.re == 0 && al.im == 0);                                                     \
    int beta_zero = be.re == 0 && be.im == 0, beta_one = be.re == 1 && be.im == 0;                          \
    if (gauge) for (long x = 0; x < n * n; x++) gauge[x] = 0;                                               \
    if (!product && beta_one) return;                                                                       \
    for (long j = 0; j < n; j++)                                                                            \
      for (long i = uplo ? j : 0; i < (uplo ? n : j + 1); i++) {                                            \
        cx_##SFX acc1 = {0, 0}, acc2 = {0, 0};                                                              \
        double g = 0;                                                                                       \
        if (product)                                                                                        \
          for (long l = 0; l < k; l++) {                                                                    \
            /* first factor: row i of op(A); second: row j of op(B) (B == A for rank-k), conjugated for the Hermitian ones */ \
            cx_##SFX x = trans ? cj_##SFX(ld_##SFX(a, l + i * lda, cplx), herm) : ld_##SFX(a, i + l * lda, cplx);  \
            cx_##SFX y = trans ? ld_##SFX(two ? b : a, l + j * (two ? ldb : lda), cplx)                     \
                               : cj_##SFX(ld_##SFX(two ? b : a, j + l * (two ? ldb : lda), cplx), herm);    \
            acc1 = add_##SFX(acc1, mul_##SFX(x, y));                                                        \
            g += a1_##SFX(x) * a1_##SFX(y);                                                                 \
            if (two) {                                                                                      \
              cx_##SFX u = trans ? cj_##SFX(ld_##SFX(b, l + i * ldb, cplx), herm) : ld_##SFX(b, i + l * ldb, cplx); \
              cx_##SFX v = trans ? ld_##SFX(a, l + j * lda, cplx) : cj_##SFX(ld_##SFX(a, j + l * lda, cplx), herm); \
              acc2 = add_##SFX(acc2, mul_##SFX(u, v));                                                      \
              g += a1_##SFX(u) * a1_##SFX(v);                                                               \
            }                                                                                               \
          }                                                                                                 \
        cx_##SFX r = mul_##SFX(al, acc1);                                                                   \
        if (two) r = add_##SFX(r, mul_##SFX(cj_##SFX(al, herm), acc2));                                     \
        g *= a1_##SFX(al);                                                                                  \
        if (!beta_zero) {                                                                                   \
          cx_##SFX c0 = ld_##SFX(c, i + j * ldc, cplx);                                                     \
          if (herm && i == j) c0.im = 0;                                                                    \
          r = add_##SFX(r, mul_##SFX(be, c0));                                                              \
          g += a1_##SFX(be) * a1_##SFX(c0);                                                                 \
        }                                                                                                   \
        if (herm && i == j) r.im = 0;                                                                       \
        st_##SFX(c, i + j * ldc, cplx, r);                                                                  \
        if (gauge) gauge[i + j * n] = g;                                                                    \
      }                                                                                                     \
  }

GEN(float, f)
GEN(double, d)

/* TRMM / TRSM.  E = op(A) restricted to its triangle, unit diagonal taken as 1 without reading A:
 * reference/dtrmmf.f:206-355, dtrsmf.f:237-407 (ztrmmf.f / ztrsmf.f for the conjugated forms).
 * TRSM substitutes in the working precision after scaling B by alpha, as the netlib code does. */
#define GEN_TR(T, SFX)                                                                                      \
  static cx_##SFX tri_##SFX(const T *a, long lda, int uplo, int trans, int unit, int cplx, long i, long k) { \
    /* E(i,k) = op(A)(i,k); zero outside the triangle */                                                    \
    long r = (trans & 1) ? k : i, c = (trans & 1) ? i : k;                                                  \
    cx_##SFX z = {0, 0};                                                                                    \
    if (r == c) { if (unit) { z.re = 1; return z; } }                                                       \
    else if (uplo ? (r < c) : (r > c)) return z;                                                            \
    return cj_##SFX(ld_##SFX(a, r + c * lda, cplx), trans >= 2);                                            \
  }                                                                                                         \
  static void trxm_##SFX(int cplx, int solve, int side, int uplo, int trans, int unit, long m, long n,       \
                         const T *alpha, const T *a, long lda, T *b, long ldb, double *gauge) {              \
    cx_##SFX al = ld_##SFX(alpha, 0, cplx);                                                                 \
    long ka = side ? n : m, nv = side ? m : n;     /* vector length, number of vectors */                   \
    int eff_lower = (uplo != 0) != ((trans & 1) != 0);                                                      \
    /* vector v: column v of B (left) or row v of B (right); right: x E = b  <=>  E^T x^T = b^T */          \
    for (long v = 0; v < nv; v++) {                                                                         \
      cx_##SFX x[ka > 0 ? ka : 1];                                                                          \
      double g[ka > 0 ? ka : 1];                                                                            \
      for (long i = 0; i < ka; i++) x[i] = ld_##SFX(b, side ? v + i * ldb : i + v * ldb, cplx);             \
      int low = side ? !eff_lower : eff_lower;     /* triangle of the matrix that multiplies the vector */   \
      if (!solve) {                                                                                         \
        cx_##SFX y[ka > 0 ? ka : 1];                                                                        \
        for (long i = 0; i < ka; i++) {                                                                     \
          cx_##SFX acc = {0, 0};                                                                            \
          double gg = 0;                                                                                    \
          for (long k = low ? 0 : i; k < (low ? i + 1 : ka); k++) {                                         \
            cx_##SFX e = side ? tri_##SFX(a, lda, uplo, trans, unit, cplx, k, i) : tri_##SFX(a, lda, uplo, trans, unit, cplx, i, k); \
            acc = add_##SFX(acc, mul_##SFX(e, x[k]));                                                       \
            gg += a1_##SFX(e) * a1_##SFX(x[k]);                                                             \
          }                                                                                                 \
          y[i] = mul_##SFX(al, acc);                                                                        \
          g[i] = gg * a1_##SFX(al);                                                                         \
        }                                                                                                   \
        for (long i = 0; i < ka; i++) x[i] = y[i];                                                          \
      } else {                                                                                              \
        for (long i = 0; i < ka; i++) { x[i] = mul_##SFX(al, x[i]); g[i] = 0; }                             \
        for (long ii = 0; ii < ka; ii++) {                                                                  \
          long i = low ? ii : ka - 1 - ii;                                                                  \
          cx_##SFX acc = x[i];                                                                              \
          for (long k = low ? 0 : i + 1; k < (low ? i : ka); k++) {                                         \
            cx_##SFX e = side ? tri_##SFX(a, lda, uplo, trans, unit, cplx, k, i) : tri_##SFX(a, lda, uplo, trans, unit, cplx, i, k); \
            cx_##SFX pr = mul_##SFX(e, x[k]);                                                               \
            acc.re -= pr.re; acc.im -= pr.im;                                                               \
          }                                                                                                 \
          if (!unit) {                                                                                      \
            cx_##SFX d = tri_##SFX(a, lda, uplo, trans, 0, cplx, i, i);                                     \
            T den = d.re * d.re + d.im * d.im;                                                              \
            cx_##SFX q; q.re = (acc.re * d.re + acc.im * d.im) / den; q.im = (acc.im * d.re - acc.re * d.im) / den; \
            if (!cplx) { q.re = acc.re / d.re; q.im = 0; }                                                  \
            acc = q;                                                                                        \
          }                                                                                                 \
          x[i] = acc;                                                                                       \
        }                                                                                                   \
      }                                                                                                     \
      for (long i = 0; i < ka; i++) {                                                                       \
        st_##SFX(b, side ? v + i * ldb : i + v * ldb, cplx, x[i]);                                          \
        if (gauge) gauge[side ? v + i * m : i + v * m] = g[i];                                              \
      }                                                                                                     \
    }                                                                                                       \
  }
GEN_TR(float, f)
GEN_TR(double, d)


/* C := alpha*A*B + beta*C (side 0) or alpha*B*A + beta*C (side 1); A symmetric (herm 0) or Hermitian
 * (herm 1), only its uplo triangle (0 upper, 1 lower) is read.  alpha/beta: 1 value, 2 for complex. */
int oracle_symm(int dtype, int herm, int side, int uplo, long m, long n, const void *alpha, const void *a, long lda,
                const void *b, long ldb, const void *beta, void *c, long ldc, double *gauge) {
  if (m <= 0 || n <= 0) return 0;
  int cplx = dtype == OR_C || dtype == OR_Z;
  if (dtype == OR_S || dtype == OR_C)
    symm_f(cplx, herm && cplx, side, uplo, m, n, (const float *)alpha, (const float *)a, lda, (const float *)b, ldb,
           (const float *)beta, (float *)c, ldc, gauge);
  else if (dtype == OR_D || dtype == OR_Z)
    symm_d(cplx, herm && cplx, side, uplo, m, n, (const double *)alpha, (const double *)a, lda, (const double *)b, ldb,
           (const double *)beta, (double *)c, ldc, gauge);
  else return -1;
  return 0;
}

/* rank-k / rank-2k update of the uplo triangle of the n x n matrix C.  trans 0: A (and B) n x k,
 * 1: k x n.  herm: HERK / HER2K (alpha real for HERK; beta real; alpha and beta are still passed as
 * (re, im) pairs for complex types).  two: SYR2K / HER2K. */
int oracle_rankk(int dtype, int herm, int two, int uplo, int trans, long n, long k, const void *alpha, const void *a,
                 long lda, const void *b, long ldb, const void *beta, void *c, long ldc, double *gauge) {
  if (n <= 0) return 0;
  int cplx = dtype == OR_C || dtype == OR_Z;
  if (dtype == OR_S || dtype == OR_C)
    rankk_f(cplx, herm && cplx, two, uplo, trans, n, k, (const float *)alpha, (const float *)a, lda, (const float *)b, ldb,
            (const float *)beta, (float *)c, ldc, gauge);
  else if (dtype == OR_D || dtype == OR_Z)
    rankk_d(cplx, herm && cplx, two, uplo, trans, n, k, (const double *)alpha, (const double *)a, lda, (const double *)b, ldb,
            (const double *)beta, (double *)c, ldc, gauge);
  else return -1;
  return 0;
}

/* ---- argument validation, info value xerbla_ would get or `ok` -------------------------
 * SYMM/HEMM: interface/symm.c:203-227 (side, uplo in {0, 1, -1}); rank-k family: syrk.c:158-165,
 * syr2k.c:280-288 (two != 0 adds the ldb check and moves ldc to 12). */
static long max1(long x) { return x > 1 ? x : 1; }
int oracle_check_symm(int side, int uplo, long m, long n, long lda, long ldb, long ldc, int ok) {
  int info = ok;
  if (ldc < max1(m)) info = 12;
  if (ldb < max1(m)) info = 9;
  if (lda < max1(side == 0 ? m : n)) info = 7;
  if (n < 0) info = 4;
  if (m < 0) info = 3;
  if (uplo < 0) info = 2;
  if (side < 0) info = 1;
  return info;
}
int oracle_check_rankk(int two, int uplo, int trans, long n, long k, long lda, long ldb, long ldc, int ok) {
  long nrowa = (trans & 1) ? k : n;
  int info = ok;
  if (ldc < max1(n)) info = two ? 12 : 10;
  if (two && ldb < max1(nrowa)) info = 9;
  if (lda < max1(nrowa)) info = 7;
  if (k < 0) info = 4;
  if (n < 0) info = 3;
  if (trans < 0) info = 2;
  if (uplo < 0) info = 1;
  return info;
}

/* B := alpha*op(A)*B / alpha*B*op(A) (solve 0) or the solution X of op(A)*X = alpha*B / X*op(A) = alpha*B
 * (solve 1), in place on b.  side 0 left, 1 right; uplo 0 upper, 1 lower; trans 0 N, 1 T, 2 conj, 3 conj-trans
 * (real types: 2 = N, 3 = T); unit 1 = unit diagonal.  gauge (m x n, ld m) only meaningful for solve 0. */
int oracle_trxm(int dtype, int solve, int side, int uplo, int trans, int unit, long m, long n, const void *alpha, const void *a,
                long lda, void *b, long ldb, double *gauge) {
  if (m <= 0 || n <= 0) return 0;
  int cplx = dtype == OR_C || dtype == OR_Z;
  if (!cplx) trans &= 1;
  if (dtype == OR_S || dtype == OR_C)
    trxm_f(cplx, solve, side, uplo, trans, unit, m, n, (const float *)alpha, (const float *)a, lda, (float *)b, ldb, gauge);
  else if (dtype == OR_D || dtype == OR_Z)
    trxm_d(cplx, solve, side, uplo, trans, unit, m, n, (const double *)alpha, (const double *)a, lda, (double *)b, ldb, gauge);
  else return -1;
  return 0;
}

/* The ctest check of a TRSM result (c_dblat3.f:1195-1235 multiplies the solution back with DMMCH):
 * max over elements of |op(A)*X - alpha*B0| / (eps * (sum |op(A)||X| + |alpha*B0|)), evaluated in long
 * double so the checker adds no rounding of its own.  A result passes ctest when this is below 16. */
double oracle_trsm_residual(int dtype, int side, int uplo, int trans, int unit, long m, long n, const void *alpha, const void *a,
                            long lda, const void *b0, long ldb0, const void *x, long ldx) {
  int cplx = dtype == OR_C || dtype == OR_Z, dbl = dtype == OR_D || dtype == OR_Z;
  if (!cplx) trans &= 1;
  double eps = dbl ? ldexp(1.0, -52) : ldexp(1.0, -23), worst = 0;
#define LD_RE(p, idx) (dbl ? (long double)((const double *)(p))[cplx ? 2 * (idx) : (idx)] : (long double)((const float *)(p))[cplx ? 2 * (idx) : (idx)])
#define LD_IM(p, idx) (!cplx ? 0.0L : dbl ? (long double)((const double *)(p))[2 * (idx) + 1] : (long double)((const float *)(p))[2 * (idx) + 1])
  long double are = LD_RE(alpha, 0), aim = LD_IM(alpha, 0);
  long ka = side ? n : m;
  for (long j = 0; j < n; j++)
    for (long i = 0; i < m; i++) {
      long double sre = 0, sim = 0, g = 0;
      for (long l = 0; l < ka; l++) {
        /* left: E(i,l) X(l,j); right: X(i,l) E(l,j) */
        long ei = side ? l : i, ek = side ? j : l;
        long r = (trans & 1) ? ek : ei, c = (trans & 1) ? ei : ek;
        long double ere, eim;
        if (r == c && unit) { ere = 1; eim = 0; }
        else if (r != c && (uplo ? (r < c) : (r > c))) continue;
        else { ere = LD_RE(a, r + c * lda); eim = LD_IM(a, r + c * lda); if (trans >= 2) eim = -eim; }
        long xi = side ? i + l * ldx : l + j * ldx;
        long double xre = LD_RE(x, xi), xim = LD_IM(x, xi);
        sre += ere * xre - eim * xim; sim += ere * xim + eim * xre;
        g += (fabsl(ere) + fabsl(eim)) * (fabsl(xre) + fabsl(xim));
      }
      long double bre = LD_RE(b0, i + j * ldb0), bim = LD_IM(b0, i + j * ldb0);
      long double tre = are * bre - aim * bim, tim = are * bim + aim * bre;
      long double err = fabsl(sre - tre) + fabsl(sim - tim);
      g += fabsl(tre) + fabsl(tim);
      double ratio = g > 0 ? (double)(err / (eps * g)) : (err > 0 ? 1e300 : 0.0);
      if (ratio > worst) worst = ratio;
    }
#undef LD_RE
#undef LD_IM
  return worst;
}

/* interface/trsm.c:188-197 (side, uplo, trans, unit in their decoded form, -1 = illegal) */
int oracle_check_trxm(int side, int uplo, int trans, int unit, long m, long n, long lda, long ldb, int ok) {
  long nrowa = (side & 1) ? n : m;
  int info = ok;
  if (ldb < max1(m)) info = 11;
  if (lda < max1(nrowa)) info = 9;
  if (n < 0) info = 6;
  if (m < 0) info = 5;
  if (unit < 0) info = 4;
  if (trans < 0) info = 3;
  if (uplo < 0) info = 2;
  if (side < 0) info = 1;
  return info;
}

/* ---- ?GEMMT: the uplo triangle of the m x m matrix  C := alpha * op(A) * op(B) + beta * C  ---------------
 * interface/gemmt.c:474-680 does this column by column with GEMV: column i of the triangle gets
 * SCAL_K(beta) (beta == 0 writes zeros without reading C, beta == 1 leaves it) and then, unless alpha == 0,
 * GEMV over the k columns / rows of op(A) restricted to the rows of the triangle.  The reference's own
 * acceptance test (utest/test_extensions/test_dgemmt.c:56-107, ?gemmt_trusted) defines the expected result
 * as the triangle of the GEMM result, which is what this restatement computes: one dot product per element
 * in the working precision.  Unlike the reference (gemmt.c:466-476 conjugates B IN PLACE for transb = R / C
 * and never undoes it), B is not modified.  PINNED BY BOUND against oracle/_ref/generic (?gemmt_ and
 * cblas_?gemmt compiled from interface/gemmt.c) in tests/test_oracle_pin.py.
 * transa / transb: 0 N, 1 T, 2 conj, 3 conj-trans (real types fold 2 -> 0, 3 -> 1). */
#define GEN_GEMMT(T, SFX)                                                                                   \
  static void gemmt_##SFX(int cplx, int uplo, int ta, int tb, long m, long k, const T *alpha, const T *a, long lda, \
                          const T *b, long ldb, const T *beta, T *c, long ldc, double *gauge) {              \
    cx_##SFX al = ld_##SFX(alpha, 0, cplx), be = ld_##SFX(beta, 0, cplx);                                   \
    int product = k > 0 && !(al.re == 0 && al.im == 0);                                                     \
    int beta_zero = be.re == 0 && be.im == 0, beta_one = be.re == 1 && be.im == 0;                          \
    if (gauge) for (long x = 0; x < m * m; x++) gauge[x] = 0;                                               \
    if (!product && beta_one) return;                                                                       \
    for (long j = 0; j < m; j++)                                                                            \
      for (long i = uplo ? j : 0; i < (uplo ? m : j + 1); i++) {                                            \
        cx_##SFX acc = {0, 0};                                                                              \
        double g = 0;                                                                                       \
        if (product)                                                                                        \
          for (long l = 0; l < k; l++) {                                                                    \
            cx_##SFX x = cj_##SFX(ld_##SFX(a, (ta & 1) ? l + i * lda : i + l * lda, cplx), ta & 2);         \
            cx_##SFX y = cj_##SFX(ld_##SFX(b, (tb & 1) ? j + l * ldb : l + j * ldb, cplx), tb & 2);         \
            acc = add_##SFX(acc, mul_##SFX(x, y));                                                          \
            g += a1_##SFX(x) * a1_##SFX(y);                                                                 \
          }                                                                                                 \
        cx_##SFX r = mul_##SFX(al, acc);                                                                    \
        g *= a1_##SFX(al);                                                                                  \
        if (!beta_zero) {                                                                                   \
          cx_##SFX c0 = ld_##SFX(c, i + j * ldc, cplx);                                                     \
          r = add_##SFX(r, mul_##SFX(be, c0));                                                              \
          g += a1_##SFX(be) * a1_##SFX(c0);                                                                 \
        }                                                                                                   \
        st_##SFX(c, i + j * ldc, cplx, r);                                                                  \
        if (gauge) gauge[i + j * m] = g;                                                                    \
      }                                                                                                     \
  }
GEN_GEMMT(float, f)
GEN_GEMMT(double, d)

int oracle_gemmt(int dtype, int uplo, int transa, int transb, long m, long k, const void *alpha, const void *a, long lda,
                 const void *b, long ldb, const void *beta, void *c, long ldc, double *gauge) {
  if (m <= 0) return 0;
  int cplx = dtype == OR_C || dtype == OR_Z;
  if (!cplx) { transa &= 1; transb &= 1; }
  if (dtype == OR_S || dtype == OR_C)
    gemmt_f(cplx, uplo, transa, transb, m, k, (const float *)alpha, (const float *)a, lda, (const float *)b, ldb,
            (const float *)beta, (float *)c, ldc, gauge);
  else if (dtype == OR_D || dtype == OR_Z)
    gemmt_d(cplx, uplo, transa, transb, m, k, (const double *)alpha, (const double *)a, lda, (const double *)b, ldb,
            (const double *)beta, (double *)c, ldc, gauge);
  else return -1;
  return 0;
}

/* interface/gemmt.c:163-186 (column-major, both ABIs): info xerbla_ gets, or `ok`.  The row-major branch
 * (:365-389) runs the same checks on the swapped problem but reports the swapped positions: 10 for the
 * leading dimension it calls lda (the caller's LDB), 8 for ldb (the caller's LDA), the 10 check last of the
 * two; 3 / 2 for the trans arguments. */
int oracle_check_gemmt(int rowmajor, int uplo, int transa, int transb, long m, long k, long lda, long ldb, long ldc, int ok) {
  int info = ok;
  if (!rowmajor) {
    long nrowa = (transa & 1) ? k : m, nrowb = (transb & 1) ? m : k;
    if (ldc < max1(m)) info = 13;
    if (ldb < max1(nrowb)) info = 10;
    if (lda < max1(nrowa)) info = 8;
    if (k < 0) info = 5;
    if (m < 0) info = 4;
    if (transb < 0) info = 3;
    if (transa < 0) info = 2;
    if (uplo < 0) info = 1;
  } else {   /* arguments as the swapped column-major problem sees them: a = B, b = A */
    long ncola = (transa & 1) ? k : m, ncolb = (transb & 1) ? m : k;
    if (ldc < max1(m)) info = 13;
    if (ldb < max1(ncolb)) info = 8;
    if (lda < max1(ncola)) info = 10;
    if (k < 0) info = 5;
    if (m < 0) info = 4;
    if (transb < 0) info = 2;
    if (transa < 0) info = 3;
    if (uplo < 0) info = 1;
  }
  return info;
}

/* ---- SBGEMV / SBDOT (bf16 in, fp32 accumulate and out) ---------------------------------------------------
 * kernel/x86_64/sbgemv_n.c:44-80, sbgemv_t.c (the C kernels every target without AVX512-BF16 runs): both
 * operands widened to fp32 (exact), one fp32 accumulator per output walked in storage order,
 * y = alpha * acc (beta == 0: y not read) or alpha * acc + beta * y.  interface/sbgemv.c:171-183: m == 0 or
 * n == 0 returns, alpha == 0 only scales y by beta (beta == 1: untouched), negative increments walk from the
 * far end.  kernel/x86_64/sbdot.c:37-58: both vectors widened to fp32, SDOT.  x and y are passed as the
 * interface passes them to the kernel (already moved to the logical first element).
 * PINNED against oracle/_ref/generic (sbgemv_, cblas_sbgemv, sbdot_, cblas_sbdot) in tests/test_oracle_pin.py:
 * bit for bit on the generic target's C kernels for SBGEMV, by bound for SBDOT (the SDOT kernel accumulates in
 * double, kernel/arm/dot.c). */
static float bf16_widen(unsigned short v) { unsigned int u = (unsigned int)v << 16; float f; memcpy(&f, &u, 4); return f; }

int oracle_sbgemv(int trans, long m, long n, float alpha, const unsigned short *a, long lda, const unsigned short *x, long incx,
                  float beta, float *y, long incy, double *gauge) {
  if (m <= 0 || n <= 0) return 0;
  long lenx = trans ? m : n, leny = trans ? n : m;
  if (incx < 0) x -= (lenx - 1) * incx;
  if (incy < 0) y -= (leny - 1) * incy;
  if (gauge) for (long i = 0; i < leny; i++) gauge[i] = 0;
  if (alpha == 0.0f) {
    if (beta != 1.0f) for (long i = 0; i < leny; i++) {
      if (gauge) gauge[i] = fabs((double)beta) * fabs((double)y[i * incy]);
      y[i * incy] = beta == 0.0f ? 0.0f : beta * y[i * incy];
    }
    return 0;
  }
  for (long i = 0; i < leny; i++) {
    float acc = 0.0f;
    double g = 0;
    for (long j = 0; j < lenx; j++) {
      float av = bf16_widen(trans ? a[j + i * lda] : a[i + j * lda]), xv = bf16_widen(x[j * incx]);
      acc += av * xv;
      g += fabs((double)av) * fabs((double)xv);
    }
    g *= fabs((double)alpha);
    if (beta == 0.0f) y[i * incy] = alpha * acc;
    else { g += fabs((double)beta) * fabs((double)y[i * incy]); y[i * incy] = alpha * acc + beta * y[i * incy]; }
    if (gauge) gauge[i] = g;
  }
  return 0;
}

double oracle_sbdot(long n, const unsigned short *x, long incx, const unsigned short *y, long incy, double *gauge) {
  if (gauge) *gauge = 0;
  if (n <= 0) return 0.0;
  if (incx < 0) x -= (n - 1) * incx;
  if (incy < 0) y -= (n - 1) * incy;
  double acc = 0, g = 0;
  for (long i = 0; i < n; i++) {
    double xv = bf16_widen(x[i * incx]), yv = bf16_widen(y[i * incy]);
    acc += xv * yv;
    g += fabs(xv) * fabs(yv);
  }
  if (gauge) *gauge = g;
  return (double)(float)acc;
}

/* interface/sbgemv.c:86-92 / 130-135 (after the row-major swap of m, n and trans) */
/* ---- SBGEMMT: the uplo triangle of the m x m fp32 matrix  C := alpha * op(A) * op(B) + beta * C,  A and B bf16 ----
 * interface/sbgemmt.c:346-445 walks the triangle column by column and hands each column to the SBGEMV KERNEL
 * (not the interface): kernel/x86_64/sbgemv_n.c:114 / sbgemv_t.c return at once when a dimension is < 1, so
 * k == 0 leaves C untouched whatever beta is, and there is no alpha == 0 shortcut -- every element is
 * alpha * acc (beta == 0, C not read) or alpha * acc + beta * C with one fp32 accumulator walked over k in order
 * (sbgemv_n.c:63-72).  R / C fold to N / T (:90-104).  PINNED against oracle/_ref/generic (sbgemmt_, cblas_sbgemmt) in
 * tests/test_f_rows_pin.py: bit for bit while a column stays on one thread, by bound otherwise. */
int oracle_sbgemmt(int uplo, int transa, int transb, long m, long k, float alpha, const unsigned short *a, long lda,
                   const unsigned short *b, long ldb, float beta, float *c, long ldc, double *gauge) {
  if (gauge) for (long x = 0; x < m * m; x++) gauge[x] = 0;
  if (m <= 0 || k <= 0) return 0;
  transa &= 1; transb &= 1;
  for (long j = 0; j < m; j++)
    for (long i = uplo ? j : 0; i < (uplo ? m : j + 1); i++) {
      float acc = 0.0f;
      double g = 0;
      for (long l = 0; l < k; l++) {
        float av = bf16_widen(transa ? a[l + i * lda] : a[i + l * lda]), bv = bf16_widen(transb ? b[j + l * ldb] : b[l + j * ldb]);
        acc += av * bv;
        g += fabs((double)av) * fabs((double)bv);
      }
      g *= fabs((double)alpha);
      float *y = c + i + j * ldc;
      if (beta == 0.0f) *y = alpha * acc;
      else { g += fabs((double)beta) * fabs((double)*y); *y = alpha * acc + beta * *y; }
      if (gauge) gauge[i + j * m] = g;
    }
  return 0;
}

/* interface/sbgemmt.c:121-137 (column-major, both ABIs) and :286-301 (row-major: the checks run on the swapped
 * problem and report the swapped positions, as in gemmt.c); unlike gemmt.c:322-323 the row-major branch does NOT
 * flip uplo (:239-240), so a row-major call updates the OTHER triangle of the caller's C -- reproduced as is. */
int oracle_check_sbgemmt(int rowmajor, int uplo, int transa, int transb, long m, long k, long lda, long ldb, long ldc, int ok) {
  return oracle_check_gemmt(rowmajor, uplo, transa, transb, m, k, lda, ldb, ldc, ok);
}

int oracle_check_sbgemv(int trans, long m, long n, long lda, long incx, long incy, int ok) {
  int info = ok;
  if (incy == 0) info = 11;
  if (incx == 0) info = 8;
  if (lda < max1(m)) info = 6;
  if (n < 0) info = 3;
  if (m < 0) info = 2;
  if (trans < 0) info = 1;
  return info;
}
