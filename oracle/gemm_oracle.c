/*
 * gemm_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C CPU restatement of the reference OpenBLAS 0.3.28.dev GEMM path, used only as
 * the checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.  Nothing
 * under openblas_b200/ may link, load or call this file; the product path has no CPU
 * fallback and fails loudly without a GPU.
 *
 * What is restated (reference file:line next to each function):
 *   - the beta pass and the early outs of the single-threaded level-3 driver
 *   - its k-blocking rule (GEMM_Q), which fixes the ORDER in which partial sums are
 *     rounded into C -- the only part of the Goto loop nest that affects values
 *   - the arithmetic of the generic micro-kernels (separate multiply and add, sequential
 *     in k inside a block, then C += alpha * partial)
 *   - the bf16 conversion rules, the ctest checker (DMMCH) and the argument validation
 * The P/R blocking, the packing copies and the thread partition move data without
 * arithmetic and are deliberately absent.
 *
 * Parity pin: compiled with -ffp-contract=off and called with the GENERIC target's block
 * sizes (x86-64: param.h:4096-4101 -- GEMM_Q = 128 for s/d/c/z; 256 for sb, param.h:81) this file is
 * BIT-IDENTICAL to the reference built with TARGET=GENERIC (oracle/_ref/generic), and
 * agrees with the reference's SIMD builds (oracle/_ref/{haswell,skylakex,...}) to the
 * summation-order bound; tests/test_oracle_pin.py checks both on seeded inputs.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { OR_S = 0, OR_D = 1, OR_C = 2, OR_Z = 3, OR_SB = 4 };

/* ---- bf16 <-> fp32 (kernel/x86_64/tobf16.c:46-96, kernel/x86_64/bf16to.c:43-100) ------ */
uint16_t oracle_f32_to_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  switch (u & 0xff800000u) {
    case 0x00000000u: return 0x0000u;            /* +denormal / +0 */
    case 0x80000000u: return 0x8000u;            /* -denormal / -0 */
    case 0x7f800000u:
    case 0xff800000u: {
      uint16_t h = (uint16_t)(u >> 16);
      if (u & 0x007fffffu) h |= 0x0040u;         /* NaN -> quiet NaN */
      return h;
    }
    default:
      u += ((u >> 16) & 1u) + 0x7fffu;           /* round to nearest even */
      return (uint16_t)(u >> 16);
  }
}

float oracle_bf16_to_f32(uint16_t h) {
  uint32_t u;
  float f;
  switch (h & 0xff80u) {
    case 0x0000u: u = 0x00000000u; break;
    case 0x8000u: u = 0x80000000u; break;
    case 0x7f80u:
    case 0xff80u:
      u = ((uint32_t)h) << 16;
      if (h & 0x007fu) u |= 0x00400000u;
      break;
    default: u = ((uint32_t)h) << 16; break;
  }
  memcpy(&f, &u, 4);
  return f;
}

void oracle_tobf16(long n, const float *in, long inc_in, uint16_t *out, long inc_out) {
  for (long i = 0; i < n; i++) out[i * inc_out] = oracle_f32_to_bf16(in[i * inc_in]);
}
void oracle_bf16to(long n, const uint16_t *in, long inc_in, float *out, long inc_out) {
  for (long i = 0; i < n; i++) out[i * inc_out] = oracle_bf16_to_f32(in[i * inc_in]);
}

/* The GEMM micro-kernel widens bf16 by a plain shift (kernel/generic/gemmkernel_2x2.c:3-16),
 * without the denormal flush of bf16to.c. */
static inline float widen_bf16(uint16_t h) {
  uint32_t u = ((uint32_t)h) << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

/* ---- argument validation (interface/gemm.c:271-290 Fortran, :411-420/:460-476 CBLAS) --
 * trans codes: 0 N, 1 T, 2 R, 3 C, -1 illegal.  Returns the info value xerbla_ would get,
 * or `ok` when all checks pass. */
int oracle_check_args(int transa, int transb, long m, long n, long k, long lda, long ldb,
                      long ldc, int ok) {
  long nrowa = (transa & 1) ? k : m;
  long nrowb = (transb & 1) ? n : k;
  int info = ok;
  if (ldc < m) info = 13;
  if (ldb < nrowb) info = 10;
  if (lda < nrowa) info = 8;
  if (k < 0) info = 5;
  if (n < 0) info = 4;
  if (m < 0) info = 3;
  if (transb < 0) info = 2;
  if (transa < 0) info = 1;
  return info;
}

/* ---- k-block length (driver/level3/level3.c:292-305) --------------------------------- */
static long next_min_l(long remaining, long q, long unroll_m) {
  long min_l = remaining;
  if (min_l >= q * 2) {
    min_l = q;
  } else if (min_l > q) {
    min_l = ((min_l / 2 + unroll_m - 1) / unroll_m) * unroll_m;
  }
  return min_l;
}

/* element accessors: op(A)(i,l) and op(B)(l,j) of the column-major stored operands */
#define A_AT(i, l) ((transa & 1) ? ((l) + (i) * lda) : ((i) + (l) * lda))
#define B_AT(l, j) ((transb & 1) ? ((j) + (l) * ldb) : ((l) + (j) * ldb))

/* ---- real GEMM: s, d, sb ---------------------------------------------------------------
 * beta pass      kernel/generic/gemm_beta.c:41-94   (beta == 0 stores zeros, never reads C)
 * driver         driver/level3/level3.c:229-259     (beta != 1 -> scale; k == 0 or
 *                                                    alpha == 0 -> return)
 * micro-kernel   kernel/generic/gemmkernel_2x2.c:37-99  res = res + a*b sequentially in k,
 *                                                    then C = C + res*alpha               */
#define REAL_GEMM(NAME, TIN, TACC, LOAD)                                                      \
  static void NAME(int transa, int transb, long m, long n, long k, TACC alpha, const TIN *a,  \
                   long lda, const TIN *b, long ldb, TACC beta, TACC *c, long ldc, long q,    \
                   long unroll_m) {                                                           \
    if (beta != (TACC)1) {                                                                    \
      for (long j = 0; j < n; j++)                                                            \
        for (long i = 0; i < m; i++) {                                                        \
          if (beta == (TACC)0) c[i + j * ldc] = (TACC)0;                                      \
          else c[i + j * ldc] *= beta;                                                        \
        }                                                                                     \
    }                                                                                         \
    if (k == 0 || alpha == (TACC)0) return;                                                   \
    for (long ls = 0; ls < k;) {                                                              \
      long min_l = next_min_l(k - ls, q, unroll_m);                                           \
      for (long j = 0; j < n; j++)                                                            \
        for (long i = 0; i < m; i++) {                                                        \
          TACC res = 0;                                                                       \
          for (long l = ls; l < ls + min_l; l++) {                                            \
            TACC prod = LOAD(a[A_AT(i, l)]) * LOAD(b[B_AT(l, j)]);                            \
            res = res + prod;                                                                 \
          }                                                                                   \
          res = res * alpha;                                                                  \
          c[i + j * ldc] = c[i + j * ldc] + res;                                              \
        }                                                                                     \
      ls += min_l;                                                                            \
    }                                                                                         \
  }

#define IDENT(x) (x)
REAL_GEMM(gemm_s, float, float, IDENT)
REAL_GEMM(gemm_d, double, double, IDENT)
REAL_GEMM(gemm_sb, uint16_t, float, widen_bf16)

/* ---- complex GEMM: c, z ----------------------------------------------------------------
 * beta pass      kernel/generic/zgemm_beta.c:47-155  (re = br*cr - bi*ci ; im = br*ci + bi*cr)
 * micro-kernel   kernel/generic/zgemmkernel_2x2.c:36-131 (NN), :133-229 (conj B), :231-327
 *                (conj A), :329-425 (both): per k  re += ar*br ; im += ai*br ; re -= ai*bi ;
 *                im += ar*bi with the signs of the conj variants; then :539-546
 *                C_re += re*alpha_r ; C_im += im*alpha_r ; C_re -= im*alpha_i ; C_im += re*alpha_i */
#define CPLX_GEMM(NAME, T)                                                                    \
  static void NAME(int transa, int transb, long m, long n, long k, const T *alpha, const T *a, \
                   long lda, const T *b, long ldb, const T *beta, T *c, long ldc, long q,     \
                   long unroll_m) {                                                           \
    const T alr = alpha[0], ali = alpha[1], ber = beta[0], bei = beta[1];                     \
    const int conja = (transa & 2) != 0, conjb = (transb & 2) != 0;                           \
    if (ber != (T)1 || bei != (T)0) {                                                         \
      for (long j = 0; j < n; j++)                                                            \
        for (long i = 0; i < m; i++) {                                                        \
          T *p = c + 2 * (i + j * ldc);                                                       \
          if (ber == (T)0 && bei == (T)0) { p[0] = 0; p[1] = 0; }                             \
          else {                                                                              \
            T cr = p[0], ci = p[1];                                                           \
            T t1 = ber * cr, t2 = bei * ci, t3 = ber * ci, t4 = bei * cr;                     \
            p[0] = t1 - t2; p[1] = t3 + t4;                                                   \
          }                                                                                   \
        }                                                                                     \
    }                                                                                         \
    if (k == 0 || (alr == (T)0 && ali == (T)0)) return;                                       \
    for (long ls = 0; ls < k;) {                                                              \
      long min_l = next_min_l(k - ls, q, unroll_m);                                           \
      for (long j = 0; j < n; j++)                                                            \
        for (long i = 0; i < m; i++) {                                                        \
          T re = 0, im = 0;                                                                   \
          for (long l = ls; l < ls + min_l; l++) {                                            \
            const T *pa = a + 2 * A_AT(i, l), *pb = b + 2 * B_AT(l, j);                       \
            T ar = pa[0], ai = pa[1], br = pb[0], bi = pb[1];                                 \
            T p_rr = ar * br, p_ir = ai * br, p_ii = ai * bi, p_ri = ar * bi;                 \
            re = re + p_rr;                                                                   \
            if (!conja) im = im + p_ir; else im = im - p_ir;                                  \
            if (conja == conjb) re = re - p_ii; else re = re + p_ii;                          \
            if (!conjb) im = im + p_ri; else im = im - p_ri;                                  \
          }                                                                                   \
          T *p = c + 2 * (i + j * ldc);                                                       \
          T l0 = re * alr; p[0] = p[0] + l0;                                                  \
          T l1 = im * alr; p[1] = p[1] + l1;                                                  \
          l0 = im * ali;   p[0] = p[0] - l0;                                                  \
          l1 = re * ali;   p[1] = p[1] + l1;                                                  \
        }                                                                                     \
      ls += min_l;                                                                            \
    }                                                                                         \
  }

CPLX_GEMM(gemm_c, float)
CPLX_GEMM(gemm_z, double)

/* GENERIC-target block sizes on x86-64: GEMM_Q param.h:4096-4101 (s/d/c/z), param.h:76-82 (sb);
 * GEMM_UNROLL_M param.h:4037-4043 */
static const long generic_q[5] = {128, 128, 128, 128, 256};
static const long generic_unroll_m[5] = {2, 2, 2, 2, 8};

/* One column-major GEMM.  q <= 0 selects the GENERIC target's GEMM_Q / GEMM_UNROLL_M.
 * Arguments are NOT validated here (use oracle_check_args); m == 0 || n == 0 returns
 * (interface/gemm.c:494). */
int oracle_gemm(int dtype, int transa, int transb, long m, long n, long k, const void *alpha,
                const void *a, long lda, const void *b, long ldb, const void *beta, void *c,
                long ldc, long q, long unroll_m) {
  if (m <= 0 || n <= 0) return 0;
  if (dtype < 0 || dtype > 4) return -1;
  if (q <= 0) { q = generic_q[dtype]; unroll_m = generic_unroll_m[dtype]; }
  if (unroll_m <= 0) unroll_m = 1;
  switch (dtype) {
    case OR_S:
      gemm_s(transa, transb, m, n, k, *(const float *)alpha, (const float *)a, lda,
             (const float *)b, ldb, *(const float *)beta, (float *)c, ldc, q, unroll_m);
      break;
    case OR_D:
      gemm_d(transa, transb, m, n, k, *(const double *)alpha, (const double *)a, lda,
             (const double *)b, ldb, *(const double *)beta, (double *)c, ldc, q, unroll_m);
      break;
    case OR_SB:
      gemm_sb(transa, transb, m, n, k, *(const float *)alpha, (const uint16_t *)a, lda,
              (const uint16_t *)b, ldb, *(const float *)beta, (float *)c, ldc, q, unroll_m);
      break;
    case OR_C:
      gemm_c(transa, transb, m, n, k, (const float *)alpha, (const float *)a, lda,
             (const float *)b, ldb, (const float *)beta, (float *)c, ldc, q, unroll_m);
      break;
    default:
      gemm_z(transa, transb, m, n, k, (const double *)alpha, (const double *)a, lda,
             (const double *)b, ldb, (const double *)beta, (double *)c, ldc, q, unroll_m);
      break;
  }
  return 0;
}

/* ---- small-matrix path (interface/gemm.c:551-571) ------------------------------------------
 * kernel/generic/gemm_small_matrix_kernel_nn.c:37-55 and zgemm_small_matrix_kernel_nn.c:41-88:
 * one running sum over the whole k, then C = C*beta + alpha*sum (B0 variant: C = alpha*sum,
 * C never read).  On x86-64 this is what every call with m*n*k <= 100^3 takes
 * (kernel/x86_64/dgemm_small_kernel_permit_skylakex.c:30-43), i.e. all ctest calls. */
int oracle_gemm_small(int dtype, int transa, int transb, long m, long n, long k, const void *alpha,
                      const void *a, long lda, const void *b, long ldb, const void *beta, void *c,
                      long ldc) {
  if (m <= 0 || n <= 0) return 0;
  for (long i = 0; i < m; i++)
    for (long j = 0; j < n; j++) {
      long ic = i + j * ldc;
      if (dtype == OR_S || dtype == OR_SB) {
        float res = 0, al = *(const float *)alpha, be = *(const float *)beta;
        for (long l = 0; l < k; l++) {
          float av = dtype == OR_S ? ((const float *)a)[A_AT(i, l)] : widen_bf16(((const uint16_t *)a)[A_AT(i, l)]);
          float bv = dtype == OR_S ? ((const float *)b)[B_AT(l, j)] : widen_bf16(((const uint16_t *)b)[B_AT(l, j)]);
          float prod = av * bv;
          res += prod;
        }
        float *pc = (float *)c + ic;
        if (be == 0.0f) *pc = al * res; else { float t = *pc * be, u = al * res; *pc = t + u; }
      } else if (dtype == OR_D) {
        double res = 0, al = *(const double *)alpha, be = *(const double *)beta;
        for (long l = 0; l < k; l++) {
          double prod = ((const double *)a)[A_AT(i, l)] * ((const double *)b)[B_AT(l, j)];
          res += prod;
        }
        double *pc = (double *)c + ic;
        if (be == 0.0) *pc = al * res; else { double t = *pc * be, u = al * res; *pc = t + u; }
      } else {
        double re = 0, im = 0;   /* accumulate in the working precision below */
        float ref_ = 0, imf_ = 0;
        const int dbl = dtype == OR_Z;
        for (long l = 0; l < k; l++) {
          long ia = A_AT(i, l), ib = B_AT(l, j);
          if (dbl) {
            double ar = ((const double *)a)[2 * ia], ai = ((const double *)a)[2 * ia + 1];
            double br = ((const double *)b)[2 * ib], bi = ((const double *)b)[2 * ib + 1];
            if (transa & 2) ai = -ai;
            if (transb & 2) bi = -bi;
            double t1 = ar * br, t2 = ai * bi, t3 = ar * bi, t4 = ai * br;
            re += (t1 - t2); im += (t3 + t4);
          } else {
            float ar = ((const float *)a)[2 * ia], ai = ((const float *)a)[2 * ia + 1];
            float br = ((const float *)b)[2 * ib], bi = ((const float *)b)[2 * ib + 1];
            if (transa & 2) ai = -ai;
            if (transb & 2) bi = -bi;
            float t1 = ar * br, t2 = ai * bi, t3 = ar * bi, t4 = ai * br;
            ref_ += (t1 - t2); imf_ += (t3 + t4);
          }
        }
        if (dbl) {
          const double *al = (const double *)alpha, *be = (const double *)beta;
          double *pc = (double *)c + 2 * ic;
          double t0 = 0, t1 = 0;
          if (!(be[0] == 0.0 && be[1] == 0.0)) { t0 = be[0] * pc[0] - be[1] * pc[1]; t1 = be[0] * pc[1] + be[1] * pc[0]; }
          pc[0] = t0 + al[0] * re - al[1] * im;
          pc[1] = t1 + al[0] * im + re * al[1];
        } else {
          const float *al = (const float *)alpha, *be = (const float *)beta;
          float *pc = (float *)c + 2 * ic;
          float t0 = 0, t1 = 0;
          if (!(be[0] == 0.0f && be[1] == 0.0f)) { t0 = be[0] * pc[0] - be[1] * pc[1]; t1 = be[0] * pc[1] + be[1] * pc[0]; }
          pc[0] = t0 + al[0] * ref_ - al[1] * imf_;
          pc[1] = t1 + al[0] * imf_ + ref_ * al[1];
        }
      }
    }
  return 0;
}

/* ---- the ctest checker (ctest/c_dblat3.f:2198-2318 DMMCH, c_zblat3.f ZMMCH) -----------
 * Recomputes ct = alpha*sum(a*b) + beta*c in the working precision and the gauge
 * g = |alpha|*sum|a||b| + |beta||c| (complex: ABS1(z) = |re|+|im|), and returns
 * max_ij |ct - cc| / (eps * g)  (g == 0 -> the absolute difference / eps).
 * c0 is C before the call, cc the result under test.  The suite passes when this is < 16. */
double oracle_mmch(int dtype, int transa, int transb, long m, long n, long k, const void *alpha,
                   const void *a, long lda, const void *b, long ldb, const void *beta,
                   const void *c0, long ldc0, const void *cc, long ldcc) {
  double err = 0.0;
  const int cplx = (dtype == OR_C || dtype == OR_Z);
  const double eps = (dtype == OR_D || dtype == OR_Z) ? ldexp(1.0, -53) * 2 : ldexp(1.0, -24) * 2;
  for (long j = 0; j < n; j++)
    for (long i = 0; i < m; i++) {
      double ctr = 0, cti = 0, g = 0;
      for (long l = 0; l < k; l++) {
        double ar, ai = 0, br, bi = 0;
        long ia = A_AT(i, l), ib = B_AT(l, j);
        switch (dtype) {
          case OR_S: ar = ((const float *)a)[ia]; br = ((const float *)b)[ib]; break;
          case OR_D: ar = ((const double *)a)[ia]; br = ((const double *)b)[ib]; break;
          case OR_SB: ar = widen_bf16(((const uint16_t *)a)[ia]); br = widen_bf16(((const uint16_t *)b)[ib]); break;
          case OR_C: ar = ((const float *)a)[2 * ia]; ai = ((const float *)a)[2 * ia + 1];
                     br = ((const float *)b)[2 * ib]; bi = ((const float *)b)[2 * ib + 1]; break;
          default:   ar = ((const double *)a)[2 * ia]; ai = ((const double *)a)[2 * ia + 1];
                     br = ((const double *)b)[2 * ib]; bi = ((const double *)b)[2 * ib + 1]; break;
        }
        if (transa & 2) ai = -ai;
        if (transb & 2) bi = -bi;
        ctr += ar * br - ai * bi;
        cti += ar * bi + ai * br;
        g += (fabs(ar) + fabs(ai)) * (fabs(br) + fabs(bi));
      }
      double alr, ali = 0, ber, bei = 0, c0r, c0i = 0, ccr, cci = 0;
      long ic0 = i + j * ldc0, icc = i + j * ldcc;
      switch (dtype) {
        case OR_D: alr = *(const double *)alpha; ber = *(const double *)beta;
                   c0r = ((const double *)c0)[ic0]; ccr = ((const double *)cc)[icc]; break;
        case OR_Z: alr = ((const double *)alpha)[0]; ali = ((const double *)alpha)[1];
                   ber = ((const double *)beta)[0]; bei = ((const double *)beta)[1];
                   c0r = ((const double *)c0)[2 * ic0]; c0i = ((const double *)c0)[2 * ic0 + 1];
                   ccr = ((const double *)cc)[2 * icc]; cci = ((const double *)cc)[2 * icc + 1]; break;
        case OR_C: alr = ((const float *)alpha)[0]; ali = ((const float *)alpha)[1];
                   ber = ((const float *)beta)[0]; bei = ((const float *)beta)[1];
                   c0r = ((const float *)c0)[2 * ic0]; c0i = ((const float *)c0)[2 * ic0 + 1];
                   ccr = ((const float *)cc)[2 * icc]; cci = ((const float *)cc)[2 * icc + 1]; break;
        default:   alr = *(const float *)alpha; ber = *(const float *)beta;
                   c0r = ((const float *)c0)[ic0]; ccr = ((const float *)cc)[icc]; break;
      }
      if (ber == 0.0 && bei == 0.0) { c0r = 0; c0i = 0; } /* beta == 0: old C is not an input */
      double rr = alr * ctr - ali * cti + ber * c0r - bei * c0i;
      double ri = alr * cti + ali * ctr + ber * c0i + bei * c0r;
      double gauge = (fabs(alr) + fabs(ali)) * g + (fabs(ber) + fabs(bei)) * (fabs(c0r) + fabs(c0i));
      double diff = cplx ? fabs(rr - ccr) + fabs(ri - cci) : fabs(rr - ccr);
      double e = diff / eps;
      if (gauge != 0.0) e /= gauge;
      if (!(e <= err)) err = e; /* NaN-propagating max */
    }
  return err;
}

/* ---- the north-star componentwise ratio ------------------------------------------------
 * max_ij |x_ij - y_ij| / (k * eps * (|alpha| (|A||B|)_ij + |beta||C0_ij|)); the bound in
 * BASELINE.json is "ratio <= c" with c a small constant.  x, y: two results for the same
 * problem (e.g. GPU and reference). */
double oracle_componentwise_ratio(int dtype, int transa, int transb, long m, long n, long k,
                                  const void *alpha, const void *a, long lda, const void *b,
                                  long ldb, const void *beta, const void *c0, long ldc0,
                                  const void *x, long ldx, const void *y, long ldy) {
  double worst = 0.0;
  const int cplx = (dtype == OR_C || dtype == OR_Z);
  const int dbl = (dtype == OR_D || dtype == OR_Z);
  const double eps = dbl ? ldexp(1.0, -52) : ldexp(1.0, -23);
  double alr, ali = 0, ber, bei = 0;
  if (dtype == OR_D) { alr = *(const double *)alpha; ber = *(const double *)beta; }
  else if (dtype == OR_Z) { alr = ((const double *)alpha)[0]; ali = ((const double *)alpha)[1];
                            ber = ((const double *)beta)[0]; bei = ((const double *)beta)[1]; }
  else if (dtype == OR_C) { alr = ((const float *)alpha)[0]; ali = ((const float *)alpha)[1];
                            ber = ((const float *)beta)[0]; bei = ((const float *)beta)[1]; }
  else { alr = *(const float *)alpha; ber = *(const float *)beta; }
  const double aabs = cplx ? hypot(alr, ali) : fabs(alr), babs = cplx ? hypot(ber, bei) : fabs(ber);
  const long kk = k > 0 ? k : 1;
  for (long j = 0; j < n; j++)
    for (long i = 0; i < m; i++) {
      double g = 0;
      for (long l = 0; l < k; l++) {
        long ia = A_AT(i, l), ib = B_AT(l, j);
        double av, bv;
        switch (dtype) {
          case OR_S: av = fabs(((const float *)a)[ia]); bv = fabs(((const float *)b)[ib]); break;
          case OR_D: av = fabs(((const double *)a)[ia]); bv = fabs(((const double *)b)[ib]); break;
          case OR_SB: av = fabs(widen_bf16(((const uint16_t *)a)[ia])); bv = fabs(widen_bf16(((const uint16_t *)b)[ib])); break;
          case OR_C: av = hypot(((const float *)a)[2 * ia], ((const float *)a)[2 * ia + 1]);
                     bv = hypot(((const float *)b)[2 * ib], ((const float *)b)[2 * ib + 1]); break;
          default:   av = hypot(((const double *)a)[2 * ia], ((const double *)a)[2 * ia + 1]);
                     bv = hypot(((const double *)b)[2 * ib], ((const double *)b)[2 * ib + 1]); break;
        }
        g += av * bv;
      }
      double cabs0 = 0, diff;
      long ic0 = i + j * ldc0, ix = i + j * ldx, iy = i + j * ldy;
      if (babs != 0.0) {
        if (dtype == OR_D) cabs0 = fabs(((const double *)c0)[ic0]);
        else if (dtype == OR_Z) cabs0 = hypot(((const double *)c0)[2 * ic0], ((const double *)c0)[2 * ic0 + 1]);
        else if (dtype == OR_C) cabs0 = hypot(((const float *)c0)[2 * ic0], ((const float *)c0)[2 * ic0 + 1]);
        else cabs0 = fabs(((const float *)c0)[ic0]);
      }
      if (dtype == OR_D) diff = fabs(((const double *)x)[ix] - ((const double *)y)[iy]);
      else if (dtype == OR_Z) diff = hypot(((const double *)x)[2 * ix] - ((const double *)y)[2 * iy],
                                           ((const double *)x)[2 * ix + 1] - ((const double *)y)[2 * iy + 1]);
      else if (dtype == OR_C) diff = hypot((double)((const float *)x)[2 * ix] - ((const float *)y)[2 * iy],
                                           (double)((const float *)x)[2 * ix + 1] - ((const float *)y)[2 * iy + 1]);
      else diff = fabs((double)((const float *)x)[ix] - ((const float *)y)[iy]);
      double bound = (double)kk * eps * (aabs * g + babs * cabs0);
      double r = bound > 0 ? diff / bound : (diff == 0 ? 0.0 : INFINITY);
      if (!(r <= worst)) worst = r;
    }
  return worst;
}
