/*
 * shim.h -- the thin internal C ABI between the host-side C interface layer
 * (interface_gemm.c, interface_batch.c, control.c) and the CUDA side (runtime.cu + kernels).
 * Plain C types only.  Public extension symbols live in include/openblas_b200.h; the
 * symbols declared here are library-internal (hidden visibility).
 */
#ifndef B200_SHIM_H
#define B200_SHIM_H

#include <stdint.h>
#include <stddef.h>
#include "../../include/openblas_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define B200_HIDDEN __attribute__((visibility("hidden")))
#define B200_EXPORT __attribute__((visibility("default")))

/* One normalised (column-major, op codes 0..3) GEMM problem.  The host-side analogue of
 * the reference's blas_arg_t (common_macro.h:2641-2663), minus the thread fields. */
typedef struct b200_problem {
  int dtype;            /* enum b200_dtype */
  int transa, transb;   /* enum b200_trans */
  int64_t m, n, k;
  int64_t lda, ldb, ldc;
  const void *alpha;    /* host pointer: 1 FLOAT, or 2 for complex */
  const void *beta;
  const void *a;
  const void *b;
  void *c;
  int algo3m;           /* 1: called through a GEMM3M entry point (complex types): three real products may be used */
} b200_problem;

/* element size in bytes of A/B and of C for a dtype */
static inline size_t b200_in_size(int dtype) {
  switch (dtype) { case B200_S: return 4; case B200_D: return 8; case B200_C: return 8;
                   case B200_Z: return 16; default: return 2; }
}
static inline size_t b200_out_size(int dtype) {
  switch (dtype) { case B200_S: return 4; case B200_D: return 8; case B200_C: return 8;
                   case B200_Z: return 16; default: return 4; }
}
static inline int b200_is_complex(int dtype) { return dtype == B200_C || dtype == B200_Z; }

/* Synchronous solve of one problem whose pointers may be host or device (runtime.cu). */
B200_HIDDEN int b200_run_problem(const b200_problem *p);

/* Synchronous solve of `count` independent problems (gemm_batch): staged together, one
 * stream, one final synchronise. */
B200_HIDDEN int b200_run_batch(const b200_problem *p, int64_t count);

/* One normalised (column-major) problem of the symmetric level-3 family; the
 * analogue of blas_arg_t as interface/symm.c / syrk.c / syr2k.c fill it. */
enum b200_l3_routine { B200_SYMM = 0, B200_HEMM, B200_SYRK, B200_HERK, B200_SYR2K, B200_HER2K, B200_TRMM, B200_TRSM, B200_GEMMT };
typedef struct b200_l3_problem {
  int routine;          /* enum b200_l3_routine */
  int dtype;            /* enum b200_dtype (S, D, C, Z) */
  int side;             /* SYMM/HEMM: 0 = C := alpha*A*B + beta*C, 1 = C := alpha*B*A + beta*C */
  int uplo;             /* 0 = upper, 1 = lower: triangle of A (SYMM/HEMM) or of C (the others) */
  int trans;            /* SYRK family: 0 = A (and B) are n x k, 1 = k x n (T, or C for HERK/HER2K);
                           TRMM/TRSM: enum b200_trans of op(A) */
  int transb;           /* GEMMT only: enum b200_trans of op(B) (trans holds op(A)'s); C is n x n, op(A) n x k, op(B) k x n */
  int unit;             /* TRMM/TRSM: 1 = unit diagonal (never read) */
  int64_t m, n, k;      /* SYMM/HEMM: C is m x n; SYRK family: C is n x n */
  int64_t lda, ldb, ldc;
  double alpha[2], beta[2];   /* by value; real scalars and HERK's real alpha/beta have im = 0 */
  const void *a;        /* the symmetric / Hermitian / triangular matrix (SYMM/HEMM/TRMM/TRSM), else A */
  const void *b;        /* the m x n matrix (SYMM/HEMM), B (SYR2K/HER2K), unused otherwise */
  void *c;              /* TRMM/TRSM: the m x n matrix B, overwritten (ldc = its leading dimension) */
} b200_l3_problem;

/* Synchronous solve; pointers may be host or device (runtime.cu). */
B200_HIDDEN int b200_run_level3(const b200_l3_problem *p);

/* SBGEMV (interface/sbgemv.c): y := alpha * op(A) x + beta * y, A m x n bf16 column-major, x bf16, y fp32; x and y
 * point at the LOGICAL first element (the interface moved them for negative increments).  SBDOT likewise. */
B200_HIDDEN int b200_run_sbgemv(int trans, int64_t m, int64_t n, float alpha, const void *a, int64_t lda, const void *x,
                                int64_t incx, float beta, void *y, int64_t incy);
B200_HIDDEN int b200_run_sbdot(int64_t n, const void *x, int64_t incx, const void *y, int64_t incy, float *result);

/* bf16 <-> fp32/fp64 conversion on the device, strided host or device arrays.
 * dir: 0 = float->bf16, 1 = double->bf16, 2 = bf16->float, 3 = bf16->double */
B200_HIDDEN int b200_run_convert(int dir, int64_t n, const void *in, int64_t inc_in, void *out,
                                 int64_t inc_out);

/* fatal: print the thread's last error and abort() -- there is no CPU fallback. */
B200_HIDDEN void b200_fatal(const char *where, int err);

#ifdef __cplusplus
}
#endif
#endif
