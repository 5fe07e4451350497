/*
 * openblas_b200.h -- C ABI of the B200-native GEMM library (libopenblas_b200.so).
 *
 * Every entry point in part 1 replaces, symbol for symbol, an entry point of the reference
 * OpenBLAS 0.3.28.dev.  The file:line next to each one is the reference declaration it
 * is bound against (paths relative to the reference tree).  A caller that was compiled
 * against the reference's own cblas.h / common_interface.h links against this library
 * unchanged: types, enum values, argument order and the xerbla_ protocol are identical.
 *
 * Part 2 holds the extension entry points (device-pointer / stream / multi-GPU panel
 * helpers) used by bench.py, the SUMMA layer and the tests; they do not exist in the
 * reference.
 *
 * There is NO CPU fallback behind any of these symbols: if no CUDA device is usable the
 * call prints a diagnostic and aborts (b200_last_error() explains why).
 */
#ifndef OPENBLAS_B200_H
#define OPENBLAS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- types (openblas_config_template.h:36-45, common.h:263-279) ------------------- */
#ifndef OPENBLAS_B200_NO_TYPES
#ifdef OPENBLAS_USE64BITINT
typedef int64_t blasint;
#else
typedef int blasint;
#endif
typedef uint16_t bfloat16;

/* cblas.h:62-63 */
typedef enum CBLAS_ORDER     {CblasRowMajor = 101, CblasColMajor = 102} CBLAS_ORDER;
typedef enum CBLAS_TRANSPOSE {CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113,
                              CblasConjNoTrans = 114} CBLAS_TRANSPOSE;
typedef CBLAS_ORDER CBLAS_LAYOUT;
/* cblas.h:64-66 */
typedef enum CBLAS_UPLO      {CblasUpper = 121, CblasLower = 122} CBLAS_UPLO;
typedef enum CBLAS_SIDE      {CblasLeft = 141, CblasRight = 142} CBLAS_SIDE;
typedef enum CBLAS_DIAG      {CblasNonUnit = 131, CblasUnit = 132} CBLAS_DIAG;
#endif

/* =====================================================================================
 * Part 1 -- drop-in symbols
 * ===================================================================================== */

/* ---- CBLAS GEMM (cblas.h:298-307, cblas.h:444-445; body interface/gemm.c:294-651) --- */
void cblas_sgemm(enum CBLAS_ORDER Order, enum CBLAS_TRANSPOSE TransA, enum CBLAS_TRANSPOSE TransB,
                 blasint M, blasint N, blasint K, float alpha, const float *A, blasint lda,
                 const float *B, blasint ldb, float beta, float *C, blasint ldc);
void cblas_dgemm(enum CBLAS_ORDER Order, enum CBLAS_TRANSPOSE TransA, enum CBLAS_TRANSPOSE TransB,
                 blasint M, blasint N, blasint K, double alpha, const double *A, blasint lda,
                 const double *B, blasint ldb, double beta, double *C, blasint ldc);
void cblas_cgemm(enum CBLAS_ORDER Order, enum CBLAS_TRANSPOSE TransA, enum CBLAS_TRANSPOSE TransB,
                 blasint M, blasint N, blasint K, const void *alpha, const void *A, blasint lda,
                 const void *B, blasint ldb, const void *beta, void *C, blasint ldc);
void cblas_zgemm(enum CBLAS_ORDER Order, enum CBLAS_TRANSPOSE TransA, enum CBLAS_TRANSPOSE TransB,
                 blasint M, blasint N, blasint K, const void *alpha, const void *A, blasint lda,
                 const void *B, blasint ldb, const void *beta, void *C, blasint ldc);
void cblas_sbgemm(enum CBLAS_ORDER Order, enum CBLAS_TRANSPOSE TransA, enum CBLAS_TRANSPOSE TransB,
                  blasint M, blasint N, blasint K, float alpha, const bfloat16 *A, blasint lda,
                  const bfloat16 *B, blasint ldb, float beta, float *C, blasint ldc);

/* ---- CBLAS GEMM3M (cblas.h:304-309; interface/gemm.c built with -DGEMM3M; driver/level3/gemm3m_level3.c):
 *      every extent >= 512 (B200_3M_MIN): three real GEMMs, error bound in (|re| + |im|) gauges as ctest's 3M
 *      driver checks it; smaller products: the 4-multiply kernel ---------------------------------------------- */
void cblas_cgemm3m(enum CBLAS_ORDER Order, enum CBLAS_TRANSPOSE TransA, enum CBLAS_TRANSPOSE TransB,
                   blasint M, blasint N, blasint K, const void *alpha, const void *A, blasint lda,
                   const void *B, blasint ldb, const void *beta, void *C, blasint ldc);
void cblas_zgemm3m(enum CBLAS_ORDER Order, enum CBLAS_TRANSPOSE TransA, enum CBLAS_TRANSPOSE TransB,
                   blasint M, blasint N, blasint K, const void *alpha, const void *A, blasint lda,
                   const void *B, blasint ldb, const void *beta, void *C, blasint ldc);

/* ---- Fortran GEMM (common_interface.h:484-497; body interface/gemm.c:181-290) -------
 * All arguments by reference, trans as one char of N/T/R/C in either case; no hidden
 * string-length arguments are read. */
void sgemm_(char *TRANSA, char *TRANSB, blasint *M, blasint *N, blasint *K, float *alpha,
            float *a, blasint *ldA, float *b, blasint *ldB, float *beta, float *c, blasint *ldC);
void dgemm_(char *TRANSA, char *TRANSB, blasint *M, blasint *N, blasint *K, double *alpha,
            double *a, blasint *ldA, double *b, blasint *ldB, double *beta, double *c, blasint *ldC);
void cgemm_(char *TRANSA, char *TRANSB, blasint *M, blasint *N, blasint *K, float *alpha,
            float *a, blasint *ldA, float *b, blasint *ldB, float *beta, float *c, blasint *ldC);
void zgemm_(char *TRANSA, char *TRANSB, blasint *M, blasint *N, blasint *K, double *alpha,
            double *a, blasint *ldA, double *b, blasint *ldB, double *beta, double *c, blasint *ldC);
void sbgemm_(char *TRANSA, char *TRANSB, blasint *M, blasint *N, blasint *K, float *alpha,
             bfloat16 *a, blasint *ldA, bfloat16 *b, blasint *ldB, float *beta, float *c,
             blasint *ldC);
void cgemm3m_(char *TRANSA, char *TRANSB, blasint *M, blasint *N, blasint *K, float *alpha,
              float *a, blasint *ldA, float *b, blasint *ldB, float *beta, float *c, blasint *ldC);
void zgemm3m_(char *TRANSA, char *TRANSB, blasint *M, blasint *N, blasint *K, double *alpha,
              double *a, blasint *ldA, double *b, blasint *ldB, double *beta, double *c,
              blasint *ldC);

/* ---- batched GEMM (cblas.h:419-430,446-447; body interface/gemm_batch.c:118-372) ---- */
void cblas_sgemm_batch(enum CBLAS_ORDER Order, const enum CBLAS_TRANSPOSE *TransA_array,
                       const enum CBLAS_TRANSPOSE *TransB_array, const blasint *M_array,
                       const blasint *N_array, const blasint *K_array, const float *alpha_array,
                       const float **A_array, const blasint *lda_array, const float **B_array,
                       const blasint *ldb_array, const float *beta_array, float **C_array,
                       const blasint *ldc_array, blasint group_count, const blasint *group_size);
void cblas_dgemm_batch(enum CBLAS_ORDER Order, const enum CBLAS_TRANSPOSE *TransA_array,
                       const enum CBLAS_TRANSPOSE *TransB_array, const blasint *M_array,
                       const blasint *N_array, const blasint *K_array, const double *alpha_array,
                       const double **A_array, const blasint *lda_array, const double **B_array,
                       const blasint *ldb_array, const double *beta_array, double **C_array,
                       const blasint *ldc_array, blasint group_count, const blasint *group_size);
void cblas_cgemm_batch(enum CBLAS_ORDER Order, const enum CBLAS_TRANSPOSE *TransA_array,
                       const enum CBLAS_TRANSPOSE *TransB_array, const blasint *M_array,
                       const blasint *N_array, const blasint *K_array, const void *alpha_array,
                       const void **A_array, const blasint *lda_array, const void **B_array,
                       const blasint *ldb_array, const void *beta_array, void **C_array,
                       const blasint *ldc_array, blasint group_count, const blasint *group_size);
void cblas_zgemm_batch(enum CBLAS_ORDER Order, const enum CBLAS_TRANSPOSE *TransA_array,
                       const enum CBLAS_TRANSPOSE *TransB_array, const blasint *M_array,
                       const blasint *N_array, const blasint *K_array, const void *alpha_array,
                       const void **A_array, const blasint *lda_array, const void **B_array,
                       const blasint *ldb_array, const void *beta_array, void **C_array,
                       const blasint *ldc_array, blasint group_count, const blasint *group_size);
void cblas_sbgemm_batch(enum CBLAS_ORDER Order, const enum CBLAS_TRANSPOSE *TransA_array,
                        const enum CBLAS_TRANSPOSE *TransB_array, const blasint *M_array,
                        const blasint *N_array, const blasint *K_array, const float *alpha_array,
                        const bfloat16 **A_array, const blasint *lda_array,
                        const bfloat16 **B_array, const blasint *ldb_array,
                        const float *beta_array, float **C_array, const blasint *ldc_array,
                        blasint group_count, const blasint *group_size);

/* ---- level-3 routines that are "GEMM with a mask" (SURVEY 8(f3)): SYMM/HEMM, SYRK/HERK,
 *      SYR2K/HER2K.  CBLAS: cblas.h:320-345, 365-378; bodies interface/symm.c:231-442,
 *      interface/syrk.c:190-402, interface/syr2k.c:190-397.  Fortran: common_interface.h:554-626;
 *      bodies interface/symm.c:150-229, syrk.c:95-188, syr2k.c:95-188.  Same argument checks,
 *      xerbla_ names and info numbers as the reference; only the referenced triangle of the
 *      symmetric / Hermitian operand is read and only the named triangle of C is written. */
/* SYMM / HEMM (cblas.h:320-327, 365-368) */
void cblas_ssymm(enum CBLAS_ORDER Order, enum CBLAS_SIDE Side, enum CBLAS_UPLO Uplo, blasint M, blasint N,
                 float alpha, const float *A, blasint lda, const float *B, blasint ldb, float beta, float *C, blasint ldc);
void cblas_dsymm(enum CBLAS_ORDER Order, enum CBLAS_SIDE Side, enum CBLAS_UPLO Uplo, blasint M, blasint N,
                 double alpha, const double *A, blasint lda, const double *B, blasint ldb, double beta, double *C, blasint ldc);
void cblas_csymm(enum CBLAS_ORDER Order, enum CBLAS_SIDE Side, enum CBLAS_UPLO Uplo, blasint M, blasint N,
                 const void *alpha, const void *A, blasint lda, const void *B, blasint ldb, const void *beta, void *C, blasint ldc);
void cblas_zsymm(enum CBLAS_ORDER Order, enum CBLAS_SIDE Side, enum CBLAS_UPLO Uplo, blasint M, blasint N,
                 const void *alpha, const void *A, blasint lda, const void *B, blasint ldb, const void *beta, void *C, blasint ldc);
void cblas_chemm(enum CBLAS_ORDER Order, enum CBLAS_SIDE Side, enum CBLAS_UPLO Uplo, blasint M, blasint N,
                 const void *alpha, const void *A, blasint lda, const void *B, blasint ldb, const void *beta, void *C, blasint ldc);
void cblas_zhemm(enum CBLAS_ORDER Order, enum CBLAS_SIDE Side, enum CBLAS_UPLO Uplo, blasint M, blasint N,
                 const void *alpha, const void *A, blasint lda, const void *B, blasint ldb, const void *beta, void *C, blasint ldc);
/* SYRK / HERK (cblas.h:329-336, 370-373): HERK takes real alpha and beta */
void cblas_ssyrk(enum CBLAS_ORDER Order, enum CBLAS_UPLO Uplo, enum CBLAS_TRANSPOSE Trans, blasint N, blasint K,
                 float alpha, const float *A, blasint lda, float beta, float *C, blasint ldc);
void cblas_dsyrk(enum CBLAS_ORDER Order, enum CBLAS_UPLO Uplo, enum CBLAS_TRANSPOSE Trans, blasint N, blasint K,
                 double alpha, const double *A, blasint lda, double beta, double *C, blasint ldc);
void cblas_csyrk(enum CBLAS_ORDER Order, enum CBLAS_UPLO Uplo, enum CBLAS_TRANSPOSE Trans, blasint N, blasint K,
                 const void *alpha, const void *A, blasint lda, const void *beta, void *C, blasint ldc);
void cblas_zsyrk(enum CBLAS_ORDER Order, enum CBLAS_UPLO Uplo, enum CBLAS_TRANSPOSE Trans, blasint N, blasint K,
                 const void *alpha, const void *A, blasint lda, const void *beta, void *C, blasint ldc);
void cblas_cherk(enum CBLAS_ORDER Order, enum CBLAS_UPLO Uplo, enum CBLAS_TRANSPOSE Trans, blasint N, blasint K,
                 float alpha, const void *A, blasint lda, float beta, void *C, blasint ldc);
void cblas_zherk(enum CBLAS_ORDER Order, enum CBLAS_UPLO Uplo, enum CBLAS_TRANSPOSE Trans, blasint N, blasint K,
                 double alpha, const void *A, blasint lda, double beta, void *C, blasint ldc);
/* SYR2K / HER2K (cblas.h:338-345, 375-378): HER2K takes a complex alpha and a real beta */
void cblas_ssyr2k(enum CBLAS_ORDER Order, enum CBLAS_UPLO Uplo, enum CBLAS_TRANSPOSE Trans, blasint N, blasint K,
                  float alpha, const float *A, blasint lda, const float *B, blasint ldb, float beta, float *C, blasint ldc);
void cblas_dsyr2k(enum CBLAS_ORDER Order, enum CBLAS_UPLO Uplo, enum CBLAS_TRANSPOSE Trans, blasint N, blasint K,
                  double alpha, const double *A, blasint lda, const double *B, blasint ldb, double beta, double *C, blasint ldc);
void cblas_csyr2k(enum CBLAS_ORDER Order, enum CBLAS_UPLO Uplo, enum CBLAS_TRANSPOSE Trans, blasint N, blasint K,
                  const void *alpha, const void *A, blasint lda, const void *B, blasint ldb, const void *beta, void *C, blasint ldc);
void cblas_zsyr2k(enum CBLAS_ORDER Order, enum CBLAS_UPLO Uplo, enum CBLAS_TRANSPOSE Trans, blasint N, blasint K,
                  const void *alpha, const void *A, blasint lda, const void *B, blasint ldb, const void *beta, void *C, blasint ldc);
void cblas_cher2k(enum CBLAS_ORDER Order, enum CBLAS_UPLO Uplo, enum CBLAS_TRANSPOSE Trans, blasint N, blasint K,
                  const void *alpha, const void *A, blasint lda, const void *B, blasint ldb, float beta, void *C, blasint ldc);
void cblas_zher2k(enum CBLAS_ORDER Order, enum CBLAS_UPLO Uplo, enum CBLAS_TRANSPOSE Trans, blasint N, blasint K,
                  const void *alpha, const void *A, blasint lda, const void *B, blasint ldb, double beta, void *C, blasint ldc);
/* Fortran ABI (common_interface.h:554-626): every argument by reference, complex scalars as pointers to (re, im) */
void ssymm_(char *SIDE, char *UPLO, blasint *M, blasint *N, float *alpha, float *a, blasint *ldA, float *b, blasint *ldB,
            float *beta, float *c, blasint *ldC);
void dsymm_(char *SIDE, char *UPLO, blasint *M, blasint *N, double *alpha, double *a, blasint *ldA, double *b, blasint *ldB,
            double *beta, double *c, blasint *ldC);
void csymm_(char *SIDE, char *UPLO, blasint *M, blasint *N, float *alpha, float *a, blasint *ldA, float *b, blasint *ldB,
            float *beta, float *c, blasint *ldC);
void zsymm_(char *SIDE, char *UPLO, blasint *M, blasint *N, double *alpha, double *a, blasint *ldA, double *b, blasint *ldB,
            double *beta, double *c, blasint *ldC);
void chemm_(char *SIDE, char *UPLO, blasint *M, blasint *N, float *alpha, float *a, blasint *ldA, float *b, blasint *ldB,
            float *beta, float *c, blasint *ldC);
void zhemm_(char *SIDE, char *UPLO, blasint *M, blasint *N, double *alpha, double *a, blasint *ldA, double *b, blasint *ldB,
            double *beta, double *c, blasint *ldC);
void ssyrk_(char *UPLO, char *TRANS, blasint *N, blasint *K, float *alpha, float *a, blasint *ldA, float *beta, float *c,
            blasint *ldC);
void dsyrk_(char *UPLO, char *TRANS, blasint *N, blasint *K, double *alpha, double *a, blasint *ldA, double *beta, double *c,
            blasint *ldC);
void csyrk_(char *UPLO, char *TRANS, blasint *N, blasint *K, float *alpha, float *a, blasint *ldA, float *beta, float *c,
            blasint *ldC);
void zsyrk_(char *UPLO, char *TRANS, blasint *N, blasint *K, double *alpha, double *a, blasint *ldA, double *beta, double *c,
            blasint *ldC);
void cherk_(char *UPLO, char *TRANS, blasint *N, blasint *K, float *alpha, float *a, blasint *ldA, float *beta, float *c,
            blasint *ldC);
void zherk_(char *UPLO, char *TRANS, blasint *N, blasint *K, double *alpha, double *a, blasint *ldA, double *beta, double *c,
            blasint *ldC);
void ssyr2k_(char *UPLO, char *TRANS, blasint *N, blasint *K, float *alpha, float *a, blasint *ldA, float *b, blasint *ldB,
             float *beta, float *c, blasint *ldC);
void dsyr2k_(char *UPLO, char *TRANS, blasint *N, blasint *K, double *alpha, double *a, blasint *ldA, double *b, blasint *ldB,
             double *beta, double *c, blasint *ldC);
void csyr2k_(char *UPLO, char *TRANS, blasint *N, blasint *K, float *alpha, float *a, blasint *ldA, float *b, blasint *ldB,
             float *beta, float *c, blasint *ldC);
void zsyr2k_(char *UPLO, char *TRANS, blasint *N, blasint *K, double *alpha, double *a, blasint *ldA, double *b, blasint *ldB,
             double *beta, double *c, blasint *ldC);
void cher2k_(char *UPLO, char *TRANS, blasint *N, blasint *K, float *alpha, float *a, blasint *ldA, float *b, blasint *ldB,
             float *beta, float *c, blasint *ldC);
void zher2k_(char *UPLO, char *TRANS, blasint *N, blasint *K, double *alpha, double *a, blasint *ldA, double *b, blasint *ldB,
             double *beta, double *c, blasint *ldC);

/* ---- TRMM / TRSM (cblas.h:347-363; common_interface.h:530-552; bodies interface/trsm.c:96-430, which the
 *      reference compiles twice, with and without -DTRMM).  B := alpha*op(A)*B or alpha*B*op(A) (TRMM), or the
 *      solution X of op(A)*X = alpha*B or X*op(A) = alpha*B (TRSM); A triangular, only its uplo triangle is read
 *      and a unit diagonal is never read. */
void cblas_strmm(enum CBLAS_ORDER Order, enum CBLAS_SIDE Side, enum CBLAS_UPLO Uplo, enum CBLAS_TRANSPOSE TransA,
                 enum CBLAS_DIAG Diag, blasint M, blasint N, float alpha, const float *A, blasint lda, float *B, blasint ldb);
void cblas_dtrmm(enum CBLAS_ORDER Order, enum CBLAS_SIDE Side, enum CBLAS_UPLO Uplo, enum CBLAS_TRANSPOSE TransA,
                 enum CBLAS_DIAG Diag, blasint M, blasint N, double alpha, const double *A, blasint lda, double *B, blasint ldb);
void cblas_ctrmm(enum CBLAS_ORDER Order, enum CBLAS_SIDE Side, enum CBLAS_UPLO Uplo, enum CBLAS_TRANSPOSE TransA,
                 enum CBLAS_DIAG Diag, blasint M, blasint N, const void *alpha, const void *A, blasint lda, void *B, blasint ldb);
void cblas_ztrmm(enum CBLAS_ORDER Order, enum CBLAS_SIDE Side, enum CBLAS_UPLO Uplo, enum CBLAS_TRANSPOSE TransA,
                 enum CBLAS_DIAG Diag, blasint M, blasint N, const void *alpha, const void *A, blasint lda, void *B, blasint ldb);
void cblas_strsm(enum CBLAS_ORDER Order, enum CBLAS_SIDE Side, enum CBLAS_UPLO Uplo, enum CBLAS_TRANSPOSE TransA,
                 enum CBLAS_DIAG Diag, blasint M, blasint N, float alpha, const float *A, blasint lda, float *B, blasint ldb);
void cblas_dtrsm(enum CBLAS_ORDER Order, enum CBLAS_SIDE Side, enum CBLAS_UPLO Uplo, enum CBLAS_TRANSPOSE TransA,
                 enum CBLAS_DIAG Diag, blasint M, blasint N, double alpha, const double *A, blasint lda, double *B, blasint ldb);
void cblas_ctrsm(enum CBLAS_ORDER Order, enum CBLAS_SIDE Side, enum CBLAS_UPLO Uplo, enum CBLAS_TRANSPOSE TransA,
                 enum CBLAS_DIAG Diag, blasint M, blasint N, const void *alpha, const void *A, blasint lda, void *B, blasint ldb);
void cblas_ztrsm(enum CBLAS_ORDER Order, enum CBLAS_SIDE Side, enum CBLAS_UPLO Uplo, enum CBLAS_TRANSPOSE TransA,
                 enum CBLAS_DIAG Diag, blasint M, blasint N, const void *alpha, const void *A, blasint lda, void *B, blasint ldb);
void strmm_(char *SIDE, char *UPLO, char *TRANSA, char *DIAG, blasint *M, blasint *N, float *alpha, float *a, blasint *ldA,
            float *b, blasint *ldB);
void dtrmm_(char *SIDE, char *UPLO, char *TRANSA, char *DIAG, blasint *M, blasint *N, double *alpha, double *a, blasint *ldA,
            double *b, blasint *ldB);
void ctrmm_(char *SIDE, char *UPLO, char *TRANSA, char *DIAG, blasint *M, blasint *N, float *alpha, float *a, blasint *ldA,
            float *b, blasint *ldB);
void ztrmm_(char *SIDE, char *UPLO, char *TRANSA, char *DIAG, blasint *M, blasint *N, double *alpha, double *a, blasint *ldA,
            double *b, blasint *ldB);
void strsm_(char *SIDE, char *UPLO, char *TRANSA, char *DIAG, blasint *M, blasint *N, float *alpha, float *a, blasint *ldA,
            float *b, blasint *ldB);
void dtrsm_(char *SIDE, char *UPLO, char *TRANSA, char *DIAG, blasint *M, blasint *N, double *alpha, double *a, blasint *ldA,
            double *b, blasint *ldB);
void ctrsm_(char *SIDE, char *UPLO, char *TRANSA, char *DIAG, blasint *M, blasint *N, float *alpha, float *a, blasint *ldA,
            float *b, blasint *ldB);
void ztrsm_(char *SIDE, char *UPLO, char *TRANSA, char *DIAG, blasint *M, blasint *N, double *alpha, double *a, blasint *ldA,
            double *b, blasint *ldB);

/* ---- ?GEMMT: the Uplo triangle of the M x M matrix C := alpha op(A) op(B) + beta C; the other
 *      triangle of C is never read or written (cblas.h:311-318; common_interface.h:506-513;
 *      interface/gemmt.c:67 Fortran, :200 CBLAS).  One triangle-masked launch of the GEMM kernels.  A and B are
 *      never written (the reference conjugates one operand in place for ConjTrans / ConjNoTrans, gemmt.c:466-476). */
void cblas_sgemmt(enum CBLAS_ORDER Order, enum CBLAS_UPLO Uplo, enum CBLAS_TRANSPOSE TransA, enum CBLAS_TRANSPOSE TransB,
                  blasint M, blasint K, float alpha, const float *A, blasint lda, const float *B, blasint ldb, float beta,
                  float *C, blasint ldc);
void cblas_dgemmt(enum CBLAS_ORDER Order, enum CBLAS_UPLO Uplo, enum CBLAS_TRANSPOSE TransA, enum CBLAS_TRANSPOSE TransB,
                  blasint M, blasint K, double alpha, const double *A, blasint lda, const double *B, blasint ldb, double beta,
                  double *C, blasint ldc);
void cblas_cgemmt(enum CBLAS_ORDER Order, enum CBLAS_UPLO Uplo, enum CBLAS_TRANSPOSE TransA, enum CBLAS_TRANSPOSE TransB,
                  blasint M, blasint K, const void *alpha, const void *A, blasint lda, const void *B, blasint ldb,
                  const void *beta, void *C, blasint ldc);
void cblas_zgemmt(enum CBLAS_ORDER Order, enum CBLAS_UPLO Uplo, enum CBLAS_TRANSPOSE TransA, enum CBLAS_TRANSPOSE TransB,
                  blasint M, blasint K, const void *alpha, const void *A, blasint lda, const void *B, blasint ldb,
                  const void *beta, void *C, blasint ldc);
void sgemmt_(char *UPLO, char *TRANSA, char *TRANSB, blasint *M, blasint *K, float *alpha, float *a, blasint *ldA, float *b,
             blasint *ldB, float *beta, float *c, blasint *ldC);
void dgemmt_(char *UPLO, char *TRANSA, char *TRANSB, blasint *M, blasint *K, double *alpha, double *a, blasint *ldA, double *b,
             blasint *ldB, double *beta, double *c, blasint *ldC);
void cgemmt_(char *UPLO, char *TRANSA, char *TRANSB, blasint *M, blasint *K, float *alpha, float *a, blasint *ldA, float *b,
             blasint *ldB, float *beta, float *c, blasint *ldC);
void zgemmt_(char *UPLO, char *TRANSA, char *TRANSB, blasint *M, blasint *K, double *alpha, double *a, blasint *ldA, double *b,
             blasint *ldB, double *beta, double *c, blasint *ldC);

/* ---- SBGEMMT: ?GEMMT with bf16 A and B, fp32 alpha, beta and C (interface/sbgemmt.c:47-52 Fortran,
 *      :143-149 CBLAS; built under BUILD_BFLOAT16, interface/Makefile:52,290,1306,1970; the reference declares it in
 *      no public header).  One triangle-masked launch of the tcgen05 SBGEMM kernel.  As in the reference: K == 0
 *      leaves C untouched whatever beta is, and a RowMajor call does NOT flip Uplo (sbgemmt.c:239-240).  One
 *      deviation: with alpha == 0 the reference still forms 0 * (A B) in its SBGEMV kernels, so NaN / Inf in A or B
 *      reach C; here alpha == 0 follows the BLAS convention (A and B are not read, the triangle is scaled by beta). */
void cblas_sbgemmt(enum CBLAS_ORDER Order, enum CBLAS_UPLO Uplo, enum CBLAS_TRANSPOSE TransA, enum CBLAS_TRANSPOSE TransB,
                   blasint M, blasint K, float alpha, const bfloat16 *A, blasint lda, const bfloat16 *B, blasint ldb,
                   float beta, float *C, blasint ldc);
void sbgemmt_(char *UPLO, char *TRANSA, char *TRANSB, blasint *M, blasint *K, float *alpha, bfloat16 *a, blasint *ldA,
              bfloat16 *b, blasint *ldB, float *beta, float *c, blasint *ldC);

/* ---- SBGEMV / SBDOT: bf16 operands, fp32 accumulation and result (cblas.h:441-442;
 *      common_interface.h:62,258; interface/sbgemv.c, interface/bf16dot.c, kernel/x86_64/sbgemv_n.c, sbgemv_t.c,
 *      sbdot.c).  HBM-bound CUDA kernels (csrc/bf16_level12.cu), deterministic (no atomics). */
float cblas_sbdot(blasint n, const bfloat16 *x, blasint incx, const bfloat16 *y, blasint incy);
void  cblas_sbgemv(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE trans, blasint m, blasint n, float alpha, const bfloat16 *a,
                   blasint lda, const bfloat16 *x, blasint incx, float beta, float *y, blasint incy);
float sbdot_(blasint *N, bfloat16 *x, blasint *INCX, bfloat16 *y, blasint *INCY);
void  sbgemv_(char *TRANS, blasint *M, blasint *N, float *ALPHA, bfloat16 *a, blasint *LDA, bfloat16 *x, blasint *INCX,
              float *BETA, float *y, blasint *INCY);

/* ---- bf16 conversion helpers callers of sbgemm need (cblas.h:433-440;
 *      interface/tobf16.c, interface/bf16to.c; rounding rule kernel/x86_64/tobf16.c:46-96) */
void cblas_sbstobf16(blasint n, const float *in, blasint incin, bfloat16 *out, blasint incout);
void cblas_sbdtobf16(blasint n, const double *in, blasint incin, bfloat16 *out, blasint incout);
void cblas_sbf16tos(blasint n, const bfloat16 *in, blasint incin, float *out, blasint incout);
void cblas_dbf16tod(blasint n, const bfloat16 *in, blasint incin, double *out, blasint incout);
void sbstobf16_(blasint *n, float *in, blasint *incin, bfloat16 *out, blasint *incout);
void sbdtobf16_(blasint *n, double *in, blasint *incin, bfloat16 *out, blasint *incout);
void sbf16tos_(blasint *n, bfloat16 *in, blasint *incin, float *out, blasint *incout);
void dbf16tod_(blasint *n, bfloat16 *in, blasint *incin, double *out, blasint *incout);

/* ---- multi-GPU: 2-D block-cyclic SUMMA over a P x Q grid of processes, one GPU each (csrc/summa.cu).
 *      Replaces the reference's threaded level-3 driver: driver/level3/level3_thread.c:219-532 (inner_thread: every
 *      worker packs its share once, announces it through job[].working flags, the others consume it, double
 *      buffered), :804-862 (choice of the nthreads_m x nthreads_n grid), gemm_thread_mn.c:43-61 (divide_rule[]).
 *      Panels travel by copy-engine pulls from the owners' windows (CUDA IPC over NVLink) announced with stream
 *      memory operations -- no SM is taken from the local GEMM; B200_SUMMA_TRANSPORT=nccl switches to ncclBroadcast
 *      on row / column communicators.  NCCL (dlopen'ed) bootstraps the group from the 128-byte id every rank gets
 *      from rank 0 by whatever means the caller has (MPI_Bcast, torch.distributed, a file).  All calls return 0 or
 *      non-zero with b200_last_error() set; create / gemm / destroy are collective. ----------------------------- */
typedef struct b200_summa b200_summa;
int      b200_summa_unique_id(void *id128);
void     b200_summa_grid(int world, int *P, int *Q);                       /* 2 -> 1x2, 4 -> 2x2, 8 -> 2x4 (divide_rule[]) */
int      b200_summa_create(b200_summa **handle, const void *id128, int rank, int world, int P, int Q);
int      b200_summa_destroy(b200_summa *handle);
int64_t  b200_summa_numroc(int64_t n, int64_t nb, int iproc, int nprocs);  /* local extent of a block-cyclic dimension */
int64_t  b200_summa_schedule(int64_t k, int64_t nb, int P, int Q, int64_t capacity, int64_t *k0, int64_t *width, int *a_owner_col,
                             int64_t *a_local_col, int *b_owner_row, int64_t *b_local_row);
/* C := alpha A B + beta C, global m x n x k, block nb; this rank (p = rank / Q, q = rank % Q) passes its local pieces
 * a_loc (numroc(m,nb,p,P) x numroc(k,nb,q,Q)), b_loc (numroc(k,nb,p,P) x numroc(n,nb,q,Q)), c_loc (numroc(m,nb,p,P) x
 * numroc(n,nb,q,Q)), column-major, device or host memory; the local products run on `stream` (cudaStream_t). */
int      b200_summa_gemm(b200_summa *handle, int dtype, int64_t m, int64_t n, int64_t k, int64_t nb, const void *alpha,
                         const void *a_loc, int64_t lda, const void *b_loc, int64_t ldb, const void *beta, void *c_loc,
                         int64_t ldc, void *stream);
uint64_t b200_summa_launches(const b200_summa *handle);
const char *b200_summa_describe(const b200_summa *handle);

/* ---- error hook (driver/others/xerbla.c:56-73): weak, a caller's own xerbla_ wins ---- */
int xerbla_(char *name, blasint *info, blasint len);

/* ---- control API kept for callers (cblas.h:13-45).  Only exported by the stand-alone
 *      library; the co-link build (libopenblas_b200_gemmonly.so) leaves them to
 *      libopenblas.a (SURVEY 8(b) "Link recipe"). "threads" maps to nothing on a GPU:
 *      set is accepted and remembered, get returns it. ------------------------------- */
void  openblas_set_num_threads(int num_threads);
void  goto_set_num_threads(int num_threads);
int   openblas_get_num_threads(void);
int   openblas_get_num_procs(void);
char *openblas_get_config(void);
char *openblas_get_corename(void);
int   openblas_get_parallel(void);

/* =====================================================================================
 * Part 2 -- extensions (not in the reference)
 * ===================================================================================== */

/* precision codes used by the generic entry points below */
enum b200_dtype { B200_S = 0, B200_D = 1, B200_C = 2, B200_Z = 3, B200_SB = 4 };
/* op codes after normalisation (interface/gemm.c:245-269): N, T, R = conj, C = conj-trans */
enum b200_trans { B200_N = 0, B200_T = 1, B200_R = 2, B200_C_ = 3 };

/* kernel families; b200_set_kernel(B200_K_AUTO) restores the size-based choice */
enum b200_kernel {
  B200_K_AUTO = 0,
  B200_K_GENERIC = 1, /* any shape/ld/alignment, all precisions: latency path          */
  B200_K_FAST = 2     /* DMMA (d,z) / FFMA (s,c) / tcgen05 (sb) roofline kernels        */
};

/* Column-major C <- alpha*op(A)*op(B) + beta*C on DEVICE pointers, enqueued on `stream`
 * (a cudaStream_t passed as void*; NULL = the legacy default stream, as everywhere in CUDA)
 * without synchronising.  alpha/beta are HOST pointers to one FLOAT (s,d,sb) or two (c,z).
 * Arguments are assumed valid (the BLAS entry points validate).  Returns 0 or a
 * cudaError_t value. */
int b200_gemm_async(int dtype, int transa, int transb, int64_t m, int64_t n, int64_t k,
                    const void *alpha, const void *A, int64_t lda, const void *B, int64_t ldb,
                    const void *beta, void *C, int64_t ldc, void *stream);

/* Same operation, synchronous, pointers may each be host (pageable or pinned) or device.
 * This is what every Part-1 entry point calls after validation. */
int b200_gemm(int dtype, int transa, int transb, int64_t m, int64_t n, int64_t k,
              const void *alpha, const void *A, int64_t lda, const void *B, int64_t ldb,
              const void *beta, void *C, int64_t ldc);

/* Force a kernel family for subsequent calls of this process (tests, profiling). */
void b200_set_kernel(int kernel);
int  b200_get_kernel(void);

/* Number of kernels this library has launched since load (bench.py "gpu_launches"). */
uint64_t b200_launch_count(void);

/* Name of the kernel the last b200_gemm*() of the calling thread dispatched to. */
const char *b200_last_kernel(void);

/* Last error text of the calling thread ("" if none). */
const char *b200_last_error(void);

/* Initialise eagerly on `device` (otherwise lazily on the current device at first call).
 * Returns 0 on success. */
int b200_init(int device);

/* Pinned host memory helpers (what a caller should hand to the BLAS symbols to get
 * full-rate DMA; pageable memory also works, through a staging ring). */
void *b200_host_alloc(size_t bytes);
void  b200_host_free(void *p);

/* Releases everything the library holds on the GPU and the host (pooled streams, events, device workspace,
 * pinned staging, the host copy threads); also runs as the library's destructor.  Replaces gotoblas_quit
 * (driver/others/memory.c:1566-1600).  The next call initialises again. */
void b200_shutdown(void);

/* Version string, e.g. "openblas_b200 0.1 (sm_100a)". */
const char *b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* OPENBLAS_B200_H */
