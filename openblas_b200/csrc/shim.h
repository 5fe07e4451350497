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
