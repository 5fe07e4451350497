/*
 * control.c -- the run-time control symbols callers of the reference expect (cblas.h:13-45;
 * reference bodies: driver/others/blas_server.c:864,936, memory.c:503-511,
 * openblas_get_config.c, openblas_get_parallel.c).  On a GPU "threads" have no meaning:
 * the setters remember the value, the getters return it.  This file is left out of the
 * co-link build (libopenblas_b200_gemmonly.so) so that libopenblas.a can supply them.
 */
#include <unistd.h>
#include "shim.h"

static int requested_threads = 1;

B200_EXPORT void openblas_set_num_threads(int n) { if (n > 0) requested_threads = n; }
B200_EXPORT void goto_set_num_threads(int n) { openblas_set_num_threads(n); }
B200_EXPORT int openblas_get_num_threads(void) { return requested_threads; }
B200_EXPORT int openblas_get_num_procs(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}
B200_EXPORT char *openblas_get_config(void) {
  return (char *)"OpenBLAS-B200 0.1 (ABI of OpenBLAS 0.3.28.dev) NO_LAPACK NO_AFFINITY "
                 "CUDA sm_100a GEMM-only";
}
B200_EXPORT char *openblas_get_corename(void) { return (char *)"B200"; }
/* 0 = OPENBLAS_SEQUENTIAL (cblas.h:40-45): the host side never spawns threads. */
B200_EXPORT int openblas_get_parallel(void) { return 0; }
