/*
 * xerbla.c -- default error hook.  Same contract as the reference's
 * driver/others/xerbla.c:56-73: print one line, return 0, never abort; exported WEAK so that
 * a program's (or the ctest harness's, ctest/c_xerbla.c:131-135) own xerbla_ overrides it.
 */
#include <stdio.h>
#include "shim.h"

B200_EXPORT __attribute__((weak)) int xerbla_(char *name, blasint *info, blasint len) {
  (void)len;
  printf(" ** On entry to %6s parameter number %2d had an illegal value\n", name, (int)*info);
  return 0;
}
