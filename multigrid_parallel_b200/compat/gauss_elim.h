/*
 * gauss_elim.h -- drop-in for the reference's gauss_elim.h on top of libmgb.
 * The dense LU factorisation and the triangular solves run on the GPU
 * (mgb_host_lu_factor / mgb_host_lu_solve) in the reference's exact operation
 * order, so the factors and solutions are bit-identical.
 * gaussianElimination (reference gauss_elim.h:65-97) belongs to the 1D
 * prototypes and is not part of this path.
 */
#ifndef GAUSS_ELIM_H
#define GAUSS_ELIM_H

#include <assert.h>
#include <stdio.h>

#include "mgb.h"

#ifndef MGB_COMPAT_CHECK_DEFINED
#define MGB_COMPAT_CHECK_DEFINED
static void mgb_compat_check(int rc, const char *what)
{
    if (rc != 0) {
        fprintf(stderr, "libmgb: %s failed: %s\n", what, mgb_last_error());
        assert(rc == 0 && "libmgb call failed (there is no CPU fallback)");
        abort();
    }
}
#endif

/* reference gauss_elim.h:9-29 */
void convertToLU_InPlace(double *a, int n)
{
    mgb_compat_check(mgb_host_lu_factor(a, n), "mgb_host_lu_factor");
}

/* reference gauss_elim.h:31-60 */
void solveWithLU(const double *__restrict__ LU, const int n, const double *__restrict__ b,
                 double *__restrict__ x)
{
    mgb_compat_check(mgb_host_lu_solve(LU, n, b, x), "mgb_host_lu_solve");
}

#endif
