/* tests/c_abi_example.c -- a plain C99 host calling the C ABI exactly as INTEGRATION.md section B shows.
 * Built with gcc (no C++, no CUDA toolkit needed on the caller's side) and run by tests/test_gpu_compat.py.
 * Exit code 0 = the solve matched the expectations printed below. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "../include/lbfgsb200.h"

int main(void)
{
    const size_t n = 10000;
    double *x0 = (double *)malloc(n * sizeof(double)), *x = (double *)malloc(n * sizeof(double));
    lbfgsb200_params_t p;
    lbfgsb200_result_t r;
    size_t i;
    int rc;
    double worst = 0.0;
    /* sequential-implementation/main.cpp as shipped: separable quadratic, x0 ~ U(-1000,1000), tol 1e-8 */
    lbfgsb200_x0_uniform(42, -1000.0, 1000.0, 0, n, x0);
    lbfgsb200_params_default(&p, LBFGSB200_FLAVOR_SEQ);
    p.line_search = LBFGSB200_LS_BACKTRACKING;
    p.max_iterations = 15000;
    p.m = 10;
    p.tolerance = 1e-8;
    rc = lbfgsb200_solve(LBFGSB200_OBJ_QUADRATIC, n, x0, x, &p, &r, NULL, 0);
    if (rc < 0) {
        fprintf(stderr, "%s: %s\n", lbfgsb200_strerror(rc), lbfgsb200_last_error());
        return 10;
    }
    for (i = 0; i < n; ++i)
        if (fabs(x[i] - 1.0) > worst) worst = fabs(x[i] - 1.0);
    printf("status=%d (%s) iterations=%lld f=%g |g|=%g max|x-1|=%g launches=%lld\n", rc, lbfgsb200_strerror(rc),
           (long long)r.iterations, r.f, r.gnorm, worst, (long long)r.kernel_launches);
    if (rc != LBFGSB200_CONVERGED || r.iterations != 2 || worst > 1e-9) return 11; /* the reference converges at k=2 */
    /* unknown line search => the reference's invalid_argument */
    p.line_search = 9;
    rc = lbfgsb200_solve(LBFGSB200_OBJ_QUADRATIC, n, x0, x, &p, &r, NULL, 0);
    if (rc != LBFGSB200_ERR_INVALID) return 12;
    printf("rejected: %s\n", lbfgsb200_last_error());
    free(x0);
    free(x);
    return 0;
}
