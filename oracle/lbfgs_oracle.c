/*
 * oracle/lbfgs_oracle.c -- CPU restatement of the reference L-BFGS hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see lbfgs_oracle.h).  Plain C, single thread, the
 * reference's operation order kept expression by expression so that, built
 * without FMA contraction (-ffp-contract=off), every intermediate is
 * bit-identical to the unmodified reference compiled for x86-64
 * (oracle/_ref/, checked by tests/test_oracle.py and the golden traces).
 *
 * seq/ = /root/reference/sequential-implementation/
 * par/ = /root/reference/parallel-implementation/
 */
#include "lbfgs_oracle.h"

#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* vector utilities: seq/vector_utils.cpp:32-86                        */
/* ------------------------------------------------------------------ */

/* TEST-ONLY switch, used to EXPLAIN a golden file, never to define parity: with it on, every reduction
 * (dot products, norms, the sums inside f) is accumulated in long double (64-bit mantissa) and rounded to double
 * once, instead of the reference's naive left-to-right double sum.  Each term is still the double the reference
 * forms.  The result is the trajectory the reference's algorithm takes when its sums carry (practically) no
 * rounding error: the distance between the two trajectories is the reference's OWN summation noise, which at
 * n = 1e7 reaches 1e-10 relative in the iterates after 20 steps (tests/golden/reference_large.json, "exact_sums").
 * The product has no such mode; its deterministic tree sums are accurate to a few ulp by construction. */
static int g_exact_sums = 0;
void oracle_set_exact_sums(int on) { g_exact_sums = on; }

/* seq/vector_utils.cpp:32-41 -- naive left-to-right sum */
double oracle_dot(const double *a, const double *b, size_t n)
{
    if (g_exact_sums) {
        long double acc = 0.0L;
        for (size_t i = 0; i < n; ++i)
            acc += (long double)(a[i] * b[i]);
        return (double)acc;
    }
    double sum = 0.;
    for (size_t i = 0; i < n; ++i)
        sum += a[i] * b[i];
    return sum;
}

/* seq/vector_utils.cpp:78-86 */
double oracle_norm(const double *a, size_t n)
{
    if (g_exact_sums) {
        long double acc = 0.0L;
        for (size_t i = 0; i < n; ++i)
            acc += (long double)(a[i] * a[i]);
        return sqrt((double)acc);
    }
    double r = 0.;
    for (size_t i = 0; i < n; ++i)
        r += a[i] * a[i];
    return sqrt(r);
}

/* ------------------------------------------------------------------ */
/* objectives                                                          */
/* ------------------------------------------------------------------ */

/* par/functions.cpp:6-14 */
static double quadratic_f(const double *x, size_t n)
{
    if (g_exact_sums) {
        long double acc = 0.0L;
        for (size_t i = 0; i < n; ++i)
            acc += (long double)((x[i] - 1) * (x[i] - 1));
        return (double)acc;
    }
    double sum = 0.0;
    for (size_t i = 0; i < n; ++i)
        sum += (x[i] - 1) * (x[i] - 1);
    return sum;
}

/* par/functions.cpp:16-24 */
static void quadratic_g(const double *x, double *g, size_t n)
{
    for (size_t i = 0; i < n; ++i)
        g[i] = 2.0 * (x[i] - 1);
}

/* par/functions.cpp:26-36 (same as seq/benchmark.cpp:58-68) */
static double rosenbrock_f(const double *x, size_t n)
{
    if (g_exact_sums) {
        long double acc = 0.0L;
        for (size_t i = 0; i + 1 < n; ++i) {
            double term1 = x[i + 1] - x[i] * x[i];
            double term2 = 1 - x[i];
            acc += (long double)(100.0 * term1 * term1 + term2 * term2);
        }
        return (double)acc;
    }
    double sum = 0.0;
    for (size_t i = 0; i + 1 < n; ++i) {
        double term1 = x[i + 1] - x[i] * x[i];
        double term2 = 1 - x[i];
        sum += 100.0 * term1 * term1 + term2 * term2;
    }
    return sum;
}

/* par/functions.cpp:38-49 (same as seq/benchmark.cpp:70-81) */
static void rosenbrock_g(const double *x, double *g, size_t n)
{
    for (size_t i = 0; i < n; ++i)
        g[i] = 0.0;
    for (size_t i = 0; i + 1 < n; ++i) {
        double term1 = 2.0 * (x[i] - 1);
        double term2 = x[i + 1] - x[i] * x[i];
        g[i] += term1 - 400.0 * x[i] * term2;
        g[i + 1] += 200.0 * term2;
    }
}

/* seq/benchmark.cpp:16-34, COEFFICIENT = 1000.0 (:13) */
static double tridiag_f(const double *x, size_t n)
{
    const double COEFFICIENT = 1000.0;
    if (g_exact_sums) {
        long double acc = 0.0L;
        for (size_t i = 0; i < n; ++i)
            acc += (long double)(COEFFICIENT * x[i] * x[i]);
        for (size_t i = 0; i + 1 < n; ++i)
            acc += (long double)((COEFFICIENT / 10.0) * x[i] * x[i + 1]);
        return (double)acc;
    }
    double result = 0.0;
    for (size_t i = 0; i < n; ++i)
        result += COEFFICIENT * x[i] * x[i];
    for (size_t i = 0; i + 1 < n; ++i)
        result += (COEFFICIENT / 10.0) * x[i] * x[i + 1];
    return result;
}

/* seq/benchmark.cpp:37-56 */
static void tridiag_g(const double *x, double *g, size_t n)
{
    const double COEFFICIENT = 1000.0;
    for (size_t i = 0; i < n; ++i)
        g[i] = 2.0 * COEFFICIENT * x[i];
    for (size_t i = 0; i + 1 < n; ++i) {
        g[i] += (COEFFICIENT / 10.0) * x[i + 1];
        g[i + 1] += (COEFFICIENT / 10.0) * x[i];
    }
}

double oracle_f(int objective, const double *x, size_t n)
{
    switch (objective) {
    case ORACLE_OBJ_QUADRATIC: return quadratic_f(x, n);
    case ORACLE_OBJ_ROSENBROCK: return rosenbrock_f(x, n);
    default: return tridiag_f(x, n);
    }
}

void oracle_grad(int objective, const double *x, double *g, size_t n)
{
    switch (objective) {
    case ORACLE_OBJ_QUADRATIC: quadratic_g(x, g, n); break;
    case ORACLE_OBJ_ROSENBROCK: rosenbrock_g(x, g, n); break;
    default: tridiag_g(x, g, n); break;
    }
}

/* ------------------------------------------------------------------ */
/* interpolation helpers                                               */
/* ------------------------------------------------------------------ */

/* seq/line_search.cpp:8-12 == par/line_search.cpp:10-15 */
double oracle_cubic(double alpha0, double alpha1, double phi0, double dphi0, double phi1,
                    double dphi1)
{
    double d1 = dphi0 + dphi1 - 3 * (phi1 - phi0) / (alpha1 - alpha0);
    double d2 = copysign(sqrt(d1 * d1 - dphi0 * dphi1), alpha1 - alpha0);
    return alpha0 + (alpha1 - alpha0) * (dphi0 + d2 - d1) / (dphi0 - dphi1 + 2 * d2);
}

/* seq/line_search.cpp:14-16 == par/line_search.cpp:17-20 (alpha1 is unused there too) */
double oracle_quadratic(double alpha0, double alpha1, double phi0, double dphi0, double phi1)
{
    (void)alpha1;
    return alpha0 - 0.5 * dphi0 * alpha0 * alpha0 / (phi1 - phi0 - dphi0 * alpha0);
}

/* par/line_search.cpp:231-296.  The try/catch blocks there guard plain double
 * arithmetic, which never throws; they are no-ops and are not restated. */
double oracle_safe_cubic(double alpha0, double alpha1, double phi0, double dphi0, double phi1,
                         double dphi1)
{
    if (alpha0 > alpha1) {
        double t;
        t = alpha0; alpha0 = alpha1; alpha1 = t;
        t = phi0; phi0 = phi1; phi1 = t;
        t = dphi0; dphi0 = dphi1; dphi1 = t;
    }
    double d1 = dphi0 + dphi1 - 3 * (phi1 - phi0) / (alpha1 - alpha0);
    if (isnan(d1) || isinf(d1))
        return 0.5 * (alpha0 + alpha1);
    double discriminant = d1 * d1 - dphi0 * dphi1;
    if (discriminant < 0)
        return 0.5 * (alpha0 + alpha1);
    double d2 = copysign(sqrt(discriminant), alpha1 - alpha0);
    double denominator = dphi0 - dphi1 + 2 * d2;
    if (fabs(denominator) < 1e-10)
        return 0.5 * (alpha0 + alpha1);
    double result = alpha0 + (alpha1 - alpha0) * (dphi0 + d2 - d1) / denominator;
    if (isnan(result) || isinf(result))
        return 0.5 * (alpha0 + alpha1);
    /* std::max(lo, std::min(hi, result)) */
    double hi = alpha1 - 0.1 * (alpha1 - alpha0);
    double lo = alpha0 + 0.1 * (alpha1 - alpha0);
    double mn = (result < hi) ? result : hi; /* std::min(hi,result): returns hi unless result<hi */
    return (lo < mn) ? mn : lo;             /* std::max(lo,mn): returns lo unless lo<mn */
}

/* ------------------------------------------------------------------ */
/* line searches over an abstract phi(alpha)                           */
/* ------------------------------------------------------------------ */

typedef struct phi_s {
    /* phi(alpha) = f(x + alpha d) ; dphi(alpha) = grad f(x + alpha d) . d */
    double (*f_at)(struct phi_s *, double alpha);
    double (*df_at)(struct phi_s *, double alpha);
    double (*f0)(struct phi_s *); /* f(x): the reference re-evaluates it, counted as an f call */
    double gd;                    /* dotProduct(gradient, d) */
    long nf, ng, ntrial;          /* ntrial = distinct trial points = f_at calls */
    void *ctx;
} phi_t;

/* constants: seq/config.h:5-17 and par/constants.h:5-21 */
static const double C1 = 1e-4;
static const double INITIAL_STEP_SIZE = 1.0;
static const double BACKTRACKING_ALPHA = 0.5;
static const double BACKTRACKING_TOL = 1e-8;
static const double WOLFE_INTERP_MIN = 1e-10;
static double c2_of(int flavor) { return flavor != ORACLE_FLAVOR_SEQ ? 0.7 : 0.9; }

/* seq/line_search.cpp:19-30 ; par/line_search.cpp:25-43 (adds the 0.5 floor) */
static double ls_backtracking(phi_t *p, int flavor)
{
    double alpha = INITIAL_STEP_SIZE;
    /* note the reference's test: f(x) - f(x+alpha d) < C1*alpha*(g.d) */
    while (p->f0(p) - p->f_at(p, alpha) < C1 * alpha * p->gd) {
        alpha *= BACKTRACKING_ALPHA;
        if (alpha < BACKTRACKING_TOL)
            break;
    }
    if (flavor == ORACLE_FLAVOR_PAR && alpha < 1e-4)
        return 0.5; /* par/line_search.cpp:38-41 */
    return alpha;
}

/* seq/line_search.cpp:57-121 ; par/line_search.cpp:156-228 */
static double ls_interpolation(phi_t *p, int flavor)
{
    const double f_x = p->f0(p);
    const double grad_dot_d = p->gd;
    double alpha = INITIAL_STEP_SIZE;
    double alpha_prev = 0.0;
    double f_prev = f_x;
    int iteration = 0;
    const int max_iterations = 20;
    while (iteration++ < max_iterations) {
        double f_new = p->f_at(p, alpha);
        if (f_new <= f_x + C1 * alpha * grad_dot_d)
            return alpha;
        if (alpha < WOLFE_INTERP_MIN)
            return WOLFE_INTERP_MIN;
        if (alpha_prev > 0) {
            double delta_alpha = alpha - alpha_prev;
            if (fabs(delta_alpha) < 1e-10) {
                alpha *= 0.5;
            } else {
                double grad_alpha = (f_new - f_x - grad_dot_d * alpha) / (alpha * alpha);
                alpha = oracle_cubic(alpha_prev, alpha, f_prev, grad_dot_d, f_new, grad_alpha);
                if (alpha < 0.1 * alpha_prev || alpha > 0.9 * alpha_prev)
                    alpha = alpha_prev * 0.5;
            }
        } else {
            alpha = oracle_quadratic(alpha, 0.0, f_new, grad_dot_d, f_x);
            if (alpha < 0.1 * INITIAL_STEP_SIZE || alpha > 0.9 * INITIAL_STEP_SIZE)
                alpha = INITIAL_STEP_SIZE * 0.5;
        }
        alpha_prev = alpha;
        f_prev = f_new;
    }
    if (flavor == ORACLE_FLAVOR_PAR && alpha < 1e-4)
        return 0.5; /* par/line_search.cpp:223-226 */
    return alpha;
}

/* seq/line_search.cpp:125-189 (cubicInterpolate, C2=0.9) ;
 * par/line_search.cpp:298-369 (safeCubicInterpolate, C2=0.7) */
static double ls_wolfe(phi_t *p, int flavor)
{
    const double C2 = c2_of(flavor);
    double (*interp)(double, double, double, double, double, double) =
        flavor == ORACLE_FLAVOR_PAR ? oracle_safe_cubic : oracle_cubic;
    const double f_x = p->f0(p);
    const double grad_dot_d = p->gd;
    double alpha = INITIAL_STEP_SIZE;
    double alpha_lo = 0.0;
    double alpha_hi = INFINITY;
    double f_lo = f_x;
    double dphi_lo = grad_dot_d;
    for (int iter = 0; iter < 20; ++iter) {
        double f_new = p->f_at(p, alpha);
        if (f_new > f_x + C1 * alpha * grad_dot_d || (f_new >= f_lo && iter > 0)) {
            alpha_hi = alpha;
            alpha = interp(alpha_lo, alpha_hi, f_lo, dphi_lo, f_new,
                           (f_new - f_x - grad_dot_d * alpha) / (alpha * alpha));
            continue;
        }
        double dphi_new = p->df_at(p, alpha);
        if (fabs(dphi_new) <= -C2 * grad_dot_d)
            return alpha;
        if (dphi_new >= 0) {
            alpha_hi = alpha;
            alpha = interp(alpha_lo, alpha_hi, f_lo, dphi_lo, f_new, dphi_new);
        } else {
            alpha_lo = alpha;
            f_lo = f_new;
            dphi_lo = dphi_new;
            if (alpha_hi == INFINITY)
                alpha *= 2;
            else
                alpha = interp(alpha_lo, alpha_hi, f_lo, dphi_lo, f_new, dphi_new);
        }
        if (alpha < WOLFE_INTERP_MIN)
            return WOLFE_INTERP_MIN;
    }
    return alpha;
}

/* seq/line_search.cpp:33-55 (x0.5 / x1.1, C2=0.9, unbounded loop) ;
 * par/line_search.cpp:45-154 (bisection, local C2=0.9, TOL=1e-10, <=20 trials;
 * its unordered_map cache only memoises values that would be recomputed
 * identically, so it is not restated) */
static double ls_backtracking_wolfe(phi_t *p, int flavor)
{
    if (flavor == ORACLE_FLAVOR_SEQ) {
        const double C2 = 0.9;
        double alpha = INITIAL_STEP_SIZE;
        long guard = 0;
        while (1) {
            double dphi_new = p->df_at(p, alpha); /* grad evaluated first, :40 */
            if (p->f_at(p, alpha) > p->f0(p) + C1 * alpha * p->gd)
                alpha *= BACKTRACKING_ALPHA;
            else if (dphi_new < C2 * p->gd)
                alpha *= 1.1;
            else
                break;
            if (alpha < BACKTRACKING_TOL)
                break;
            if (++guard > 100000) /* the reference has no bound; the oracle refuses to hang */
                break;
        }
        return alpha;
    }
    const double lC1 = 1e-4, lC2 = 0.9, lTOL = 1e-10;
    double alpha = 1.0;
    int iter = 0;
    double f_current = p->f0(p);
    double gradient_dot_d = p->gd;
    double alpha_lo = 0.0;
    double alpha_hi = DBL_MAX;
    while (iter++ < 20) {
        double f_new = p->f_at(p, alpha);
        if (f_new <= f_current + lC1 * alpha * gradient_dot_d) {
            double gnd = p->df_at(p, alpha);
            if (gnd >= lC2 * gradient_dot_d)
                break;
            else
                alpha_lo = alpha;
        } else {
            alpha_hi = alpha;
        }
        if (alpha_hi < DBL_MAX)
            alpha = (alpha_lo + alpha_hi) / 2.0;
        else
            alpha = 2.0 * alpha_lo;
        if (alpha < lTOL)
            break;
    }
    return alpha;
}

static double run_ls(int ls, int flavor, phi_t *p)
{
    switch (ls) {
    case ORACLE_LS_BACKTRACKING: return ls_backtracking(p, flavor);
    case ORACLE_LS_INTERPOLATION: return ls_interpolation(p, flavor);
    case ORACLE_LS_WOLFE: return ls_wolfe(p, flavor);
    default: return ls_backtracking_wolfe(p, flavor);
    }
}


/* ------------------------------------------------------------------ */
/* the searches as INLINED in the CUDA solvers (ORACLE_FLAVOR_PAR_INLINED) */
/* ------------------------------------------------------------------ */
/* What the solver loop carries from one search to the next: the reference keeps evaluating trial
 * points into the host vector x_host and, at the top of the next iteration, takes "f(x_k)" as
 * f(x_host) (par/L-BFGS-Wolfe.cu:270, par/L-BFGS-Interpolation.cu:267,
 * par/L-BFGS-Backtracking_Wolfe.cu:266).  f_xhost is that value; f_initial is f(x0)
 * (par/L-BFGS-Wolfe.cu:172). */
typedef struct {
    double f_xhost, f_initial;
    int success;
} inl_t;

static double inl_f_at(phi_t *p, inl_t *q, double alpha)
{
    q->f_xhost = p->f_at(p, alpha); /* D2H of the trial point into x_host, then f(x_host) */
    return q->f_xhost;
}

/* par/L-BFGS-Wolfe.cu:260-349 */
static double ls_inl_wolfe(phi_t *p, inl_t *q)
{
    const double C2 = 0.7; /* par/constants.h:6 */
    const double grad_dot_d = p->gd;
    double alpha_current = INITIAL_STEP_SIZE;
    double alpha_lo = 0.0;
    double alpha_hi = INFINITY;
    double f_lo = q->f_initial; /* :267 -- f(x0) in every iteration */
    double dphi_lo = grad_dot_d;
    const double f_x = q->f_xhost; /* :270 */
    p->nf++;
    q->success = 0;
    for (int iter = 0; iter < 20; ++iter) {
        const double f_new = inl_f_at(p, q, alpha_current);
        if (f_new > f_x + C1 * alpha_current * grad_dot_d || (f_new >= f_lo && iter > 0)) {
            alpha_hi = alpha_current;
            alpha_current = oracle_safe_cubic(alpha_lo, alpha_hi, f_lo, dphi_lo, f_new,
                                              (f_new - f_x - grad_dot_d * alpha_current) /
                                                  (alpha_current * alpha_current));
            continue;
        }
        const double dphi_new = p->df_at(p, alpha_current);
        if (fabs(dphi_new) <= -C2 * grad_dot_d) {
            q->success = 1;
            break;
        }
        if (dphi_new >= 0) {
            alpha_hi = alpha_current;
            alpha_current = oracle_safe_cubic(alpha_lo, alpha_hi, f_lo, dphi_lo, f_new, dphi_new);
        } else {
            alpha_lo = alpha_current;
            f_lo = f_new;
            dphi_lo = dphi_new;
            if (alpha_hi == INFINITY)
                alpha_current *= 2;
            else
                alpha_current = oracle_safe_cubic(alpha_lo, alpha_hi, f_lo, dphi_lo, f_new, dphi_new);
        }
        if (alpha_current < WOLFE_INTERP_MIN) {
            alpha_current = WOLFE_INTERP_MIN; /* x_temp is updated, x_host is not (:339-346) */
            break;
        }
    }
    return alpha_current;
}

/* par/L-BFGS-Interpolation.cu:259-342 */
static double ls_inl_interpolation(phi_t *p, inl_t *q)
{
    const double grad_dot_d = p->gd;
    double alpha_current = INITIAL_STEP_SIZE;
    double alpha_prev = 0.0;
    double f_prev = q->f_initial; /* :265 */
    const double f_x = q->f_xhost; /* :267 */
    p->nf++;
    q->success = 0;
    for (int iter = 0; iter < 20; ++iter) {
        const double f_new = inl_f_at(p, q, alpha_current);
        if (f_new <= f_x + C1 * alpha_current * grad_dot_d) {
            q->success = 1;
            break;
        }
        if (alpha_current < WOLFE_INTERP_MIN) {
            alpha_current = WOLFE_INTERP_MIN;
            break;
        }
        if (alpha_prev > 0) {
            const double delta_alpha = alpha_current - alpha_prev;
            if (fabs(delta_alpha) < 1e-10) {
                alpha_current *= 0.5;
            } else {
                const double grad_alpha = (f_new - f_x - grad_dot_d * alpha_current) / (alpha_current * alpha_current);
                double next_alpha = oracle_cubic(alpha_prev, alpha_current, f_prev, grad_dot_d, f_new, grad_alpha);
                if (next_alpha < 0.1 * alpha_prev || next_alpha > 0.9 * alpha_prev)
                    next_alpha = alpha_prev * 0.5;
                alpha_current = next_alpha;
            }
        } else {
            double next_alpha = oracle_quadratic(alpha_current, 0.0, f_new, grad_dot_d, f_x);
            if (next_alpha < 0.1 * INITIAL_STEP_SIZE || next_alpha > 0.9 * INITIAL_STEP_SIZE)
                next_alpha = INITIAL_STEP_SIZE * 0.5;
            alpha_current = next_alpha;
        }
        alpha_prev = alpha_current;
        f_prev = f_new;
    }
    if (alpha_current < 1e-4) /* :338-341, after the loop: applies to every exit */
        alpha_current = 0.5;
    return alpha_current;
}

/* par/L-BFGS-Backtracking.cu:292-341 (local constants :153-156; f_current re-read from d_x :308-312) */
static double ls_inl_backtracking(phi_t *p, inl_t *q)
{
    const double lTOL = 1e-10;
    double step_size = INITIAL_STEP_SIZE;
    const double f_current = p->f0(p);
    q->success = 0;
    for (;;) {
        const double f_trial = inl_f_at(p, q, step_size);
        if (f_trial <= f_current + C1 * step_size * p->gd) {
            q->success = 1;
            break;
        }
        step_size *= BACKTRACKING_ALPHA;
        if (step_size < lTOL) {
            step_size = 0.5;
            break;
        }
    }
    return step_size;
}

/* par/L-BFGS-Backtracking_Wolfe.cu:262-397 (the unordered_map caches only memoise) */
static double ls_inl_btwolfe(phi_t *p, inl_t *q)
{
    const double lC1 = 1e-4, lC2 = 0.9, lTOL = 1e-10;
    const double f_x = q->f_xhost; /* :266 */
    const double grad_dot_d = p->gd;
    double alpha_current = 1.0, alpha_lo = 0.0, alpha_hi = DBL_MAX;
    p->nf++;
    q->success = 0;
    for (int iter = 0; iter < 20; ++iter) {
        const double f_new = inl_f_at(p, q, alpha_current);
        if (f_new <= f_x + lC1 * alpha_current * grad_dot_d) {
            const double gnd = p->df_at(p, alpha_current);
            if (gnd >= lC2 * grad_dot_d) {
                q->success = 1;
                break;
            }
            alpha_lo = alpha_current;
        } else {
            alpha_hi = alpha_current;
        }
        if (alpha_hi < DBL_MAX)
            alpha_current = (alpha_lo + alpha_hi) / 2.0;
        else
            alpha_current = 2.0 * alpha_lo;
        if (alpha_current < lTOL) {
            alpha_current = lTOL;
            (void)inl_f_at(p, q, alpha_current); /* :371-396: evaluated at TOL, x_host follows */
            break;
        }
    }
    return alpha_current;
}

static double run_ls_inlined(int ls, phi_t *p, inl_t *q)
{
    switch (ls) {
    case ORACLE_LS_BACKTRACKING: return ls_inl_backtracking(p, q);
    case ORACLE_LS_INTERPOLATION: return ls_inl_interpolation(p, q);
    case ORACLE_LS_WOLFE: return ls_inl_wolfe(p, q);
    default: return ls_inl_btwolfe(p, q);
    }
}

/* ---- 1-D polynomial phi, for the host state-machine tests ---- */
static double poly_f(phi_t *p, double a)
{
    const double *c = (const double *)p->ctx;
    p->nf++; p->ntrial++;
    return c[0] + a * (c[1] + a * (c[2] + a * (c[3] + a * c[4])));
}
static double poly_df(phi_t *p, double a)
{
    const double *c = (const double *)p->ctx;
    p->ng++;
    return c[1] + a * (2 * c[2] + a * (3 * c[3] + a * 4 * c[4]));
}
static double poly_f0(phi_t *p)
{
    const double *c = (const double *)p->ctx;
    p->nf++;
    return c[0];
}

double oracle_ls_poly(int line_search, int flavor, const double coef[5], int *nf, int *ng)
{
    phi_t p;
    memset(&p, 0, sizeof p);
    p.f_at = poly_f;
    p.df_at = poly_df;
    p.f0 = poly_f0;
    p.gd = coef[1];
    p.ctx = (void *)coef;
    double a = run_ls(line_search, flavor, &p);
    if (nf) *nf = (int)p.nf;
    if (ng) *ng = (int)p.ng;
    return a;
}

/* ---- vector phi ---- */
typedef struct {
    int objective;
    size_t n;
    const double *x, *d;
    double *xt, *gt;
} vec_ctx_t;

static void make_trial(vec_ctx_t *c, double alpha)
{
    /* add(x, scalarProduct(alpha, d)) and x[i] + alpha*d[i] round identically without FMA */
    for (size_t i = 0; i < c->n; ++i)
        c->xt[i] = c->x[i] + alpha * c->d[i];
}
static double vec_f(phi_t *p, double a)
{
    vec_ctx_t *c = (vec_ctx_t *)p->ctx;
    make_trial(c, a);
    p->nf++; p->ntrial++;
    return oracle_f(c->objective, c->xt, c->n);
}
static double vec_df(phi_t *p, double a)
{
    vec_ctx_t *c = (vec_ctx_t *)p->ctx;
    make_trial(c, a);
    p->ng++;
    oracle_grad(c->objective, c->xt, c->gt, c->n);
    return oracle_dot(c->gt, c->d, c->n);
}
static double vec_f0(phi_t *p)
{
    vec_ctx_t *c = (vec_ctx_t *)p->ctx;
    p->nf++;
    return oracle_f(c->objective, c->x, c->n);
}

/* inlined searches over the polynomial: f_xhost is what the solver loop would hand over as
 * "f(x_k)", f_initial as f(x0).  *f_xhost_out = f at the last point the search evaluated. */
double oracle_ls_poly_inlined(int line_search, const double coef[5], double f_xhost, double f_initial,
                              int *nf, int *ng, int *success, double *f_xhost_out)
{
    phi_t p;
    memset(&p, 0, sizeof p);
    p.f_at = poly_f; p.df_at = poly_df; p.f0 = poly_f0;
    p.gd = coef[1];
    p.ctx = (void *)coef;
    inl_t q = { f_xhost, f_initial, 0 };
    const double a = run_ls_inlined(line_search, &p, &q);
    if (nf) *nf = (int)p.ntrial;
    if (ng) *ng = (int)p.ng;
    if (success) *success = q.success;
    if (f_xhost_out) *f_xhost_out = q.f_xhost;
    return a;
}

/* ------------------------------------------------------------------ */
/* two-loop recursion: seq/lbfgs.cpp:93-143                            */
/* ------------------------------------------------------------------ */
static int two_loop_ptrs(const double *g, double *const *S, double *const *Y, int h, size_t n,
                         double *d, double *q, double *alpha)
{
    memcpy(q, g, n * sizeof(double));
    for (int i = h - 1; i >= 0; --i) {
        double rho = 1.0 / oracle_dot(Y[i], S[i], n); /* :102 */
        if (!isfinite(rho))
            return 1; /* :103-108 */
        alpha[i] = rho * oracle_dot(S[i], q, n); /* :109 */
        for (size_t j = 0; j < n; ++j)
            q[j] -= alpha[i] * Y[i][j]; /* :110-113 */
    }
    double gamma = oracle_dot(S[h - 1], Y[h - 1], n) / oracle_dot(Y[h - 1], Y[h - 1], n); /* :117 */
    if (gamma <= 0 || !isfinite(gamma))
        return 1; /* :119-124 */
    for (size_t j = 0; j < n; ++j)
        q[j] *= gamma; /* r = gamma*q, :126-130 */
    for (int i = 0; i < h; ++i) {
        double rho = 1.0 / oracle_dot(Y[i], S[i], n); /* :135 */
        double beta = rho * oracle_dot(Y[i], q, n);   /* :136 */
        double c = alpha[i] - beta;
        for (size_t j = 0; j < n; ++j)
            q[j] += S[i][j] * c; /* :137-140 */
    }
    for (size_t j = 0; j < n; ++j)
        d[j] = -q[j]; /* :143 */
    return 0;
}

int oracle_two_loop(const double *g, const double *S, const double *Y, int h, size_t n, double *d)
{
    if (h <= 0) {
        for (size_t j = 0; j < n; ++j) d[j] = -g[j];
        return 0;
    }
    double **s = (double **)malloc(sizeof(double *) * h);
    double **y = (double **)malloc(sizeof(double *) * h);
    for (int i = 0; i < h; ++i) {
        s[i] = (double *)(S + (size_t)i * n);
        y[i] = (double *)(Y + (size_t)i * n);
    }
    double *q = (double *)malloc(sizeof(double) * n);
    double *a = (double *)malloc(sizeof(double) * h);
    int rc = two_loop_ptrs(g, s, y, h, n, d, q, a);
    if (rc)
        for (size_t j = 0; j < n; ++j) d[j] = -g[j];
    free(q); free(a); free(s); free(y);
    return rc;
}

/* ------------------------------------------------------------------ */
/* solver: seq/lbfgs.cpp:17-203                                        */
/* ------------------------------------------------------------------ */
int oracle_lbfgs(const oracle_params_t *p, size_t n, const double *x0, double *x_out,
                 double *trace, size_t trace_rows, oracle_result_t *res)
{
    const int m = p->m;
    long nf = 0, ng = 0;
    int status = ORACLE_STATUS_MAX_ITER;
    double *x = (double *)malloc(n * sizeof(double));
    double *g = (double *)malloc(n * sizeof(double));
    double *d = (double *)malloc(n * sizeof(double));
    double *q = (double *)malloc(n * sizeof(double));
    double *xn = (double *)malloc(n * sizeof(double));
    double *gn = (double *)malloc(n * sizeof(double));
    double *xt = (double *)malloc(n * sizeof(double));
    double *gt = (double *)malloc(n * sizeof(double));
    double *alpha = (double *)malloc(sizeof(double) * (m > 0 ? m : 1));
    /* deque of accepted pairs, oldest first (:32, :182-190) */
    double **S = (double **)calloc(m > 0 ? m : 1, sizeof(double *));
    double **Y = (double **)calloc(m > 0 ? m : 1, sizeof(double *));
    int h = 0;

    memcpy(x, x0, n * sizeof(double));
    double f_current = oracle_f(p->objective, x, n); nf++; /* :29 */
    oracle_grad(p->objective, x, g, n); ng++;              /* :30 */

    int k;
    for (k = 0; k < p->max_iterations; ++k) {
        if (oracle_norm(g, n) < p->tolerance) { /* :80-84 */
            status = ORACLE_STATUS_CONVERGED;
            break;
        }
        int steepest = 1;
        if (!(k == 0 || h == 0)) /* :87 */
            steepest = two_loop_ptrs(g, S, Y, h, n, d, q, alpha);
        if (steepest)
            for (size_t j = 0; j < n; ++j) d[j] = -g[j];

        double grad_dot_d = oracle_dot(g, d, n); /* :146 */
        if (grad_dot_d >= 0) {                   /* :147-153 */
            for (size_t j = 0; j < n; ++j) d[j] = -g[j];
            grad_dot_d = oracle_dot(g, d, n);
        }

        vec_ctx_t c = { p->objective, n, x, d, xt, gt };
        phi_t phi;
        memset(&phi, 0, sizeof phi);
        phi.f_at = vec_f; phi.df_at = vec_df; phi.f0 = vec_f0;
        phi.gd = grad_dot_d; phi.ctx = &c;
        double a = run_ls(p->line_search, p->flavor, &phi); /* :156 */
        nf += phi.nf; ng += phi.ng;

        for (size_t j = 0; j < n; ++j)
            xn[j] = x[j] + a * d[j];             /* :159 add(x, scalarProduct(alpha,d)) */
        f_current = oracle_f(p->objective, xn, n); nf++; /* :160-161 */

        if (a < 1e-10) { /* :164-168: returns the OLD x */
            status = ORACLE_STATUS_LS_FAILED;
            break;
        }
        oracle_grad(p->objective, xn, gn, n); ng++; /* :171 */

        /* :174-195 */
        double *s_k = (double *)malloc(n * sizeof(double));
        double *y_k = (double *)malloc(n * sizeof(double));
        for (size_t j = 0; j < n; ++j) {
            s_k[j] = xn[j] - x[j];
            y_k[j] = gn[j] - g[j];
        }
        double sy = oracle_dot(s_k, y_k, n);
        if (sy > 0 && m > 0) {
            if (h >= m) {
                free(S[0]); free(Y[0]);
                memmove(S, S + 1, sizeof(double *) * (m - 1));
                memmove(Y, Y + 1, sizeof(double *) * (m - 1));
                h = m - 1;
            }
            S[h] = s_k; Y[h] = y_k; h++;
        } else {
            free(s_k); free(y_k);
        }
        memcpy(x, xn, n * sizeof(double)); /* :197-198 */
        memcpy(g, gn, n * sizeof(double));

        if (trace && (size_t)k < trace_rows) {
            double *row = trace + (size_t)k * ORACLE_TRACE_COLS;
            row[ORACLE_TR_K] = (double)k;
            row[ORACLE_TR_F] = f_current;
            row[ORACLE_TR_GNORM] = oracle_norm(g, n);
            row[ORACLE_TR_ALPHA] = a;
            row[ORACLE_TR_TRIALS] = (double)phi.ntrial;
            row[ORACLE_TR_HIST] = (double)h;
            row[ORACLE_TR_X0] = x[0];
            row[ORACLE_TR_XMID] = x[n / 2];
        }
    }

    if (x_out) memcpy(x_out, x, n * sizeof(double));
    if (res) {
        res->status = status;
        res->iterations = k;
        res->f_evals = nf;
        res->g_evals = ng;
        res->f = oracle_f(p->objective, x, n);
        oracle_grad(p->objective, x, gn, n);
        res->gnorm = oracle_norm(gn, n);
    }
    for (int i = 0; i < h; ++i) { free(S[i]); free(Y[i]); }
    free(S); free(Y); free(alpha);
    free(x); free(g); free(d); free(q); free(xn); free(gn); free(xt); free(gt);
    return status;
}

/* ------------------------------------------------------------------ */
/* CUDA-tree outer loop: par/L-BFGS.cu:195-357 (see the header for the pin)  */
/* ------------------------------------------------------------------ */
int oracle_lbfgs_cuda_profile(const oracle_params_t *p, size_t n, const double *x0, double *x_out,
                              double *trace, size_t trace_rows, oracle_result_t *res)
{
    const int m = p->m;
    long nf = 0, ng = 0;
    int status = ORACLE_STATUS_MAX_ITER;
    double *x = (double *)malloc(n * sizeof(double)), *g = (double *)malloc(n * sizeof(double));
    double *d = (double *)malloc(n * sizeof(double)), *q = (double *)malloc(n * sizeof(double));
    double *xn = (double *)malloc(n * sizeof(double)), *gn = (double *)malloc(n * sizeof(double));
    double *xt = (double *)malloc(n * sizeof(double)), *gt = (double *)malloc(n * sizeof(double));
    double *S = (double *)calloc((size_t)m * n, sizeof(double)), *Y = (double *)calloc((size_t)m * n, sizeof(double));
    double *alpha = (double *)calloc(m, sizeof(double)), *rho = (double *)calloc(m, sizeof(double));
    int *skip = (int *)calloc(m, sizeof(int));
    inl_t inl = { 0.0, 0.0, 0 };
    memcpy(x, x0, n * sizeof(double));
    double f_cur = oracle_f(p->objective, x, n);
    /* par/L-BFGS.cu and par/L-BFGS-Backtracking.cu never evaluate f(x0) outside a search */
    if (p->flavor == ORACLE_FLAVOR_PAR_INLINED && p->line_search != ORACLE_LS_BACKTRACKING) nf++;
    oracle_grad(p->objective, x, g, n); ng++; /* :199 */
    double *g0 = (double *)malloc(n * sizeof(double));
    memcpy(g0, g, n * sizeof(double));
    int k;
    for (k = 0; k < p->max_iterations; ++k) {
        if (k == 0) {
            for (size_t j = 0; j < n; ++j) d[j] = -g[j]; /* :208 */
        } else {
            memcpy(q, g, n * sizeof(double)); /* :212 */
            const int lo = k - m > 0 ? k - m : 0;
            for (int i = k - 1; i >= lo; --i) { /* :216-236 */
                const double *s = S + (size_t)(i % m) * n, *y = Y + (size_t)(i % m) * n;
                const double si_yi = oracle_dot(s, y, n);
                skip[i % m] = si_yi <= 1e-10; /* :222-223 */
                if (skip[i % m]) continue;
                rho[i % m] = 1.0 / si_yi;
                alpha[i % m] = rho[i % m] * oracle_dot(s, q, n);
                for (size_t j = 0; j < n; ++j) q[j] -= alpha[i % m] * y[j];
            }
            const int last = (k - 1) % m; /* :239-255 */
            const double ys = oracle_dot(S + (size_t)last * n, Y + (size_t)last * n, n);
            const double yy = oracle_dot(Y + (size_t)last * n, Y + (size_t)last * n, n);
            const double gamma = (yy > 0 && ys > 1e-10) ? ys / yy : 1.0;
            for (size_t j = 0; j < n; ++j) q[j] *= gamma;
            for (int i = lo; i < k; ++i) { /* :263-274 */
                if (skip[i % m]) continue; /* (the reference would reuse stale rho/alpha here) */
                const double *s = S + (size_t)(i % m) * n, *y = Y + (size_t)(i % m) * n;
                const double beta = rho[i % m] * oracle_dot(y, q, n);
                const double c = alpha[i % m] - beta;
                for (size_t j = 0; j < n; ++j) q[j] += s[j] * c;
            }
            for (size_t j = 0; j < n; ++j) d[j] = -q[j]; /* :276 */
        }
        vec_ctx_t c = { p->objective, n, x, d, xt, gt };
        phi_t phi;
        memset(&phi, 0, sizeof phi);
        phi.f_at = vec_f; phi.df_at = vec_df; phi.f0 = vec_f0;
        /* par/L-BFGS.cu:293 passes `gradient`, assigned once at k == 0 (:199) */
        phi.gd = oracle_dot(p->flavor == ORACLE_FLAVOR_PAR_STALE_GRADIENT ? g0 : g, d, n); phi.ctx = &c;
        double a;
        if (p->flavor == ORACLE_FLAVOR_PAR_INLINED) {
            if (k == 0) { inl.f_xhost = f_cur; inl.f_initial = f_cur; } /* par/L-BFGS-Wolfe.cu:165-172 */
            a = run_ls_inlined(p->line_search, &phi, &inl);
            nf += phi.nf; ng += phi.ng;
            /* par/L-BFGS-Wolfe.cu:353, par/L-BFGS-Interpolation.cu:345, par/L-BFGS-Backtracking_Wolfe.cu:401;
             * the inlined backtracking has no failure exit */
            if (p->line_search != ORACLE_LS_BACKTRACKING && !inl.success && a < 1e-10) {
                status = ORACLE_STATUS_LS_FAILED;
                break;
            }
        } else {
            a = run_ls(p->line_search, ORACLE_FLAVOR_PAR, &phi); /* :293 (with the CURRENT gradient unless ..._STALE_GRADIENT) */
            nf += phi.nf; ng += phi.ng;
            if (a < 1e-10) { status = ORACLE_STATUS_LS_FAILED; break; } /* :295-305 */
        }
        for (size_t j = 0; j < n; ++j) xn[j] = x[j] + a * d[j]; /* :309 (no FMA restated) */
        oracle_grad(p->objective, xn, gn, n);                    /* :323 */
        /* the inlined Wolfe searches keep the gradient of the accepted trial instead of evaluating it again
         * (par/L-BFGS-Wolfe.cu:385-395, par/L-BFGS-Backtracking_Wolfe.cu:432): same values, one call less */
        if (!(p->flavor == ORACLE_FLAVOR_PAR_INLINED && inl.success &&
              (p->line_search == ORACLE_LS_WOLFE || p->line_search == ORACLE_LS_BACKTRACKING_WOLFE)))
            ng++;
        double *s = S + (size_t)(k % m) * n, *y = Y + (size_t)(k % m) * n; /* :332-333 */
        for (size_t j = 0; j < n; ++j) { s[j] = xn[j] - x[j]; y[j] = gn[j] - g[j]; }
        memcpy(x, xn, n * sizeof(double));
        memcpy(g, gn, n * sizeof(double));
        f_cur = oracle_f(p->objective, x, n); nf++; /* :351 (printed only) */
        const double norm_g = sqrt(oracle_dot(g, g, n)); /* :346-347 */
        if (trace && (size_t)k < trace_rows) {
            double *row = trace + (size_t)k * ORACLE_TRACE_COLS;
            row[ORACLE_TR_K] = (double)k; row[ORACLE_TR_F] = f_cur; row[ORACLE_TR_GNORM] = norm_g;
            row[ORACLE_TR_ALPHA] = a; row[ORACLE_TR_TRIALS] = (double)phi.ntrial;
            row[ORACLE_TR_HIST] = (double)(k + 1 < m ? k + 1 : m);
            row[ORACLE_TR_X0] = x[0]; row[ORACLE_TR_XMID] = x[n / 2];
        }
        if (norm_g <= p->tolerance) { status = ORACLE_STATUS_CONVERGED; ++k; break; } /* :353-357 */
    }
    if (x_out) memcpy(x_out, x, n * sizeof(double));
    if (res) {
        res->status = status; res->iterations = k; res->f_evals = nf; res->g_evals = ng;
        res->f = oracle_f(p->objective, x, n);
        oracle_grad(p->objective, x, gn, n);
        res->gnorm = oracle_norm(gn, n);
    }
    free(x); free(g); free(d); free(q); free(xn); free(gn); free(xt); free(gt);
    free(S); free(Y); free(alpha); free(rho); free(skip); free(g0);
    (void)f_cur;
    return status;
}

/* ------------------------------------------------------------------ */
/* x0 generator: libstdc++ mt19937 + uniform_real_distribution<double>  */
/* (seq/main.cpp:34-43, par/L-BFGS-Wolfe.cu:458-465)                    */
/* ------------------------------------------------------------------ */
void oracle_x0(unsigned seed, double lo, double hi, size_t n, double *out)
{
    /* MT19937 (Matsumoto & Nishimura 1998), init_genrand seeding as std::mt19937(seed) */
    uint32_t mt[624];
    int idx = 624;
    mt[0] = seed;
    for (int i = 1; i < 624; ++i)
        mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    for (size_t t = 0; t < n; ++t) {
        double draws[2];
        for (int r = 0; r < 2; ++r) {
            if (idx >= 624) {
                for (int i = 0; i < 624; ++i) {
                    uint32_t yv = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
                    mt[i] = mt[(i + 397) % 624] ^ (yv >> 1) ^ ((yv & 1u) ? 0x9908b0dfu : 0u);
                }
                idx = 0;
            }
            uint32_t yv = mt[idx++];
            yv ^= yv >> 11;
            yv ^= (yv << 7) & 0x9d2c5680u;
            yv ^= (yv << 15) & 0xefc60000u;
            yv ^= yv >> 18;
            draws[r] = (double)yv;
        }
        /* std::generate_canonical<double,53>: two 32-bit draws, low word first */
        double sum = draws[0] + draws[1] * 4294967296.0;
        double c = sum / 18446744073709551616.0;
        if (c >= 1.0)
            c = nextafter(1.0, 0.0);
        out[t] = c * (hi - lo) + lo;
    }
}
