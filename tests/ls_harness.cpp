// tests/ls_harness.cpp -- host build of the product's line-search state machines
// (cuda-lbfgs_b200/csrc/ls_logic.h), driven over a 1-D polynomial phi so the not-gpu tests
// can compare them, decision by decision, with the oracle's restatement of the reference
// loops (oracle_ls_poly).  Test-only: the product runs ls_step() in the GPU scalar kernel.
#include "../cuda-lbfgs_b200/csrc/ls_logic.h"

using namespace lb;

static double poly(const double *c, double a) { return c[0] + a * (c[1] + a * (c[2] + a * (c[3] + a * c[4]))); }
static double dpoly(const double *c, double a) { return c[1] + a * (2 * c[2] + a * (3 * c[3] + a * 4 * c[4])); }

extern "C" double harness_ls_poly(int kind, int flavor, const double *coef, int *trials)
{
    LsParams p;
    p.kind = kind;
    p.flavor = flavor;
    p.max_trials = 20;
    p.c1 = 1e-4;
    p.c2 = flavor != FLAVOR_SEQ ? 0.7 : 0.9;
    p.step0 = 1.0;
    p.shrink = 0.5;
    p.bt_tol = 1e-8;
    p.wolfe_min = 1e-10;
    LsState s = {};
    int go = ls_begin(p, s, coef[0], coef[1]);
    while (go) go = ls_step(p, s, poly(coef, s.alpha), dpoly(coef, s.alpha));
    if (trials) *trials = s.trials;
    return s.alpha;
}

// FLAVOR_PAR_INLINED: the previous search left x_host at a point with f = f_xhost (stale), f(x0) = f_initial.
// *f_xhost_out = what the NEXT search would take as f(x_k): f_last if this search returned a step it did not
// evaluate, f at the returned step otherwise.
extern "C" double harness_ls_poly_inlined(int kind, const double *coef, double f_xhost, double f_initial, int *trials,
                                          int *success, double *f_xhost_out)
{
    LsParams p;
    p.kind = kind;
    p.flavor = FLAVOR_PAR_INLINED;
    p.max_trials = 20;
    p.c1 = 1e-4;
    p.c2 = 0.7;
    p.step0 = 1.0;
    p.shrink = 0.5;
    p.bt_tol = 1e-10; // par/L-BFGS-Backtracking.cu:155
    p.wolfe_min = 1e-10;
    LsState s = {};
    s.stale = 1;
    s.f_last = f_xhost;
    int go = ls_begin(p, s, coef[0], coef[1], f_initial);
    while (go) go = ls_step(p, s, poly(coef, s.alpha), dpoly(coef, s.alpha));
    if (trials) *trials = s.trials;
    if (success) *success = s.success;
    if (f_xhost_out) *f_xhost_out = s.stale ? s.f_last : poly(coef, s.alpha);
    return s.alpha;
}

extern "C" double harness_cubic(double a0, double a1, double p0, double d0, double p1, double d1)
{
    return cubic_interpolate(a0, a1, p0, d0, p1, d1);
}
extern "C" double harness_safe_cubic(double a0, double a1, double p0, double d0, double p1, double d1)
{
    return safe_cubic_interpolate(a0, a1, p0, d0, p1, d1);
}
extern "C" double harness_quadratic(double a0, double p0, double d0, double p1)
{
    return quadratic_interpolate(a0, p0, d0, p1);
}
