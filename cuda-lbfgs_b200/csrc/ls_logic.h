// ls_logic.h -- the line-search decision logic as resumable scalar state machines.
//
// The reference's line searches are loops around host callbacks f() / grad()
// (seq/line_search.cpp, par/line_search.cpp).  Here a trial evaluation is a fused GPU
// kernel, so each search is restated as:  ls_begin() -> first alpha ;  ls_step(f_new,
// dphi_new) -> either "evaluate this next alpha" or "finished with this alpha".
// One thread of the scalar kernel runs it on the device; the same code compiles for the
// host (plain C++, no CUDA needed) so the not-gpu tests can drive it against the oracle.
// Compiled with FMA contraction off (-fmad=false / -ffp-contract=off): the expressions
// keep the reference's evaluation order and round identically to the x86 build.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define LB_HD __host__ __device__ inline
#else
#define LB_HD inline
#endif

namespace lb {

enum { LS_BACKTRACKING = 0, LS_INTERPOLATION = 1, LS_WOLFE = 2, LS_BACKTRACKING_WOLFE = 3 };
// FLAVOR_SEQ: seq/line_search.cpp + seq/config.h.  FLAVOR_PAR: par/line_search.cpp + par/constants.h (what
// par/L-BFGS.cu calls).  FLAVOR_PAR_INLINED: the searches as they are inlined in the other CUDA solvers
// (par/L-BFGS-Wolfe.cu:260-349, par/L-BFGS-Interpolation.cu:259-342, par/L-BFGS-Backtracking.cu:292-341,
// par/L-BFGS-Backtracking_Wolfe.cu:262-397), which differ from par/line_search.cpp in a few places --
// see ls_begin(..., f0) and the FLAVOR_PAR_INLINED branches of ls_step().
enum { FLAVOR_SEQ = 0, FLAVOR_PAR = 1, FLAVOR_PAR_INLINED = 2 };

struct LsParams {
    int kind;   // LS_*
    int flavor; // FLAVOR_*
    int max_trials;
    double c1, c2, step0, shrink, bt_tol, wolfe_min;
};

struct LsState {
    double alpha;  // trial step to evaluate next / final step when finished
    double f_x;    // f(x)
    double gd;     // grad f(x) . d
    int trials;    // trial points evaluated so far
    // interpolation
    double alpha_prev, f_prev;
    // wolfe / bisection wolfe
    double lo, hi, f_lo, dphi_lo;
    // bookkeeping of the last evaluated trial (FLAVOR_PAR_INLINED: the reference's x_host)
    double alpha_last, f_last;
    int success;   // the search ended by accepting an evaluated trial
    int stale;     // it returned a step that differs from the last evaluated trial
};

// seq/line_search.cpp:8-12 == par/line_search.cpp:10-15
LB_HD double cubic_interpolate(double alpha0, double alpha1, double phi0, double dphi0,
                               double phi1, double dphi1)
{
    double d1 = dphi0 + dphi1 - 3 * (phi1 - phi0) / (alpha1 - alpha0);
    double d2 = copysign(sqrt(d1 * d1 - dphi0 * dphi1), alpha1 - alpha0);
    return alpha0 + (alpha1 - alpha0) * (dphi0 + d2 - d1) / (dphi0 - dphi1 + 2 * d2);
}

// seq/line_search.cpp:14-16 == par/line_search.cpp:17-20
LB_HD double quadratic_interpolate(double alpha0, double phi0, double dphi0, double phi1)
{
    return alpha0 - 0.5 * dphi0 * alpha0 * alpha0 / (phi1 - phi0 - dphi0 * alpha0);
}

// par/line_search.cpp:231-296 (its try/catch blocks guard arithmetic that cannot throw)
LB_HD double safe_cubic_interpolate(double alpha0, double alpha1, double phi0, double dphi0,
                                    double phi1, double dphi1)
{
    if (alpha0 > alpha1) {
        double t;
        t = alpha0; alpha0 = alpha1; alpha1 = t;
        t = phi0; phi0 = phi1; phi1 = t;
        t = dphi0; dphi0 = dphi1; dphi1 = t;
    }
    const double mid = 0.5 * (alpha0 + alpha1);
    double d1 = dphi0 + dphi1 - 3 * (phi1 - phi0) / (alpha1 - alpha0);
    if (isnan(d1) || isinf(d1)) return mid;
    double discriminant = d1 * d1 - dphi0 * dphi1;
    if (discriminant < 0) return mid;
    double d2 = copysign(sqrt(discriminant), alpha1 - alpha0);
    double denominator = dphi0 - dphi1 + 2 * d2;
    if (fabs(denominator) < 1e-10) return mid;
    double result = alpha0 + (alpha1 - alpha0) * (dphi0 + d2 - d1) / denominator;
    if (isnan(result) || isinf(result)) return mid;
    const double hi = alpha1 - 0.1 * (alpha1 - alpha0);
    const double lo = alpha0 + 0.1 * (alpha1 - alpha0);
    const double mn = (result < hi) ? result : hi; // std::min(hi, result)
    return (lo < mn) ? mn : lo;                    // std::max(lo, mn)
}

LB_HD double ls_interp(const LsParams &p, double a0, double a1, double p0, double dp0, double p1,
                       double dp1)
{
    return p.flavor != FLAVOR_SEQ ? safe_cubic_interpolate(a0, a1, p0, dp0, p1, dp1)
                                  : cubic_interpolate(a0, a1, p0, dp0, p1, dp1);
}

// "return 0.5 when the step got tiny": par/line_search.cpp:38-41, :223-226
LB_HD double ls_par_floor(const LsParams &p, double alpha)
{
    if (p.flavor == FLAVOR_PAR && alpha < 1e-4) return 0.5;
    return alpha;
}

// Start a search from f(x) and grad.d.  Always asks for a first trial (returns 1).
LB_HD int ls_begin(const LsParams &p, LsState &s, double f_x, double gd)
{
    s.alpha = p.step0;
    s.f_x = f_x;
    s.gd = gd;
    s.trials = 0;
    s.alpha_prev = 0.0;
    s.f_prev = f_x;
    s.lo = 0.0;
    s.hi = (p.kind == LS_BACKTRACKING_WOLFE) ? 1.7976931348623157e308 : INFINITY;
    s.f_lo = f_x;
    s.dphi_lo = gd;
    s.success = 0;
    return 1;
}

// Start a search inside the solver loop: f_cur = f(x_k), f0 = f(x_0).  For FLAVOR_SEQ / FLAVOR_PAR this is
// ls_begin(p, s, f_cur, gd).  The inlined CUDA searches (FLAVOR_PAR_INLINED) instead
//  * take "f(x_k)" from f(x_host), where x_host still holds the LAST TRIAL POINT the previous search evaluated
//    (par/L-BFGS-Wolfe.cu:270 with :282-288, par/L-BFGS-Interpolation.cu:267 with :279-285,
//    par/L-BFGS-Backtracking_Wolfe.cu:266 with :297-303): equal to f(x_k) only when that search returned
//    the step it evaluated last; the inlined backtracking re-reads d_x and has no such dependence
//    (par/L-BFGS-Backtracking.cu:308-312);
//  * seed the Wolfe bracket's f_lo with f(x_0) in EVERY iteration (par/L-BFGS-Wolfe.cu:267, initial_f :172).
// s must still hold the state the previous search left behind (all zero before the first one).
LB_HD int ls_begin(const LsParams &p, LsState &s, double f_cur, double gd, double f0)
{
    double f_x = f_cur;
    if (p.flavor == FLAVOR_PAR_INLINED && p.kind != LS_BACKTRACKING && s.stale) f_x = s.f_last;
    ls_begin(p, s, f_x, gd);
    if (p.flavor == FLAVOR_PAR_INLINED && p.kind == LS_WOLFE) s.f_lo = f0;
    return 1;
}

LB_HD int ls_step_core(const LsParams &p, LsState &s, double f_new, double dphi_new)
{
    const double alpha = s.alpha;
    const int iter = s.trials; // index of the trial just evaluated
    s.trials = iter + 1;

    switch (p.kind) {
    case LS_BACKTRACKING: {
        if (p.flavor == FLAVOR_PAR_INLINED) {
            // par/L-BFGS-Backtracking.cu:314-341: the textbook Armijo test, and 0.5 once the step is tiny
            if (f_new <= s.f_x + p.c1 * alpha * s.gd) { s.success = 1; return 0; }
            const double a = alpha * p.shrink;
            s.alpha = a;
            if (a < p.bt_tol) { s.alpha = 0.5; return 0; }
            return 1;
        }
        // seq/line_search.cpp:23-27.  The reference's test is
        //   f(x) - f(x+alpha d) < C1*alpha*(g.d)   (continue shrinking while true)
        if (s.f_x - f_new < p.c1 * alpha * s.gd) {
            double a = alpha * p.shrink;
            s.alpha = a;
            if (a < p.bt_tol) { // break: the returned step was never evaluated
                s.alpha = ls_par_floor(p, a);
                return 0;
            }
            return 1;
        }
        s.alpha = ls_par_floor(p, alpha);
        return 0;
    }
    case LS_INTERPOLATION: {
        // seq/line_search.cpp:73-118
        // the inlined copy applies its "alpha < 1e-4 => 0.5" AFTER the loop, i.e. to every exit
        // (par/L-BFGS-Interpolation.cu:338-341); par/line_search.cpp only to the exhausted one (:223-226)
        const bool inl = p.flavor == FLAVOR_PAR_INLINED;
        if (f_new <= s.f_x + p.c1 * alpha * s.gd) { // :83-85 (no floor on this path)
            s.success = 1;
            if (inl && alpha < 1e-4) s.alpha = 0.5;
            return 0;
        }
        if (alpha < p.wolfe_min) { s.alpha = inl ? 0.5 : p.wolfe_min; return 0; } // :87-89
        double a;
        if (s.alpha_prev > 0) {
            double delta_alpha = alpha - s.alpha_prev;
            if (fabs(delta_alpha) < 1e-10) {
                a = alpha * 0.5;
            } else {
                double grad_alpha = (f_new - s.f_x - s.gd * alpha) / (alpha * alpha);
                a = cubic_interpolate(s.alpha_prev, alpha, s.f_prev, s.gd, f_new, grad_alpha);
                if (a < 0.1 * s.alpha_prev || a > 0.9 * s.alpha_prev) a = s.alpha_prev * 0.5;
            }
        } else {
            a = quadratic_interpolate(alpha, f_new, s.gd, s.f_x);
            if (a < 0.1 * p.step0 || a > 0.9 * p.step0) a = p.step0 * 0.5;
        }
        s.alpha_prev = a; // :116-117 (assigned AFTER the update, as in the reference)
        s.f_prev = f_new;
        s.alpha = a;
        if (s.trials < p.max_trials) return 1;
        s.alpha = (inl && a < 1e-4) ? 0.5 : ls_par_floor(p, a); // :120 / par :223-227
        return 0;
    }
    case LS_WOLFE: {
        // seq/line_search.cpp:143-188 ; par/line_search.cpp:317-368
        double a;
        if (f_new > s.f_x + p.c1 * alpha * s.gd || (f_new >= s.f_lo && iter > 0)) {
            s.hi = alpha;
            a = ls_interp(p, s.lo, s.hi, s.f_lo, s.dphi_lo, f_new,
                          (f_new - s.f_x - s.gd * alpha) / (alpha * alpha));
            s.alpha = a; // `continue`: skips the alpha < MIN test
            return s.trials < p.max_trials ? 1 : 0;
        }
        if (fabs(dphi_new) <= -p.c2 * s.gd) { s.success = 1; return 0; } // strong Wolfe: accept alpha
        if (dphi_new >= 0) {
            s.hi = alpha;
            a = ls_interp(p, s.lo, s.hi, s.f_lo, s.dphi_lo, f_new, dphi_new);
        } else {
            s.lo = alpha;
            s.f_lo = f_new;
            s.dphi_lo = dphi_new;
            if (s.hi == INFINITY)
                a = alpha * 2;
            else
                a = ls_interp(p, s.lo, s.hi, s.f_lo, s.dphi_lo, f_new, dphi_new);
        }
        if (a < p.wolfe_min) { s.alpha = p.wolfe_min; return 0; }
        s.alpha = a;
        return s.trials < p.max_trials ? 1 : 0;
    }
    default: { // LS_BACKTRACKING_WOLFE
        if (p.flavor == FLAVOR_SEQ) {
            // seq/line_search.cpp:37-53: x0.5 on Armijo failure, x1.1 on curvature failure, C2=0.9.
            // The reference loop is unbounded; ls_max_trials*5000 bounds it here.
            double a = alpha;
            if (f_new > s.f_x + p.c1 * alpha * s.gd) a = alpha * p.shrink;
            else if (dphi_new < p.c2 * s.gd) a = alpha * 1.1;
            else { s.success = 1; return 0; }
            s.alpha = a;
            if (a < p.bt_tol) return 0;
            return s.trials < p.max_trials * 5000 ? 1 : 0;
        }
        // par/line_search.cpp:52-153: bisection, local constants C1=1e-4, C2=0.9, TOL=1e-10
        const double DMAX = 1.7976931348623157e308;
        if (f_new <= s.f_x + 1e-4 * alpha * s.gd) {
            if (dphi_new >= 0.9 * s.gd) { s.success = 1; return 0; }
            s.lo = alpha;
        } else {
            s.hi = alpha;
        }
        double a;
        if (s.hi < DMAX) a = (s.lo + s.hi) / 2.0;
        else a = 2.0 * s.lo;
        s.alpha = a;
        if (a < 1e-10) {
            // the inlined copy steps to exactly TOL and evaluates f and the gradient there
            // (par/L-BFGS-Backtracking_Wolfe.cu:371-396), so x_host ends up AT the returned step
            if (p.flavor == FLAVOR_PAR_INLINED) { s.alpha = 1e-10; s.alpha_last = 1e-10; }
            return 0;
        }
        return s.trials < 20 ? 1 : 0;
    }
    }
}

// Consume the evaluation at s.alpha.  Returns 1: evaluate the new s.alpha; 0: finished,
// s.alpha is the step the search returns.
LB_HD int ls_step(const LsParams &p, LsState &s, double f_new, double dphi_new)
{
    s.alpha_last = s.alpha;
    s.f_last = f_new;
    const int cont = ls_step_core(p, s, f_new, dphi_new);
    if (!cont) s.stale = (s.alpha != s.alpha_last);
    return cont;
}

} // namespace lb
