// oracle/ref_driver.cpp -- C-callable driver around the UNMODIFIED reference.
//
// TEST INFRASTRUCTURE ONLY.  oracle/Makefile compiles this file together with the
// reference's own sources, taken where they lie under /root/reference (never
// copied), into oracle/_ref/libref_seq.so and oracle/_ref/libref_hybrid.so:
//
//   libref_seq.so    = seq/lbfgs.cpp + seq/vector_utils.cpp + seq/line_search.cpp
//                      + par/functions.cpp + seq/benchmark.cpp (tridiagonal generator)
//   libref_hybrid.so = seq/lbfgs.cpp + seq/vector_utils.cpp + par/line_search.cpp
//                      (+ par/constants.h as config.h)   -- SURVEY.md 8(c) "hybrid oracle":
//                      the sequential outer loop with the CUDA tree's line searches.
//
// Nothing here restates the algorithm; it only calls the reference's symbols.
#include <chrono>
#include <cstring>
#include <functional>
#include <iostream>
#include <random>
#include <sstream>
#include <string>
#include <vector>

using namespace std;

// ---- the reference's symbols (declared, not defined, here) ----
// seq/lbfgs.cpp:17-25
vector<double> LBFGS(const function<double(vector<double>)> f,
                     const function<vector<double>(vector<double>)> grad, const vector<double> x0,
                     const string line_search_method, const int max_iterations, const int m,
                     const double tolerance, const bool verbose);
// seq/vector_utils.cpp:32-41, :78-86
double dotProduct(const vector<double> &v1, const vector<double> &v2);
double vectorNorm(const vector<double> &v);
// par/functions.cpp:6-49
double quadratic(const vector<double> &X);
vector<double> quadratic_grad(const vector<double> &X);
double rosenbrock(const vector<double> &X);
vector<double> rosenbrock_grad(const vector<double> &X);
// seq/benchmark.cpp:16-56
std::function<double(const std::vector<double> &)> generate_quadratic_function(int n);
std::function<std::vector<double>(const std::vector<double> &)> generate_quadratic_gradient(int n);
// seq/line_search.cpp:8-16 / par/line_search.cpp:10-20
double cubicInterpolate(double, double, double, double, double, double);
double quadraticInterpolate(double, double, double, double, double);
#ifdef REF_HYBRID
// par/line_search.cpp:231-296
double safeCubicInterpolate(double, double, double, double, double, double);
#endif

namespace {
const char *kMethods[] = { "backtracking", "interpolation", "wolfe", "backtracking_wolfe" };

struct Objective {
    function<double(vector<double>)> f;
    function<vector<double>(vector<double>)> g;
};

Objective make_objective(int objective, size_t n)
{
    Objective o;
    if (objective == 0) {
        o.f = quadratic;
        o.g = quadratic_grad;
    } else if (objective == 1) {
        o.f = rosenbrock;
        o.g = rosenbrock_grad;
    } else {
        o.f = generate_quadratic_function((int)n);
        o.g = generate_quadratic_gradient((int)n);
    }
    return o;
}
} // namespace

extern "C" {

// Runs the reference LBFGS as-is.  status: 0 "Converged!", 1 "Maximum iterations
// reached", 2 "Line search failed" (parsed from the reference's stdout, the only
// place it reports them: seq/lbfgs.cpp:82, :166, :201).  iter_seconds (may be NULL,
// capacity iter_cap) receives the wall-clock time stamp of each outer-loop gradient
// evaluation, for the CPU baseline.
int ref_lbfgs(int objective, int line_search, size_t n, const double *x0, int max_it, int m,
              double tol, double *x_out, long *f_evals, long *g_evals, double *seconds)
{
    Objective o = make_objective(objective, n);
    long nf = 0, ng = 0;
    function<double(vector<double>)> f = [&](vector<double> x) { ++nf; return o.f(x); };
    function<vector<double>(vector<double>)> g = [&](vector<double> x) { ++ng; return o.g(x); };
    vector<double> x(x0, x0 + n);
    ostringstream captured;
    streambuf *old = cout.rdbuf(captured.rdbuf());
    auto t0 = chrono::steady_clock::now();
    vector<double> r;
    try {
        r = LBFGS(f, g, x, kMethods[line_search & 3], max_it, m, tol, false);
    } catch (...) {
        cout.rdbuf(old);
        return -1;
    }
    auto t1 = chrono::steady_clock::now();
    cout.rdbuf(old);
    if (seconds) *seconds = chrono::duration<double>(t1 - t0).count();
    if (x_out) memcpy(x_out, r.data(), n * sizeof(double));
    if (f_evals) *f_evals = nf;
    if (g_evals) *g_evals = ng;
    const string s = captured.str();
    if (s.find("Line search failed") != string::npos) return 2;
    if (s.find("Converged!") != string::npos) return 0;
    return 1;
}

// The reference's own stdout for a verbose run (seq/lbfgs.cpp:76-78, :82, :166, :201), copied into
// out (capacity cap, NUL-terminated).  Used to generate tests/golden/verbose_*.txt.
int ref_lbfgs_verbose(int objective, int line_search, size_t n, const double *x0, int max_it, int m, double tol,
                      char *out, size_t cap)
{
    Objective o = make_objective(objective, n);
    vector<double> x(x0, x0 + n);
    ostringstream captured;
    streambuf *old = cout.rdbuf(captured.rdbuf());
    try {
        LBFGS(o.f, o.g, x, kMethods[line_search & 3], max_it, m, tol, true);
    } catch (...) {
        cout.rdbuf(old);
        return -1;
    }
    cout.rdbuf(old);
    const string s = captured.str();
    if (s.size() + 1 > cap) return -2;
    memcpy(out, s.c_str(), s.size() + 1);
    return (int)s.size();
}

double ref_dot(const double *a, const double *b, size_t n)
{
    return dotProduct(vector<double>(a, a + n), vector<double>(b, b + n));
}
double ref_norm(const double *a, size_t n) { return vectorNorm(vector<double>(a, a + n)); }

double ref_f(int objective, const double *x, size_t n)
{
    return make_objective(objective, n).f(vector<double>(x, x + n));
}
void ref_grad(int objective, const double *x, double *g, size_t n)
{
    vector<double> r = make_objective(objective, n).g(vector<double>(x, x + n));
    memcpy(g, r.data(), n * sizeof(double));
}

double ref_cubic(double a0, double a1, double p0, double dp0, double p1, double dp1)
{
    return cubicInterpolate(a0, a1, p0, dp0, p1, dp1);
}
double ref_quadratic(double a0, double a1, double p0, double dp0, double p1)
{
    return quadraticInterpolate(a0, a1, p0, dp0, p1);
}
double ref_safe_cubic(double a0, double a1, double p0, double dp0, double p1, double dp1)
{
#ifdef REF_HYBRID
    return safeCubicInterpolate(a0, a1, p0, dp0, p1, dp1);
#else
    (void)a0; (void)a1; (void)p0; (void)dp0; (void)p1; (void)dp1;
    return 0.0 / 0.0;
#endif
}

// x0 exactly as the reference mains draw it: seq/main.cpp:34-43, par/L-BFGS-Wolfe.cu:458-465
void ref_x0(unsigned seed, double lo, double hi, size_t n, double *out)
{
    std::mt19937 gen(seed);
    std::uniform_real_distribution<> dis(lo, hi);
    for (size_t i = 0; i < n; ++i)
        out[i] = dis(gen);
}

int ref_is_hybrid(void)
{
#ifdef REF_HYBRID
    return 1;
#else
    return 0;
#endif
}

} // extern "C"
