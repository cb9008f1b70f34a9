// tests/compat_host_check.cpp -- the HOST-side behaviour of the C++ shim (include/lbfgsb200_compat.hpp), runnable
// without a GPU: which built-in device objective a std::function pair is recognised as, and the reference's
// exceptions.  Objectives are written as a caller would (from their definitions), the
// tridiagonal one like sequential-implementation/benchmark.cpp:16-56 (capturing n, asserting on the dimension).
// Prints one "name: result" line per check; tests/test_host_logic.py compares them.
#define LBFGSB200_COMPAT_IMPLEMENTATION
#include "../include/lbfgsb200_compat.hpp"

#include <cmath>
#include <cstdio>
#include <iostream>
#include <stdexcept>

#include "../oracle/lbfgs_oracle.h" // test infrastructure: supplies a trace for print_cuda_log

using std::function;
using std::string;
using std::vector;

vector<double> LBFGS(const function<double(vector<double>)> f, const function<vector<double>(vector<double>)> grad,
                     const vector<double> x0, const string line_search_method, const int max_iterations, const int m,
                     const double tolerance, bool verbose);
vector<double> LBFGS_CUDA(const function<double(vector<double>)> f, const function<vector<double>(vector<double>)> grad,
                          const vector<double> x0, const string line_search_method, const int max_iterations, const int m,
                          const double tolerance);

static double quad(vector<double> x) { double s = 0; for (double v : x) s += (v - 1) * (v - 1); return s; }
static vector<double> quad_g(vector<double> x) { vector<double> g(x.size()); for (size_t i = 0; i < x.size(); ++i) g[i] = 2.0 * (x[i] - 1); return g; }
static double rosen(vector<double> x)
{
    double s = 0;
    for (size_t i = 0; i + 1 < x.size(); ++i) { double t1 = x[i + 1] - x[i] * x[i], t2 = 1 - x[i]; s += 100.0 * t1 * t1 + t2 * t2; }
    return s;
}
static vector<double> rosen_g(vector<double> x)
{
    vector<double> g(x.size(), 0.0);
    for (size_t i = 0; i + 1 < x.size(); ++i) {
        double a = 2.0 * (x[i] - 1), b = x[i + 1] - x[i] * x[i];
        g[i] += a - 400.0 * x[i] * b;
        g[i + 1] += 200.0 * b;
    }
    return g;
}
// a generator that insists on its dimension exactly as the reference's does: assert() = abort, not an exception
static void insist(bool ok)
{
    if (!ok) {
        std::printf("ABORT: objective called with a vector of the wrong dimension\n");
        std::fflush(stdout);
        std::abort();
    }
}
static function<double(vector<double>)> tridiag_f(size_t n)
{
    return [n](vector<double> x) {
        insist(x.size() == n);
        double s = 0;
        for (size_t i = 0; i < n; ++i) s += 1000.0 * x[i] * x[i];
        for (size_t i = 0; i + 1 < n; ++i) s += (1000.0 / 10.0) * x[i] * x[i + 1];
        return s;
    };
}
static function<vector<double>(vector<double>)> tridiag_g(size_t n)
{
    return [n](vector<double> x) {
        insist(x.size() == n);
        vector<double> g(n);
        for (size_t i = 0; i < n; ++i) g[i] = 2.0 * 1000.0 * x[i];
        for (size_t i = 0; i + 1 < n; ++i) { g[i] += (1000.0 / 10.0) * x[i + 1]; g[i + 1] += (1000.0 / 10.0) * x[i]; }
        return g;
    };
}
static double quartic(vector<double> x) { double s = 0; for (double v : x) s += v * v * v * v; return s; }
static vector<double> quartic_g(vector<double> x) { for (double &v : x) v = 4 * v * v * v; return x; }

int main()
{
    using lbfgsb200::identify_objective;
    std::printf("quadratic: %d\n", identify_objective(quad, quad_g));
    std::printf("rosenbrock: %d\n", identify_objective(rosen, rosen_g));
    std::printf("tridiag6: %d\n", identify_objective(tridiag_f(6), tridiag_g(6)));
    std::printf("tridiag10000: %d\n", identify_objective(tridiag_f(10000), tridiag_g(10000), 10000));
    std::printf("rosenbrock4097: %d\n", identify_objective(rosen, rosen_g, 4097));
    std::printf("quadratic1: %d\n", identify_objective(quad, quad_g, 1));
    std::printf("quartic: %d\n", identify_objective(quartic, quartic_g));
    std::printf("mismatched_gradient: %d\n", identify_objective(rosen, quad_g));
    const vector<double> x0 = {0.5, -0.5, 1.5, 2.0, -1.0, 0.25};
    try {
        LBFGS(rosen, rosen_g, x0, "newton", 10, 5, 1e-5, false);
        std::printf("unknown_method: no exception\n");
    } catch (const std::invalid_argument &e) { // seq/lbfgs.cpp:69
        std::printf("unknown_method: invalid_argument %s\n", e.what());
    }
    try {
        LBFGS_CUDA(quartic, quartic_g, x0, "wolfe", 10, 5, 1e-5);
        std::printf("foreign_objective: no exception\n");
    } catch (const std::invalid_argument &e) {
        std::printf("foreign_objective: invalid_argument %s\n", std::string(e.what()).substr(0, 60).c_str());
    }
    // a dimension-bound objective through the entry point itself: recognised (probed at ITS dimension), then either
    // solved (GPU present) or refused because there is no device -- never "not a built-in", never an abort
    try {
        vector<double> x64(64, 0.5);
        vector<double> r = LBFGS(tridiag_f(64), tridiag_g(64), x64, "backtracking", 50, 10, 1e-5, false);
        std::printf("bound_objective: solved %zu\n", r.size());
    } catch (const std::invalid_argument &e) {
        std::printf("bound_objective: invalid_argument %s\n", e.what());
    } catch (const std::runtime_error &e) {
        std::printf("bound_objective: runtime_error %s\n", e.what());
    }
    // the CUDA tree's progress lines, rebuilt from a trace: the restatement of par/L-BFGS-Wolfe.cu on the golden
    // case wolfe_rosen_1e4 (the test compares this text with the stdout the reference printed on a B200)
    {
        const size_t n = 10000;
        const int K = 20;
        vector<double> xs(n), xo(n), tr((size_t)K * ORACLE_TRACE_COLS);
        oracle_x0(42, -2, 2, n, xs.data());
        oracle_params_t op = {ORACLE_OBJ_ROSENBROCK, ORACLE_LS_WOLFE, ORACLE_FLAVOR_PAR_INLINED, 10, K, 0.0};
        oracle_result_t res;
        const int status = oracle_lbfgs_cuda_profile(&op, n, xs.data(), xo.data(), tr.data(), K, &res);
        std::printf("cuda_log_begin: %d\n", status);
        std::fflush(stdout);
        lbfgsb200::print_cuda_log(std::cout, tr.data(), K, res.iterations, status);
        std::cout.flush();
        std::printf("cuda_log_end: %ld\n", res.iterations);
    }
    // naming the objective skips the probing altogether
    lbfgsb200::compat_options().objective = LBFGSB200_OBJ_ROSENBROCK;
    try {
        int calls = 0;
        function<double(vector<double>)> counted = [&](vector<double> x) { ++calls; return rosen(x); };
        LBFGS(counted, rosen_g, x0, "newton", 10, 5, 1e-5, false);
    } catch (const std::invalid_argument &e) {
        std::printf("named_objective: invalid_argument %s\n", e.what());
    }
    return 0;
}
