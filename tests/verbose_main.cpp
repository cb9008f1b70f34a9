// tests/verbose_main.cpp -- a caller written like the reference's mains (objective functions as
// plain C++ functions over std::vector, x0 from mt19937(42)), calling LBFGS(..., verbose=true)
// through include/lbfgsb200_compat.hpp.  tests/test_gpu_compat.py compares its stdout with the
// reference's own verbose stdout (tests/golden/verbose_*.txt).
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iostream>
#include <string>
#include <vector>

#define LBFGSB200_COMPAT_IMPLEMENTATION
#include "../include/lbfgsb200_compat.hpp"

using std::vector;

// user-side objectives (what a caller of the reference writes; cf. parallel-implementation/functions.cpp)
static double my_rosenbrock(const vector<double> &X)
{
    double s = 0.0;
    for (size_t i = 0; i + 1 < X.size(); i++) {
        double t1 = X[i + 1] - X[i] * X[i], t2 = 1 - X[i];
        s += 100.0 * t1 * t1 + t2 * t2;
    }
    return s;
}
static vector<double> my_rosenbrock_grad(const vector<double> &X)
{
    vector<double> g(X.size(), 0.0);
    for (size_t i = 0; i + 1 < X.size(); i++) {
        double t1 = 2.0 * (X[i] - 1), t2 = X[i + 1] - X[i] * X[i];
        g[i] += t1 - 400.0 * X[i] * t2;
        g[i + 1] += 200.0 * t2;
    }
    return g;
}
static double my_tridiag(const vector<double> &x)
{
    double r = 0.0;
    for (size_t i = 0; i < x.size(); i++) r += 1000.0 * x[i] * x[i];
    for (size_t i = 0; i + 1 < x.size(); i++) r += 100.0 * x[i] * x[i + 1];
    return r;
}
static vector<double> my_tridiag_grad(const vector<double> &x)
{
    vector<double> g(x.size(), 0.0);
    for (size_t i = 0; i < x.size(); i++) g[i] = 2000.0 * x[i];
    for (size_t i = 0; i + 1 < x.size(); i++) {
        g[i] += 100.0 * x[i + 1];
        g[i + 1] += 100.0 * x[i];
    }
    return g;
}
static double my_unknown(const vector<double> &x) { return x.empty() ? 0.0 : x[0] * x[0] * x[0]; }
static vector<double> my_unknown_grad(const vector<double> &x) { return vector<double>(x.size(), 1.0); }

int main(int argc, char **argv)
{
    const std::string which = argc > 1 ? argv[1] : "rosen5";
    if (which == "rosen5" || which == "tridiag64") {
        const size_t n = which == "rosen5" ? 5 : 64;
        vector<double> x0(n);
        lbfgsb200_x0_uniform(42, -2, 2, 0, n, x0.data());
        if (which == "rosen5") LBFGS(my_rosenbrock, my_rosenbrock_grad, x0, "backtracking", 12, 10, 1e-5, true);
        else LBFGS(my_tridiag, my_tridiag_grad, x0, "interpolation", 20, 10, 1e-5, true);
        return 0;
    }
    if (which == "unknown_objective") {
        try {
            LBFGS(my_unknown, my_unknown_grad, vector<double>(8, 1.0), "backtracking", 5, 3, 1e-5, false);
        } catch (const std::invalid_argument &e) {
            std::cout << "invalid_argument: " << e.what() << std::endl;
            return 3;
        }
        return 0;
    }
    if (which == "unknown_method") {
        try {
            LBFGS(my_rosenbrock, my_rosenbrock_grad, vector<double>(8, 0.5), "newton", 5, 3, 1e-5, false);
        } catch (const std::invalid_argument &e) {
            std::cout << "invalid_argument: " << e.what() << std::endl;
            return 4;
        }
        return 0;
    }
    return 2;
}
