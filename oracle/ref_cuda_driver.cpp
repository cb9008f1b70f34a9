// oracle/ref_cuda_driver.cpp -- C-callable driver around the UNMODIFIED CUDA tree of the reference.
//
// TEST INFRASTRUCTURE ONLY.  oracle/Makefile (target `cudaref`) compiles each of the reference's
// parallel-implementation/*.cu solvers where it lies under /root/reference, with the reference's own
// command line (par/run.sh: `nvcc <file>.cu functions.cpp line_search.cpp vector_utils.cpp -lcublas`, nvcc
// defaults, i.e. -fmad=true) except for the architecture (sm_100 instead of sm_75) and `-Dmain=...` (each
// file carries its own main()), together with this driver, into
//
//   oracle/_ref/libref_cuda_host.so          par/L-BFGS.cu               (host line search, takes its name)
//   oracle/_ref/libref_cuda_wolfe.so         par/L-BFGS-Wolfe.cu         (inlined Wolfe)
//   oracle/_ref/libref_cuda_backtracking.so  par/L-BFGS-Backtracking.cu  (inlined Armijo backtracking)
//   oracle/_ref/libref_cuda_interpolation.so par/L-BFGS-Interpolation.cu (inlined Armijo interpolation)
//   oracle/_ref/libref_cuda_btwolfe.so       par/L-BFGS-Backtracking_Wolfe.cu (inlined bisection Wolfe)
//
// They are built in the (GPU-less) build container and RUN on the GPU box by tests/test_gpu_cuda_reference.py:
// the reference's CUDA solver and this repository's solver execute side by side on the same B200.
// Nothing here restates the algorithm; it only calls the reference's LBFGS_CUDA.
#include <cstring>
#include <functional>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

using namespace std;

#ifdef REF_CUDA_HOST_LS
// par/L-BFGS.cu:105-112
vector<double> LBFGS_CUDA(const function<double(vector<double>)> f,
                          const function<vector<double>(vector<double>)> grad, const vector<double> x0,
                          const string line_search_method, const int max_iterations, const int m,
                          const double tolerance);
#else
// par/L-BFGS-Wolfe.cu:105-111 and the other inlined variants
vector<double> LBFGS_CUDA(const function<double(vector<double>)> f,
                          const function<vector<double>(vector<double>)> grad, const vector<double> x0,
                          const int max_iterations, const int m, const double tolerance);
#endif
// par/functions.cpp:6-49
double quadratic(const vector<double> &X);
vector<double> quadratic_grad(const vector<double> &X);
double rosenbrock(const vector<double> &X);
vector<double> rosenbrock_grad(const vector<double> &X);

int reference_own_main(); // the .cu file's main(), renamed at compile time

extern "C" {

// Runs the reference's CUDA solver as-is.  objective: 0 quadratic, 1 rosenbrock.  line_search is used by
// the host-line-search build only.  The reference's stdout ("alpha: ...", "Iteration k: norm_g = ...",
// "Optimum value: ...", warnings) is captured into log (capacity log_cap, truncated, NUL-terminated).
// Returns 0, or -1 when the reference threw.  (On a CUDA error the reference calls exit(1) itself.)
int ref_cuda_lbfgs(int objective, const char *line_search, size_t n, const double *x0, int max_it, int m, double tol,
                   double *x_out, long *f_evals, long *g_evals, char *log, size_t log_cap)
{
    long nf = 0, ng = 0;
    function<double(vector<double>)> f = [&](vector<double> x) { ++nf; return objective == 0 ? quadratic(x) : rosenbrock(x); };
    function<vector<double>(vector<double>)> g = [&](vector<double> x) { ++ng; return objective == 0 ? quadratic_grad(x) : rosenbrock_grad(x); };
    vector<double> x(x0, x0 + n);
    ostringstream captured;
    streambuf *old = cout.rdbuf(captured.rdbuf());
    vector<double> r;
    try {
#ifdef REF_CUDA_HOST_LS
        r = LBFGS_CUDA(f, g, x, line_search ? line_search : "wolfe", max_it, m, tol);
#else
        (void)line_search;
        r = LBFGS_CUDA(f, g, x, max_it, m, tol);
#endif
    } catch (...) {
        cout.rdbuf(old);
        return -1;
    }
    cout.rdbuf(old);
    if (x_out) memcpy(x_out, r.data(), n * sizeof(double));
    if (f_evals) *f_evals = nf;
    if (g_evals) *g_evals = ng;
    if (log && log_cap) {
        const string s = captured.str();
        const size_t c = s.size() < log_cap - 1 ? s.size() : log_cap - 1;
        memcpy(log, s.data(), c);
        log[c] = 0;
    }
    return 0;
}

// The reference program itself: the file's own main() (renamed by -Dmain=reference_own_main; par/L-BFGS-Wolfe.cu:456-484
// and its siblings: x0 ~ U(-2,2) from mt19937(42), n = 50 000, LBFGS_CUDA(rosenbrock, ..., 50000, 10, 1e-1), prints x0,
// the progress lines, the solution and "Optimum value").  stdout is captured into log.  Returns main()'s value.
int ref_cuda_own_main(char *log, size_t log_cap)
{
    ostringstream captured;
    streambuf *old = cout.rdbuf(captured.rdbuf());
    int rc = -1;
    try {
        rc = reference_own_main();
    } catch (...) {
        rc = -2;
    }
    cout.rdbuf(old);
    if (log && log_cap) {
        const string s = captured.str();
        const size_t c = s.size() < log_cap - 1 ? s.size() : log_cap - 1;
        memcpy(log, s.data(), c);
        log[c] = 0;
    }
    return rc;
}

int ref_cuda_has_line_search_argument(void)
{
#ifdef REF_CUDA_HOST_LS
    return 1;
#else
    return 0;
#endif
}

} // extern "C"
