"""ctypes access to the CPU checker.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  Two things live behind it:

* ``Oracle``  -- oracle/liblbfgs_oracle.so, the C restatement (oracle/lbfgs_oracle.c).
* ``Ref``     -- oracle/_ref/libref_{seq,hybrid}.so, the UNMODIFIED reference sources
                 compiled by oracle/Makefile (present only after ``make ref`` ran in the
                 build container; the .so files travel to the GPU box).
* ``CudaRef`` -- oracle/_ref/libref_cuda_*.so, the reference's UNMODIFIED CUDA solvers
                 (parallel-implementation/*.cu) cross-compiled for sm_100 by ``make cudaref``;
                 they need a GPU, so only ``-m gpu`` tests and oracle/make_golden_cuda.py run them.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = "/root/reference"

OBJ = {"quadratic": 0, "rosenbrock": 1, "tridiag": 2}
LS = {"backtracking": 0, "interpolation": 1, "wolfe": 2, "backtracking_wolfe": 3}
FLAVOR = {"seq": 0, "par": 1, "par_inlined": 2, "par_stale_gradient": 3}
TRACE_COLS = 8
TR_K, TR_F, TR_GNORM, TR_ALPHA, TR_TRIALS, TR_HIST, TR_X0, TR_XMID = range(8)

_dp = C.POINTER(C.c_double)


def _p(a):
    return a.ctypes.data_as(_dp)


def build(ref=True):
    """Compile the restatement and, where /root/reference exists, the reference itself."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref and os.path.isdir(REFERENCE_ROOT):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref", "refmain", "cudaref", "cudamain"])


class _Params(C.Structure):
    _fields_ = [("objective", C.c_int), ("line_search", C.c_int), ("flavor", C.c_int),
                ("m", C.c_int), ("max_iterations", C.c_int), ("tolerance", C.c_double)]


class _Result(C.Structure):
    _fields_ = [("status", C.c_int), ("iterations", C.c_long), ("f_evals", C.c_long),
                ("g_evals", C.c_long), ("f", C.c_double), ("gnorm", C.c_double)]


class Oracle:
    def __init__(self):
        path = os.path.join(HERE, "liblbfgs_oracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = self.L = C.CDLL(path)
        L.oracle_lbfgs.restype = C.c_int
        L.oracle_lbfgs.argtypes = [C.POINTER(_Params), C.c_size_t, _dp, _dp, _dp, C.c_size_t,
                                   C.POINTER(_Result)]
        L.oracle_lbfgs_cuda_profile.restype = C.c_int
        L.oracle_lbfgs_cuda_profile.argtypes = L.oracle_lbfgs.argtypes
        L.oracle_dot.restype = C.c_double
        L.oracle_dot.argtypes = [_dp, _dp, C.c_size_t]
        L.oracle_norm.restype = C.c_double
        L.oracle_norm.argtypes = [_dp, C.c_size_t]
        L.oracle_f.restype = C.c_double
        L.oracle_f.argtypes = [C.c_int, _dp, C.c_size_t]
        L.oracle_grad.restype = None
        L.oracle_grad.argtypes = [C.c_int, _dp, _dp, C.c_size_t]
        L.oracle_two_loop.restype = C.c_int
        L.oracle_two_loop.argtypes = [_dp, _dp, _dp, C.c_int, C.c_size_t, _dp]
        for name in ("oracle_cubic", "oracle_safe_cubic"):
            fn = getattr(L, name)
            fn.restype = C.c_double
            fn.argtypes = [C.c_double] * 6
        L.oracle_quadratic.restype = C.c_double
        L.oracle_quadratic.argtypes = [C.c_double] * 5
        L.oracle_ls_poly.restype = C.c_double
        L.oracle_ls_poly.argtypes = [C.c_int, C.c_int, _dp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.oracle_x0.restype = None
        L.oracle_x0.argtypes = [C.c_uint, C.c_double, C.c_double, C.c_size_t, _dp]

    def x0(self, n, lo, hi, seed=42):
        out = np.empty(n, dtype=np.float64)
        self.L.oracle_x0(seed, lo, hi, n, _p(out))
        return out

    def set_exact_sums(self, on):
        """TEST-ONLY: accumulate every reduction in long double (explains the summation noise of the reference)."""
        self.L.oracle_set_exact_sums.restype = None
        self.L.oracle_set_exact_sums.argtypes = [C.c_int]
        self.L.oracle_set_exact_sums(1 if on else 0)

    def dot(self, a, b):
        return self.L.oracle_dot(_p(a), _p(b), a.size)

    def norm(self, a):
        return self.L.oracle_norm(_p(a), a.size)

    def f(self, objective, x):
        return self.L.oracle_f(OBJ[objective], _p(x), x.size)

    def grad(self, objective, x):
        g = np.empty_like(x)
        self.L.oracle_grad(OBJ[objective], _p(x), _p(g), x.size)
        return g

    def two_loop(self, g, S, Y):
        """S, Y: (h, n) arrays, oldest pair first.  Returns (d, fell_back)."""
        h = S.shape[0]
        S = np.ascontiguousarray(S)
        Y = np.ascontiguousarray(Y)
        d = np.empty_like(g)
        rc = self.L.oracle_two_loop(_p(g), _p(S), _p(Y), h, g.size, _p(d))
        return d, bool(rc)

    def cubic(self, *a):
        return self.L.oracle_cubic(*a)

    def safe_cubic(self, *a):
        return self.L.oracle_safe_cubic(*a)

    def quadratic_interp(self, *a):
        return self.L.oracle_quadratic(*a)

    def ls_poly(self, line_search, flavor, coef):
        c = np.ascontiguousarray(coef, dtype=np.float64)
        nf, ng = C.c_int(0), C.c_int(0)
        a = self.L.oracle_ls_poly(LS[line_search], FLAVOR[flavor], _p(c), C.byref(nf), C.byref(ng))
        return a, nf.value, ng.value

    def ls_poly_inlined(self, line_search, coef, f_xhost, f_initial):
        """-> (alpha, trials, grad_evals, success, f at the last evaluated point)"""
        coef = np.ascontiguousarray(coef, dtype=np.float64)
        nf, ng, ok, fl = C.c_int(), C.c_int(), C.c_int(), C.c_double()
        self.L.oracle_ls_poly_inlined.restype = C.c_double
        self.L.oracle_ls_poly_inlined.argtypes = [C.c_int, _dp, C.c_double, C.c_double, C.POINTER(C.c_int),
                                                  C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double)]
        a = self.L.oracle_ls_poly_inlined(LS[line_search], _p(coef), f_xhost, f_initial, C.byref(nf), C.byref(ng),
                                          C.byref(ok), C.byref(fl))
        return a, nf.value, ng.value, ok.value, fl.value

    def lbfgs(self, objective, x0, line_search="backtracking", flavor="seq", m=10,
              max_iterations=1000, tolerance=1e-5, trace_rows=0, profile="seq"):
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        p = _Params(OBJ[objective], LS[line_search], FLAVOR[flavor], m, max_iterations, tolerance)
        r = _Result()
        x = np.empty_like(x0)
        trace = np.zeros((max(trace_rows, 1), TRACE_COLS), dtype=np.float64)
        fn = self.L.oracle_lbfgs if profile == "seq" else self.L.oracle_lbfgs_cuda_profile
        fn(C.byref(p), x0.size, _p(x0), _p(x), _p(trace) if trace_rows else None, trace_rows, C.byref(r))
        info = dict(status=r.status, iterations=r.iterations, f_evals=r.f_evals,
                    g_evals=r.g_evals, f=r.f, gnorm=r.gnorm)
        rows = min(trace_rows, r.iterations)
        return x, info, trace[:rows]


class Ref:
    """The unmodified reference (seq outer loop; flavor picks the line-search tree)."""

    def __init__(self, flavor="seq"):
        name = {"seq": "libref_seq.so", "par": "libref_hybrid.so"}[flavor]
        path = os.path.join(HERE, "_ref", name)
        if not os.path.exists(path):
            if not os.path.isdir(REFERENCE_ROOT):
                raise FileNotFoundError(path + " (oracle/_ref is only buildable where /root/reference exists)")
            build(ref=True)
        L = self.L = C.CDLL(path)
        L.ref_lbfgs.restype = C.c_int
        L.ref_lbfgs.argtypes = [C.c_int, C.c_int, C.c_size_t, _dp, C.c_int, C.c_int, C.c_double, _dp,
                                C.POINTER(C.c_long), C.POINTER(C.c_long), _dp]
        L.ref_dot.restype = C.c_double
        L.ref_dot.argtypes = [_dp, _dp, C.c_size_t]
        L.ref_norm.restype = C.c_double
        L.ref_norm.argtypes = [_dp, C.c_size_t]
        L.ref_f.restype = C.c_double
        L.ref_f.argtypes = [C.c_int, _dp, C.c_size_t]
        L.ref_grad.restype = None
        L.ref_grad.argtypes = [C.c_int, _dp, _dp, C.c_size_t]
        for name in ("ref_cubic", "ref_safe_cubic"):
            fn = getattr(L, name)
            fn.restype = C.c_double
            fn.argtypes = [C.c_double] * 6
        L.ref_quadratic.restype = C.c_double
        L.ref_quadratic.argtypes = [C.c_double] * 5
        L.ref_x0.restype = None
        L.ref_x0.argtypes = [C.c_uint, C.c_double, C.c_double, C.c_size_t, _dp]

    @staticmethod
    def available(flavor="seq"):
        name = {"seq": "libref_seq.so", "par": "libref_hybrid.so"}[flavor]
        return os.path.exists(os.path.join(HERE, "_ref", name))

    def x0(self, n, lo, hi, seed=42):
        out = np.empty(n, dtype=np.float64)
        self.L.ref_x0(seed, lo, hi, n, _p(out))
        return out

    def dot(self, a, b):
        return self.L.ref_dot(_p(a), _p(b), a.size)

    def norm(self, a):
        return self.L.ref_norm(_p(a), a.size)

    def f(self, objective, x):
        return self.L.ref_f(OBJ[objective], _p(x), x.size)

    def grad(self, objective, x):
        g = np.empty_like(x)
        self.L.ref_grad(OBJ[objective], _p(x), _p(g), x.size)
        return g

    def lbfgs(self, objective, x0, line_search="backtracking", m=10, max_iterations=1000,
              tolerance=1e-5):
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        x = np.empty_like(x0)
        nf, ng = C.c_long(0), C.c_long(0)
        sec = C.c_double(0)
        st = self.L.ref_lbfgs(OBJ[objective], LS[line_search], x0.size, _p(x0), max_iterations, m,
                              tolerance, _p(x), C.byref(nf), C.byref(ng), C.byref(sec))
        return x, dict(status=st, f_evals=nf.value, g_evals=ng.value, seconds=sec.value)


class CudaRef:
    """The reference's own CUDA solver (one of parallel-implementation/*.cu), run as-is on the GPU."""

    VARIANTS = {"host": "par/L-BFGS.cu", "wolfe": "par/L-BFGS-Wolfe.cu", "backtracking": "par/L-BFGS-Backtracking.cu",
                "interpolation": "par/L-BFGS-Interpolation.cu", "btwolfe": "par/L-BFGS-Backtracking_Wolfe.cu"}

    def __init__(self, variant):
        assert variant in self.VARIANTS, variant
        path = os.path.join(HERE, "_ref", "libref_cuda_%s.so" % variant)
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle cudaref` in the build container)")
        self.variant = variant
        L = self.L = C.CDLL(path)
        L.ref_cuda_lbfgs.restype = C.c_int
        L.ref_cuda_lbfgs.argtypes = [C.c_int, C.c_char_p, C.c_size_t, _dp, C.c_int, C.c_int, C.c_double, _dp,
                                     C.POINTER(C.c_long), C.POINTER(C.c_long), C.c_char_p, C.c_size_t]

    @staticmethod
    def available(variant):
        return os.path.exists(os.path.join(HERE, "_ref", "libref_cuda_%s.so" % variant))

    def own_main(self):
        """Run the .cu file's own main() (the program the reference ships) and return its stdout."""
        cap = 1 << 26
        log = C.create_string_buffer(cap)
        self.L.ref_cuda_own_main.restype = C.c_int
        self.L.ref_cuda_own_main.argtypes = [C.c_char_p, C.c_size_t]
        rc = self.L.ref_cuda_own_main(log, cap)
        if rc != 0:
            raise RuntimeError("the reference's main() returned %d" % rc)
        return log.value.decode(errors="replace")

    def lbfgs(self, objective, x0, line_search="wolfe", m=10, max_iterations=1000, tolerance=1e-5):
        """Returns (x, info); info["log"] is the reference's stdout, info["alphas"] / ["gnorms"] are parsed
        from its "alpha: ..." / "Iteration k: norm_g = ..." lines (6 significant digits, as printed)."""
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        x = np.empty_like(x0)
        nf, ng = C.c_long(), C.c_long()
        cap = 1 << 22
        log = C.create_string_buffer(cap)
        rc = self.L.ref_cuda_lbfgs(OBJ[objective], line_search.encode(), x0.size, _p(x0), max_iterations, m, tolerance,
                                   _p(x), C.byref(nf), C.byref(ng), log, cap)
        if rc != 0:
            raise RuntimeError("the CUDA reference threw")
        text = log.value.decode(errors="replace")
        alphas, gnorms = [], []
        for line in text.splitlines():
            if line.startswith("alpha: "):
                alphas.append(float(line.split()[1]))
            elif line.startswith("Iteration ") and "norm_g" in line:
                gnorms.append(float(line.rsplit("=", 1)[1]))
        status = 2 if "Line search failed" in text else (0 if "Convergence achieved" in text else 1)
        return x, dict(status=status, f_evals=nf.value, g_evals=ng.value, log=text, alphas=alphas, gnorms=gnorms)
