"""ctypes harness over the C ABI of liblbfgsb200.so (include/lbfgsb200.h).

This is NOT the product: the product is the C-ABI library (hand-written sm_100a kernels + C++
host state machine).  Python only drives it for tests and bench.py, mirroring how a C++ host
would call it.  There is no CPU path here: if the shared library is missing it is built with
nvcc, and every compute call fails loudly on a machine without a CUDA device.

The directory name (``cuda-lbfgs_b200``) is not a valid Python identifier; load this module with
``importlib`` (see ``tests/conftest.py``, ``bench.py``, ``__graft_entry__.py``).
"""
import ctypes as C
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _load_build():
    spec = importlib.util.spec_from_file_location("lbfgsb200_build", os.path.join(HERE, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


build_module = _load_build()
LIB_PATH = build_module.LIB

OBJ = {"quadratic": 0, "rosenbrock": 1, "tridiag": 2}
LS = {"backtracking": 0, "interpolation": 1, "wolfe": 2, "backtracking_wolfe": 3}
FLAVOR = {"seq": 0, "par": 1, "par_inlined": 2}
PROFILE = {"seq": 0, "cuda": 1}
DIRECTION = {"two_loop": 0, "compact": 1, "auto": 2}
STATUS = {0: "converged", 1: "max_iter", 2: "ls_failed", 3: "running"}
TRACE_COLS = 8
UNIQUE_ID_BYTES = 128

_dp = C.POINTER(C.c_double)


class Params(C.Structure):
    _fields_ = [("m", C.c_int), ("max_iterations", C.c_int), ("tolerance", C.c_double),
                ("line_search", C.c_int), ("flavor", C.c_int), ("profile", C.c_int),
                ("direction", C.c_int), ("c1", C.c_double), ("c2", C.c_double),
                ("step0", C.c_double), ("shrink", C.c_double), ("backtracking_tol", C.c_double),
                ("wolfe_min", C.c_double), ("ls_max_trials", C.c_int), ("use_graph", C.c_int),
                ("verbose", C.c_int), ("grid_ctas", C.c_int), ("num_gpus", C.c_int)]


class Result(C.Structure):
    _fields_ = [("status", C.c_int), ("iterations", C.c_int64), ("trial_evals", C.c_int64),
                ("kernel_launches", C.c_int64), ("f", C.c_double), ("gnorm", C.c_double),
                ("device_ms", C.c_double), ("bytes_moved", C.c_double), ("f0", C.c_double),
                ("gnorm0", C.c_double), ("flow", C.c_int), ("graph", C.c_int), ("num_gpus", C.c_int), ("reserved", C.c_int)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class LbfgsError(RuntimeError):
    pass


_lib = None

# every symbol include/lbfgsb200.h declares (tests check the library exports all of them)
EXPORTS = [
    "lbfgsb200_create_callback", "lbfgsb200_create_callback_sharded", "lbfgsb200_device_halo", "lbfgsb200_checkpoint_save", "lbfgsb200_checkpoint_load", "lbfgsb200_version", "lbfgsb200_strerror", "lbfgsb200_last_error", "lbfgsb200_device_count",
    "lbfgsb200_params_default", "lbfgsb200_solve", "lbfgsb200_create", "lbfgsb200_set_x0",
    "lbfgsb200_iterate", "lbfgsb200_iterate_profiled", "lbfgsb200_get_x", "lbfgsb200_get_result",
    "lbfgsb200_get_trace", "lbfgsb200_local_size", "lbfgsb200_destroy", "lbfgsb200_shard_range",
    "lbfgsb200_comm_unique_id", "lbfgsb200_comm_create", "lbfgsb200_comm_destroy",
    "lbfgsb200_dot", "lbfgsb200_nrm2", "lbfgsb200_axpy", "lbfgsb200_scal", "lbfgsb200_eval_trial",
    "lbfgsb200_two_loop", "lbfgsb200_accept", "lbfgsb200_x0_uniform", "lbfgsb200_host_alloc",
    "lbfgsb200_host_free", "lbfgsb200_device_alloc", "lbfgsb200_device_free", "lbfgsb200_memcpy",
    "lbfgsb200_set_device", "lbfgsb200_device_sync", "lbfgsb200_trim_memory", "lbfgsb200_mem_info", "lbfgsb200_debug_timeline",
    "lbfgsb200_resolve_num_gpus", "lbfgsb200_comm_create_local",
]


def lib():
    """Load (building first if needed) the C-ABI library.  Never falls back to anything else."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("LBFGSB200_LIB") or build_module.build()  # LBFGSB200_LIB: an alternative build (kernel tuning A/B)
    L = C.CDLL(path)
    L.lbfgsb200_version.restype = C.c_int
    L.lbfgsb200_strerror.restype = C.c_char_p
    L.lbfgsb200_strerror.argtypes = [C.c_int]
    L.lbfgsb200_last_error.restype = C.c_char_p
    L.lbfgsb200_device_count.restype = C.c_int
    L.lbfgsb200_params_default.argtypes = [C.POINTER(Params), C.c_int]
    L.lbfgsb200_solve.argtypes = [C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.POINTER(Params),
                                  C.POINTER(Result), C.c_void_p, C.c_size_t]
    L.lbfgsb200_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_size_t, C.POINTER(Params),
                                   C.c_void_p, C.c_size_t]
    L.lbfgsb200_create_callback.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(Params),
                                            C.c_size_t]
    L.lbfgsb200_create_callback_sharded.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(Params),
                                                    C.c_void_p, C.c_size_t]
    L.lbfgsb200_device_halo.restype = C.c_void_p
    L.lbfgsb200_device_halo.argtypes = [C.c_void_p]
    L.lbfgsb200_set_x0.argtypes = [C.c_void_p, C.c_void_p]
    L.lbfgsb200_checkpoint_save.argtypes = [C.c_void_p, C.c_char_p]
    L.lbfgsb200_checkpoint_load.argtypes = [C.c_void_p, C.c_char_p]
    L.lbfgsb200_iterate.argtypes = [C.c_void_p, C.c_int64]
    L.lbfgsb200_iterate_profiled.argtypes = [C.c_void_p, C.c_int64, _dp, C.POINTER(C.c_int64)]
    L.lbfgsb200_get_x.argtypes = [C.c_void_p, C.c_void_p]
    L.lbfgsb200_get_result.argtypes = [C.c_void_p, C.POINTER(Result)]
    L.lbfgsb200_get_trace.restype = C.c_int64
    L.lbfgsb200_get_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.lbfgsb200_local_size.restype = C.c_size_t
    L.lbfgsb200_local_size.argtypes = [C.c_void_p]
    L.lbfgsb200_destroy.restype = None
    L.lbfgsb200_destroy.argtypes = [C.c_void_p]
    L.lbfgsb200_shard_range.restype = None
    L.lbfgsb200_shard_range.argtypes = [C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_size_t),
                                        C.POINTER(C.c_size_t)]
    L.lbfgsb200_comm_unique_id.argtypes = [C.c_char_p]
    L.lbfgsb200_comm_create.argtypes = [C.POINTER(C.c_void_p), C.c_char_p, C.c_int, C.c_int]
    L.lbfgsb200_comm_destroy.restype = None
    L.lbfgsb200_comm_destroy.argtypes = [C.c_void_p]
    L.lbfgsb200_dot.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    L.lbfgsb200_nrm2.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    L.lbfgsb200_axpy.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.lbfgsb200_scal.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.lbfgsb200_eval_trial.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                       C.c_void_p, C.c_void_p, C.c_void_p]
    L.lbfgsb200_two_loop.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_size_t,
                                     C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
    L.lbfgsb200_accept.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.lbfgsb200_x0_uniform.restype = None
    L.lbfgsb200_x0_uniform.argtypes = [C.c_uint, C.c_double, C.c_double, C.c_size_t, C.c_size_t,
                                       C.c_void_p]
    L.lbfgsb200_host_alloc.restype = C.c_void_p
    L.lbfgsb200_host_alloc.argtypes = [C.c_size_t]
    L.lbfgsb200_host_free.restype = None
    L.lbfgsb200_host_free.argtypes = [C.c_void_p]
    L.lbfgsb200_device_alloc.restype = C.c_void_p
    L.lbfgsb200_device_alloc.argtypes = [C.c_size_t]
    L.lbfgsb200_device_free.restype = None
    L.lbfgsb200_device_free.argtypes = [C.c_void_p]
    L.lbfgsb200_memcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.lbfgsb200_set_device.argtypes = [C.c_int]
    L.lbfgsb200_resolve_num_gpus.argtypes = [C.c_int, C.c_size_t]
    L.lbfgsb200_resolve_num_gpus.restype = C.c_int
    _lib = L
    return L


def _check(rc, what):
    if rc < 0:
        L = lib()
        raise LbfgsError("%s: %s (%s)" % (what, L.lbfgsb200_strerror(rc).decode(),
                                          L.lbfgsb200_last_error().decode()))
    return rc


def trim_memory():
    """Hand the arenas cached by destroyed solvers back to the driver (lbfgsb200_trim_memory)."""
    _check(lib().lbfgsb200_trim_memory(), "trim_memory")


def mem_info():
    """(free, total) device bytes; arenas cached by destroyed solvers count as used."""
    f, t = C.c_size_t(), C.c_size_t()
    _check(lib().lbfgsb200_mem_info(C.byref(f), C.byref(t)), "mem_info")
    return f.value, t.value


def default_params(flavor="seq", **overrides):
    p = Params()
    _check(lib().lbfgsb200_params_default(C.byref(p), FLAVOR[flavor]), "params_default")
    for k, v in overrides.items():
        if k == "line_search" and isinstance(v, str):
            v = LS[v]
        elif k == "profile" and isinstance(v, str):
            v = PROFILE[v]
        elif k == "direction" and isinstance(v, str):
            v = DIRECTION[v]
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


def shard_range(n, rank, nranks):
    off, ln = C.c_size_t(0), C.c_size_t(0)
    lib().lbfgsb200_shard_range(n, rank, nranks, C.byref(off), C.byref(ln))
    return off.value, ln.value


def x0_uniform(n, lo, hi, seed=42, offset=0, out=None):
    """x0 as the reference mains draw it (mt19937(seed) + uniform_real_distribution(lo,hi))."""
    if out is None:
        out = np.empty(n, dtype=np.float64)
    lib().lbfgsb200_x0_uniform(seed, lo, hi, offset, n, out.ctypes.data)
    return out


class PinnedArray:
    """float64 numpy view over cudaHostAlloc'd memory."""

    def __init__(self, n):
        self.n = n
        self.ptr = lib().lbfgsb200_host_alloc(max(n, 1) * 8)
        if not self.ptr:
            raise LbfgsError("host_alloc failed: " + lib().lbfgsb200_last_error().decode())
        self.array = np.ctypeslib.as_array(C.cast(self.ptr, _dp), shape=(n,))

    def free(self):
        if self.ptr:
            self.array = None
            lib().lbfgsb200_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        self.free()


class DeviceBuffer:
    """Raw device allocation of n doubles (cudaMalloc through the C ABI)."""

    def __init__(self, n, init=None):
        self.n = int(n)
        self.ptr = lib().lbfgsb200_device_alloc(max(self.n, 1) * 8)
        if not self.ptr:
            raise LbfgsError("device_alloc failed: " + lib().lbfgsb200_last_error().decode())
        if init is not None:
            self.upload(init)

    def upload(self, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64).ravel()
        assert a.size <= self.n
        _check(lib().lbfgsb200_memcpy(self.ptr, a.ctypes.data, a.size * 8), "memcpy H2D")

    def download(self, count=None):
        out = np.empty(self.n if count is None else count, dtype=np.float64)
        _check(lib().lbfgsb200_memcpy(out.ctypes.data, self.ptr, out.size * 8), "memcpy D2H")
        return out

    def free(self):
        if self.ptr:
            lib().lbfgsb200_device_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Comm:
    """NCCL communicator created from a 128-byte unique id (rank 0 makes it, everyone gets it
    through torch.distributed / any side channel)."""

    def __init__(self, unique_id, rank, nranks):
        self.h = C.c_void_p()
        _check(lib().lbfgsb200_comm_create(C.byref(self.h), unique_id, rank, nranks), "comm_create")
        self.rank, self.nranks = rank, nranks

    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(UNIQUE_ID_BYTES)
        _check(lib().lbfgsb200_comm_unique_id(buf), "comm_unique_id")
        return buf.raw

    def destroy(self):
        if self.h:
            lib().lbfgsb200_comm_destroy(self.h)
            self.h = C.c_void_p()


class Solver:
    """Resumable solver handle (lbfgsb200_create / set_x0 / iterate / get_x)."""

    def __init__(self, objective, n_global, params, comm=None, trace_rows=0, callback=None, user=None):
        """objective: a built-in name, or "callback" with callback = address of a
        lbfgsb200_fg_device_fn (a C function pointer from a user library) and user = its context."""
        self.h = C.c_void_p()
        self.params = params
        self.trace_rows = trace_rows
        if objective == "callback":
            _check(lib().lbfgsb200_create_callback_sharded(C.byref(self.h), callback, user, n_global, C.byref(params),
                                                           comm.h if comm else None, trace_rows), "create_callback")
        else:
            _check(lib().lbfgsb200_create(C.byref(self.h), OBJ[objective], n_global, C.byref(params),
                                          comm.h if comm else None, trace_rows), "create")
        self.n_local = lib().lbfgsb200_local_size(self.h)

    def device_halo(self):
        """device address of { xL, xR, dL, dR, gL, gR } (lbfgsb200_device_halo): what a sharded user objective reads"""
        return lib().lbfgsb200_device_halo(self.h)

    def set_x0(self, x0):
        """x0: numpy array (host) or int device pointer holding this rank's shard."""
        if isinstance(x0, np.ndarray):
            assert x0.dtype == np.float64 and x0.size == self.n_local
            x0 = np.ascontiguousarray(x0)
            ptr = x0.ctypes.data
        else:
            ptr = int(x0)
        _check(lib().lbfgsb200_set_x0(self.h, ptr), "set_x0")

    def save(self, path):
        _check(lib().lbfgsb200_checkpoint_save(self.h, os.fsencode(path)), "checkpoint_save")

    def load(self, path):
        _check(lib().lbfgsb200_checkpoint_load(self.h, os.fsencode(path)), "checkpoint_load")

    def iterate(self, iterations):
        return _check(lib().lbfgsb200_iterate(self.h, iterations), "iterate")

    def iterate_profiled(self, iterations):
        ms = (C.c_double * 6)()
        cnt = (C.c_int64 * 6)()
        rc = _check(lib().lbfgsb200_iterate_profiled(self.h, iterations, ms, cnt), "iterate_profiled")
        names = ("two_loop_pass", "trial", "accept", "other", "gram_rows", "combine")
        return rc, {n: dict(ms=ms[i], launches=cnt[i]) for i, n in enumerate(names)}

    def x(self, out=None):
        if out is None:
            out = np.empty(self.n_local, dtype=np.float64)
        _check(lib().lbfgsb200_get_x(self.h, out.ctypes.data), "get_x")
        return out

    def result(self):
        r = Result()
        _check(lib().lbfgsb200_get_result(self.h, C.byref(r)), "get_result")
        return r.as_dict()

    def trace(self):
        if not self.trace_rows:
            return np.zeros((0, TRACE_COLS))
        rows = np.zeros((self.trace_rows, TRACE_COLS), dtype=np.float64)
        got = lib().lbfgsb200_get_trace(self.h, rows.ctypes.data, self.trace_rows)
        _check(int(got), "get_trace")
        return rows[:got]

    def destroy(self):
        if self.h:
            lib().lbfgsb200_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def solve(objective, x0, line_search="backtracking", flavor="seq", trace_rows=0, **overrides):
    """One-shot lbfgsb200_solve() on host buffers: the call a user of the reference's
    ``LBFGS(f, grad, x0, method, max_iterations, m, tolerance)`` makes."""
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    overrides.setdefault("num_gpus", 1)  # explicit: the automatic choice depends on the box
    p = default_params(flavor, line_search=line_search, **overrides)
    x = np.empty_like(x0)
    r = Result()
    trace = np.zeros((max(trace_rows, 1), TRACE_COLS), dtype=np.float64)
    _check(lib().lbfgsb200_solve(OBJ[objective], x0.size, x0.ctypes.data, x.ctypes.data, C.byref(p),
                                 C.byref(r), trace.ctypes.data if trace_rows else None, trace_rows),
           "solve")
    info = r.as_dict()
    return x, info, trace[:min(trace_rows, info["iterations"])]
