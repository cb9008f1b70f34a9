import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import CudaRef, Ref, Oracle
import importlib.util
spec = importlib.util.spec_from_file_location("pkg", os.path.join(ROOT, "cuda-lbfgs_b200", "__init__.py"))
pkg = importlib.util.module_from_spec(spec); sys.modules["pkg"] = pkg; spec.loader.exec_module(pkg)
host = Ref("seq"); orc = Oracle()
cases = [("wolfe", None, "wolfe"), ("backtracking", None, "backtracking"), ("interpolation", None, "interpolation"),
         ("btwolfe", None, "backtracking_wolfe"), ("host", "wolfe", "wolfe"), ("host", "backtracking", "backtracking")]
n = 10000
x0 = host.x0(n, -2, 2)
for variant, lsarg, ls in cases:
    ref = CudaRef(variant)
    for K in (1, 2, 3, 5, 10, 20):
        xr, info = ref.lbfgs("rosenbrock", x0, lsarg or "wolfe", 10, K, 0.0)
        xm, im, _ = pkg.solve("rosenbrock", x0, ls, "par", profile="cuda", m=10, max_iterations=K, tolerance=0.0)
        xo, io, _ = orc.lbfgs("rosenbrock", x0, ls, "par", 10, K, 0.0, profile="cuda")
        print(variant, lsarg, "K", K, "ref-vs-mine %.3e" % np.max(np.abs(xr - xm)), "ref-vs-oracle %.3e" % np.max(np.abs(xr - xo)),
              "mine-vs-oracle %.3e" % np.max(np.abs(xm - xo)), "ref evals", info["f_evals"], info["g_evals"], "mine", im.get("f_evals"), im.get("g_evals"),
              "alphas", info["alphas"][-3:], flush=True)
