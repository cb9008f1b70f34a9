"""Run the same solve six times in one process (the 2nd..6th re-use the first one's arena from the pool) and report
the first iteration whose f differs bit-wise: python benchmarks/repeat_check.py [n] [m] [graph]"""
import sys, os, hashlib, numpy as np
sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import importlib.util
spec = importlib.util.spec_from_file_location("cuda_lbfgs_b200", "cuda-lbfgs_b200/__init__.py")
gpu = importlib.util.module_from_spec(spec); sys.modules["cuda_lbfgs_b200"] = gpu; spec.loader.exec_module(gpu)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
m = int(sys.argv[2]) if len(sys.argv) > 2 else 8
graph = int(sys.argv[3]) if len(sys.argv) > 3 else 1
x0 = gpu.x0_uniform(n, -2, 2)
p = gpu.default_params("par", m=m, line_search="wolfe", max_iterations=12, tolerance=0.0, use_graph=graph)
ref = None
for rep in range(6):
    s = gpu.Solver("rosenbrock", n, p, trace_rows=12)
    s.set_x0(x0)
    s.iterate(12)
    x = s.x(); tr = s.trace(); r = s.result()
    s.destroy()
    fs = [float(t[1]).hex() for t in tr]
    if ref is None:
        ref = (x.copy(), fs)
        print("rep0 flow=%s graph=%s f=%s" % (r.get("flow"), r.get("graph"), fs[-1]))
    else:
        first = next((k for k in range(len(fs)) if fs[k] != ref[1][k]), None)
        nd = int(np.count_nonzero(x != ref[0]))
        print("rep%d first differing iteration %s, differing x entries %d, max |dx| %.3e" % (rep, first, nd, float(np.max(np.abs(x - ref[0])))))
        if first is not None:
            print("   f there: %s vs %s" % (fs[first], ref[1][first]))
