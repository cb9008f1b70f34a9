import json, sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from oracle import Oracle
o = Oracle()
d = json.load(open(os.path.join(ROOT, "tests/golden/cuda_reference_traces.json")))
LSMAP = {"wolfe": "wolfe", "backtracking": "backtracking", "interpolation": "interpolation", "btwolfe": "backtracking_wolfe"}
for name, c in d["traces"].items():
    ls = c["line_search"] or LSMAP[c["variant"]]
    x0 = o.x0(c["n"], c["lo"], c["hi"])
    for K, w in sorted(c["steps"].items(), key=lambda kv: int(kv[0])):
        x, info, _ = o.lbfgs(c["objective"], x0, ls, "par" if c["variant"] == "host" else "par_inlined", c["m"], int(K), c["tolerance"], profile="cuda")
        n = c["n"]
        xs = x[:: max(1, n // 64)][:64]
        ws = np.array([float.fromhex(v) for v in w["x_sample"]])
        print(name, K, "max|dx| %.2e" % np.max(np.abs(xs - ws)), "evals", info["f_evals"], info["g_evals"], "ref", w["f_evals"], w["g_evals"])
