"""Extract the reference's unused dense SPD fixtures into tests/golden/dense_spd.json.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/make_golden_dense.py

sequential-implementation/matrices.h holds mat{N} (row-major N x N, float literals),
linear{N} and minimum{N} for the objective f(x) = x^T A x + b^T x, whose minimiser satisfies
2 A x + b = 0.  The header is included by the reference's main.cpp /
benchmark.h but never referenced; the values (float32-rounded, as the compiler would read the
`f`-suffixed literals) become known-answer tests for user (callback) objectives.
"""
import json
import os
import re

import numpy as np

SRC = "/root/reference/sequential-implementation/matrices.h"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "dense_spd.json")
SIZES = (2, 3, 4, 5, 10, 50)


def block(text, name):
    m = re.search(r"\b%s\b[^=]*=\s*\{(.*?)\};" % re.escape(name), text, re.S)
    vals = re.findall(r"[-+]?\d*\.\d+(?:[eE][-+]?\d+)?f?", m.group(1))
    return [float(np.float32(v.rstrip("f"))) for v in vals]


def main():
    text = open(SRC).read()
    out = {"generator": "oracle/make_golden_dense.py", "source": "sequential-implementation/matrices.h (float literals)",
           "objective": "f(x) = x^T A x + b^T x ; grad = (A + A^T) x + b", "cases": {}}
    for n in SIZES:
        A = block(text, "mat%d" % n)
        b = block(text, "linear%d" % n)
        xmin = block(text, "minimum%d" % n)
        assert len(A) == n * n and len(b) == n and len(xmin) == n, (n, len(A), len(b), len(xmin))
        An = np.array(A).reshape(n, n)
        resid = np.max(np.abs(2 * An @ np.array(xmin) + np.array(b)))
        out["cases"][str(n)] = {"A": A, "b": b, "minimum": xmin, "stationarity_residual_of_fixture": float(resid)}
        print(n, "residual of the reference's own minimum:", resid)
    json.dump(out, open(OUT, "w"))
    print("wrote", os.path.normpath(OUT))


if __name__ == "__main__":
    main()
