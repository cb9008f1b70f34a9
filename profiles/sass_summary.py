#!/usr/bin/env python
"""Per-kernel SASS mnemonic histogram of liblbfgsb200.so (cuobjdump -sass): the static evidence that the hot kernels
use TMA (UTMALDG = cp.async.bulk.tensor), mbarriers (SYNCS), 128-bit shared/global accesses and FP64 math, and that
nothing spills.  (LD.E / ST.E are generic-address accesses: the vector pointers come out of DevState, so the compiler
cannot prove them global; same width, same coalescing.)  Runs on the CPU build box:  python profiles/sass_summary.py > profiles/sass_r02.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cuda-lbfgs_b200", "lib", "liblbfgsb200.so")
WANT = ["UTMALDG", "UTMAPF", "UBLKCP", "SYNCS", "LDGSTS", "LDG.E.128", "LD.E.128", "LDG.E.64", "LD.E.64", "STG.E.128", "ST.E.128", "STG.E.64", "ST.E.64", "LDS.128", "LDS.64",
        "STS.128", "STS.64", "DFMA", "DADD", "DMUL", "MUFU.RCP64H", "SHFL", "BAR.SYNC", "LDL", "STL", "S2R", "S2UR"]
KEEP = re.compile(r"k_accept_gram|k_combine_trial|k_gram_tma2d|k_trialI|k_scalar|k_two_loop_pass|k_combineE|k_acceptI|k_gramI")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for w in WANT:
                if op == w or op.startswith(w + ".") or (w.count(".") and op.startswith(w)):
                    kernels[cur][w] += 1
    demangle = subprocess.run(["cu++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("# cuobjdump -sass cuda-lbfgs_b200/lib/liblbfgsb200.so  (sm_100a; -fmad=false: DFMA only where the source says fma())")
    for (name, c), pretty in zip(kernels.items(), demangle):
        if not KEEP.search(name):
            continue
        pretty = re.sub(r"\(.*", "", pretty).replace("void lb::", "")
        cells = " ".join("%s=%d" % (w, c[w]) for w in WANT if c[w])
        print("%-58s instr=%-6d %s" % (pretty[:58], c["_total"], cells))


if __name__ == "__main__":
    main()
