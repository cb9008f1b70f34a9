"""Build recipe for liblbfgsb200.so (sm_100a only; nvcc cross-compiles without a GPU).

The library is built IN-TREE (cuda-lbfgs_b200/lib/) so the .so travels with the repo
snapshot to the GPU box.  -fmad=false / -ffp-contract=off: element-wise math and the
line-search scalar logic keep the reference's operation order and round like its x86-64
build ("Rounding contract").
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "liblbfgsb200.so")
SOURCES = ["solver.cu", "comm.cpp", "x0gen.cpp"]
HEADERS = ["state.h", "ls_logic.h", "kernels.cuh", "scalar_ops.cuh", "comm.h", "compact.cuh", "accept_gram.cuh", "comm.cpp", "x0gen.cpp",
           os.path.join("..", "..", "include", "lbfgsb200.h")]


def nvcc_path():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    ccbin = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []  # dynamic libstdc++
    cmd = [nvcc_path()] + ccbin + ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off,-O2,-pthread", "-shared",
           "-Xptxas", "-v" if verbose else "-O3"]
    cmd += os.environ.get("LBFGSB200_BUILD_DEFS", "").split()  # e.g. "-DLB_AG_GROUPS=1" (kernel tuning experiments)
    cmd += [os.path.join(CSRC, f) for f in SOURCES]
    cmd += ["-ldl", "-lpthread", "-o", LIB]  # NCCL is dlopen()ed at run time (csrc/comm.cpp)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    if verbose:
        sys.stderr.write(r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
