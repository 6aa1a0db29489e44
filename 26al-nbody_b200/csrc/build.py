"""Build libal26b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python 26al-nbody_b200/csrc/build.py [--force] [--verbose]

Every unit is compiled with --fmad=false (see UNITS below).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
# AL26_BUILD_TAG=timing (with AL26_NVCC_EXTRA="-DAL26_FUSE_TIMING") builds a second, instrumented library next to the
# product one: libal26b200_timing.so, objects under build_timing/; select it at run time with AL26_LIB=<path>
TAG = os.environ.get("AL26_BUILD_TAG", "")
SO = os.path.join(HERE, "libal26b200" + ("_" + TAG if TAG else "") + ".so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++"] + ARCH
COMMON += os.environ.get("AL26_NVCC_EXTRA", "").split()  # e.g. -DAL26_FUSE_TIMING (diagnostic builds)
# --fmad=false everywhere: the corrector / ladder / enrichment arithmetic has to follow the oracle's and the
# reference's evaluation order bit for bit; the force kernel writes every FMA explicitly (fma()), so the
# flag does not change its SASS.
UNITS = [
    ("hermite_force.cu", ["--fmad=false"]),
    ("hermite_step.cu", ["--fmad=false"]),
    ("hermite_loop.cu", ["--fmad=false"]),
    ("hermite_engine.cu", ["--fmad=false"]),
    ("hermite_chip.cu", ["--fmad=false"]),
    ("enrich.cu", ["--fmad=false"]),
    ("analysis.cu", ["--fmad=false"]),
    ("api.cu", ["--fmad=false"]),
]
DEPS = ["al26_internal.cuh", "hermite_force.cuh", "hermite_step.cuh", os.path.join("..", "..", "include", "al26_b200.h")]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    deps = [os.path.join(HERE, d) for d in DEPS] + [os.path.abspath(__file__)]
    for src, extra in UNITS:
        s = os.path.join(HERE, src)
        o = os.path.join(HERE, "build" + ("_" + TAG if TAG else ""), src.replace(".cu", ".o"))
        os.makedirs(os.path.dirname(o), exist_ok=True)
        if force or _newer(o, [s] + deps):
            cmd = [nvcc] + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.check_call(cmd)
        objs.append(o)
    if force or _newer(SO, objs):
        cmd = [nvcc, "-shared", "-o", SO] + objs + ARCH + ["-ccbin", "/usr/bin/g++", "-ldl"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
