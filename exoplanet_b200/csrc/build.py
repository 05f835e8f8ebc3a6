"""Build exoplanet_b200/libxo_b200.so in-tree with nvcc for sm_100a.

    python exoplanet_b200/csrc/build.py [--force] [--verbose]

No JIT cache, no torch extension machinery: the C-ABI library has no torch/pybind
dependency, so a plain `nvcc -shared` is enough and the .so travels with the repo
snapshot to the GPU box.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "libxo_b200.so")
BUILD = os.path.join(HERE, "_build")
SOURCES = ["xo_ops.cu", "xo_transit.cu", "xo_model.cu", "xo_celerite.cu", "xo_host.cu", "xo_diag.cu", "xo_joint.cu", "xo_sampler.cu", "xo_nuts.cu"]
HEADERS = ["xo_math.cuh", "xo_common.cuh", "xo_contact.cuh", "xo_transit.cuh", "xo_prologue.cuh", "xo_posterior.cuh", os.path.join("..", "..", "include", "xo_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-ccbin", "/usr/bin/g++",
]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    mt = os.path.getmtime(target)
    return any(os.path.getmtime(d) > mt for d in deps)


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    hdrs = [os.path.join(HERE, h) for h in HEADERS] + [os.path.abspath(__file__)]
    objs = []
    jobs = []
    for src in SOURCES:
        s = os.path.join(HERE, src)
        o = os.path.join(BUILD, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))

    if jobs:
        with ThreadPoolExecutor(max_workers=4) as ex:
            list(ex.map(run, jobs))
    if jobs or force or _stale(OUT, objs):
        cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++",
                                                      "-cudart", "shared", "-ldl"]
        run(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
