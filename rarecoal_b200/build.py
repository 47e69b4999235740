"""Builds librarecoal_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "librarecoal_b200.so")
SOURCES = ["rc_api.cu", "rc_kernels.cu", "rc_plan.cpp", "rc_frontend.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC,-O2,-Wall", "--fmad=true", "-cudart", "shared"]


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(os.path.dirname(HERE), "include", "rarecoal_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    defs = os.environ.get("RC_NVCC_DEFS", "").split()
    cmd = [nvcc] + NVCC_FLAGS + defs + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO] + [os.path.join(CSRC, s) for s in SOURCES]
    env = dict(os.environ)
    env.pop("CXX", None); env.pop("CC", None)
    r = subprocess.run(cmd, cwd=CSRC, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("nvcc failed building librarecoal_b200.so")
    if verbose:
        print(r.stdout)
    return SO


if __name__ == "__main__":
    build_library(force=True, verbose="-v" in sys.argv)
    print(SO)
