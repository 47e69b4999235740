"""Builds librarecoal_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

Every source is compiled to an object under csrc/build/ (only when it or a header changed, sources in parallel) and the
objects are linked into rarecoal_b200/librarecoal_b200.so."""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
SO = os.path.join(HERE, "librarecoal_b200.so")
SOURCES = ["rc_pat25.cu", "rc_pat36.cu", "rc_pat12.cu", "rc_kernels.cu", "rc_api.cu", "rc_plan.cpp", "rc_frontend.cpp", "rc_multi.cpp", "rc_nccl.cpp"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ARCH + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-O2,-Wall", "--fmad=true"]
LDFLAGS = ARCH + ["-shared", "-cudart", "shared", "-ldl", "-lpthread"]


def _headers():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".h") or f.endswith(".cuh")] + \
           [os.path.join(os.path.dirname(HERE), "include", "rarecoal_b200.h")]


def _defs():
    return os.environ.get("RC_NVCC_DEFS", "").split()


def _obj(src):
    tag = hashlib.sha1(" ".join(_defs()).encode()).hexdigest()[:8] if _defs() else "std"
    return os.path.join(OBJ, f"{os.path.splitext(src)[0]}.{tag}.o")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build():
    return _stale(SO, [os.path.join(CSRC, s) for s in SOURCES] + _headers())


def _run(cmd):
    env = dict(os.environ)
    env.pop("CXX", None)
    env.pop("CC", None)
    r = subprocess.run(cmd, cwd=CSRC, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return r.returncode, r.stdout


def build_library(force=False, verbose=False, out=None):
    out = out or SO
    if not force and out == SO and not needs_build():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers()

    def compile_one(src):
        obj = _obj(src)
        if not force and not _stale(obj, [os.path.join(CSRC, src)] + hdrs):
            return 0, ""
        return _run([nvcc] + CFLAGS + _defs() + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj])

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    log = "".join(o for _, o in results)
    if any(rc for rc, _ in results):
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed compiling librarecoal_b200.so")
    rc, o = _run([nvcc] + LDFLAGS + ["-o", out] + [_obj(s) for s in SOURCES])
    if rc:
        sys.stderr.write(o)
        raise RuntimeError("nvcc failed linking librarecoal_b200.so")
    if verbose:
        print(log)
    return out


if __name__ == "__main__":
    build_library(force="-f" in sys.argv or "-v" in sys.argv, verbose="-v" in sys.argv)
    print(SO)
