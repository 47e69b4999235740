"""Builds a kernel variant next to the product library for A/B runs:  python tools/build_variant.py NAME -DMACRO=1 ...
-> rarecoal_b200/_v_NAME.so (git-ignored; select it with RARECOAL_B200_LIB)."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rarecoal_b200 import build

name, defs = sys.argv[1], sys.argv[2:]
out = os.path.join(build.HERE, f"_v_{name}.so")
env = dict(os.environ); env.pop("CXX", None); env.pop("CC", None)
cmd = [os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")] + build.NVCC_FLAGS + defs + ["-Xptxas", "-v", "-o", out] + \
      [os.path.join(build.CSRC, s) for s in build.SOURCES]
r = subprocess.run(cmd, cwd=build.CSRC, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
lines = r.stdout.splitlines()
for i, l in enumerate(lines):
    if "rc_propagate_keyed_kernelILi8" in l or "rc_propagate_rows" in l:
        print("\n".join(x.strip() for x in lines[i:i + 4] if "Compiling" not in x))
if r.returncode:
    print(r.stdout); sys.exit(1)
print(out)
