"""Builds a kernel variant next to the product library for A/B runs:  python tools/build_variant.py NAME -DMACRO=1 ...
-> rarecoal_b200/_v_NAME.so (git-ignored; select it with RARECOAL_B200_LIB)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rarecoal_b200 import build

name, defs = sys.argv[1], sys.argv[2:]
os.environ["RC_NVCC_DEFS"] = " ".join(defs)
out = os.path.join(build.HERE, f"_v_{name}.so")
print(build.build_library(out=out))
