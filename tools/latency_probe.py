"""Times rc_eval (host buffers in, host logl out) on the small configurations, keyed vs self-contained kernels."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rarecoal_b200 as R
from rarecoal_b200 import workloads as Wk

f = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "fourpop.json")))
T = R.defaultTimes()

def spec(P, ev):
    return R.ModelSpec(P, T, 0.0005, [1.0] * P, 0.0, False, [R.ModelEvent.from_tuple(e) for e in ev])

cases = []
for maxaf in (4, 10):
    std = R.computeStandardOrder(f["nVec"], maxaf)
    cases.append((f"fourPop m={maxaf}", f["nVec"], maxaf, std, spec(4, f["events"])))
nv = [100] * 10
cases.append(("C5 10-pop m=4", nv, 4, R.computeStandardOrder(nv, 4), spec(10, Wk.c5_events())))
nvec, pats, _ = Wk.c3_problem(max_af=6)
cases.append(("8-pop tree m=6", nvec, 6, pats, spec(8, Wk.c3_events())))
for name, nvec, maxaf, pats, sp in cases:
    for keyed in (1, 0):
        eng = R.Engine(0)
        eng.set_option("keyed", keyed)
        eng.set_problem(nvec, maxaf, pats, np.ones(len(pats), dtype=np.int64), 10 ** 9)
        for B in (1, 32):
            specs = [sp] * B
            eng.eval_models(specs)
            t0 = time.perf_counter()
            n = 20
            for _ in range(n):
                eng.eval_models(specs)
            dt = (time.perf_counter() - t0) / n
            st = eng.stats()
            print(f"{name:16s} patterns {len(pats):6d} keyed={keyed} B={B:3d}: {dt*1e3:8.3f} ms/call  launches {st['launches']:3d} kernel {st['kernel_ms']:.3f} ms")
        eng.close()
