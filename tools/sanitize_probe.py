"""Small runs of every kernel tier for compute-sanitizer (memcheck / racecheck)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rarecoal_b200 as R
from rarecoal_b200 import workloads as Wk

T = R.defaultTimes()[:260]
f = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "fourpop.json")))

def spec(P, ev, ns=False):
    return R.ModelSpec(P, T, 0.0005, [1.0] * P, 0.0, ns, [R.ModelEvent.from_tuple(e) for e in ev])

out = []
for keyed in (1, 0):
    eng = R.Engine(0)
    eng.set_option("keyed", keyed)
    nvec, pats, cnt = Wk.c3_problem(max_af=4)                 # 8-pop tree: state mode + row mode epochs
    eng.set_problem(nvec, 4, pats, cnt, Wk.TOTAL_SITES)
    ev = [(t * 0.25, ty, k, l, v) for (t, ty, k, l, v) in Wk.c3_events()]
    out.append(eng.eval_models([spec(8, ev), spec(8, ev, True)])[0])
    std = R.computeStandardOrder(f["nVec"], 8)                 # splits: general live sets, big tier for > 256 states
    eng.set_problem(f["nVec"], 8, std, np.ones(len(std), dtype=np.int64), 10 ** 9)
    ev4 = [(t * 0.2, ty, k, l, v) for (t, ty, k, l, v) in f["events"]]
    out.append(eng.eval_models([spec(4, ev4)])[0])
    eng.close()
print(out)
assert np.allclose(out[0], out[2], rtol=1e-12) and np.allclose(out[1], out[3], rtol=1e-12)
print("sanitize probe ok")
