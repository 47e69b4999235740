"""One warm evaluation of a workload for ncu launch lists:  python tools/one_eval.py c2 64 [graphs]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rarecoal_b200 as R
from rarecoal_b200 import workloads as Wk
key, B = sys.argv[1], int(sys.argv[2])
wl = Wk.get(key)
nvec, pats, cnt = wl.problem()
eng = R.Engine(0)
eng.set_option("graphs", 1.0 if len(sys.argv) > 3 else 0.0)
eng.set_problem(nvec, wl.max_af, pats, cnt, wl.total_sites)
eng.set_times(R.defaultTimes())
for i, evs in enumerate(wl.proposal_batches(3, B, seed=5)):
    flat = [e for m in evs for e in m]
    off = np.cumsum([0] + [len(m) for m in evs]).astype(np.int32)
    logl, _, st = eng.eval_raw(R.api._events_array(flat), off, np.full(B, wl.theta), np.ones(B * wl.P), False, 1.0)
    s = eng.stats()
    print(i, "kernel_ms", s["kernel_ms"], "launches", s["launches"], "logl0", logl[0], flush=True)
eng.close()
