"""Single-model latency and small-batch cost through the host-buffer call, with and without launch graphs.
    python tools/latency_probe.py [c3|c2|c5|c1 ...]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rarecoal_b200 as R
from rarecoal_b200 import workloads as Wk

for key in (sys.argv[1:] or ["c3", "c2", "c5", "c1"]):
    wl = Wk.get(key)
    nvec, pats, cnt = wl.problem()
    for graphs in (0, 1):
        eng = R.Engine(0)
        eng.set_option("graphs", graphs)
        eng.set_problem(nvec, wl.max_af, pats, cnt, wl.total_sites)
        eng.set_times(R.defaultTimes())
        for B in (1, 4, 16, 64):
            batches = wl.proposal_batches(30, B, seed=5)
            packed = []
            for evs in batches:
                flat = [e for m in evs for e in m]
                off = np.cumsum([0] + [len(m) for m in evs]).astype(np.int32)
                packed.append((R.api._events_array(flat), off, np.full(B, wl.theta), np.ones(B * wl.P)))
            for i in range(5):
                eng.eval_raw(*packed[i], False, 1.0)
            t0 = time.perf_counter()
            for i in range(5, 30):
                eng.eval_raw(*packed[i], False, 1.0)
            ms = (time.perf_counter() - t0) * 1e3 / 25
            st = eng.stats()
            print(f"{key} graphs={graphs} B={B:3d}: {ms:7.3f} ms/call  host {st['host_ms']:.3f} gpu {st['total_ms']:.3f} kernel {st['kernel_ms']:.3f} "
                  f"traj {st['traj_ms']:.3f} launches {st['launches']} replay {st['graph_replay']}", flush=True)
        eng.close()
