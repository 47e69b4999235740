"""Checksums (rc_plan_last_checksum) of the transition plans of the bench workloads: a planner change that is meant to be a
pure speed-up must leave them unchanged.    python tools/plan_checksums.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rarecoal_b200 as R
from rarecoal_b200 import workloads as Wk

T = R.defaultTimes()
nvec, pats, _ = Wk.c3_problem()
print("c3", hex(R.planStats(nvec, 10, pats, T, Wk.c3_events())["checksum"]))
for ev in Wk.c4_batch(3)[1:]:
    print("c4", hex(R.planStats(nvec, 10, pats, T, ev)["checksum"]))
wl = Wk.get("c2")
nv, ps, _ = wl.problem()
for ev in wl.proposal_batches(1, 2, seed=3)[0]:
    print("c2", hex(R.planStats(nv, wl.max_af, ps, T, ev)["checksum"]))
nv = [100] * 10
print("c5", hex(R.planStats(nv, 4, R.computeStandardOrder(nv, 4), T, Wk.c5_events())["checksum"]))
wl = Wk.get("c1")
nv, ps, _ = wl.problem()
print("c1", hex(R.planStats(nv, wl.max_af, ps, T, wl.proposal_batches(1, 1, seed=3)[0][0])["checksum"]))
