"""One rank's share of the pattern-sharded multi-GPU run on one GPU:  python tools/shard_probe.py WORLD MODELS
(rank 0 of WORLD evaluates its pattern shard for MODELS models; rc_eval_partial, no all-reduce)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rarecoal_b200 as R
from rarecoal_b200 import workloads as Wk

world, B = int(sys.argv[1]), int(sys.argv[2])
nvec, pats, cnt = Wk.c3_problem()
T = R.defaultTimes()
specs = [R.ModelSpec(8, T, 0.0005, [1.0] * 8, 0.0, False, [R.ModelEvent.from_tuple(e) for e in ev]) for ev in Wk.c4_batch(B)]
eng = R.Engine(0)
eng.set_problem(nvec, 10, pats, cnt, 10 ** 9)
eng.set_shard(0, world)
if os.environ.get("RC_PROBE_ONCE"):            # three evaluations and out: for an ncu launch list
    for _ in range(3):
        eng.eval_models(specs)
    print(eng.stats()["kernel_ms"])
    sys.exit(0)
k = []
for it in range(6):
    eng.eval_models(specs)
    st = eng.stats()
    if it >= 2:
        k.append((st["kernel_ms"], st["traj_ms"], st["total_ms"]))
print("world", world, "B", B, "kernel_ms %.3f traj_ms %.3f total_ms %.3f" % tuple(np.median(np.array(k), 0)))
