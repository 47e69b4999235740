"""Times the C3 batch (44 models) with whatever library RARECOAL_B200_LIB names; with a library built with
-DRC_PROBE_NOLOOP the propagation kernels skip their grid steps, so the difference to the normal library is the time
of the main loops and the remainder is per-launch set-up / hand-over.  Timing only: the probe build's results are wrong."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rarecoal_b200 as R
from rarecoal_b200 import workloads as Wk

B = int(sys.argv[1]) if len(sys.argv) > 1 else 44
nvec, pats, cnt = Wk.c3_problem()
T = R.defaultTimes()
specs = [R.ModelSpec(8, T, 0.0005, [1.0] * 8, 0.0, False, [R.ModelEvent.from_tuple(e) for e in ev]) for ev in Wk.c4_batch(B)]
eng = R.Engine(0)
eng.set_problem(nvec, 10, pats, cnt, 10 ** 9)
for _ in range(3):
    eng.eval_models(specs)
k = []
for _ in range(8):
    eng.eval_models(specs)
    st = eng.stats()
    k.append((st["kernel_ms"], st["traj_ms"], st["total_ms"]))
k = np.array(k)
print("lib", os.environ.get("RARECOAL_B200_LIB", "default"), "B", B, "kernel_ms %.3f traj_ms %.3f total_ms %.3f" % tuple(np.median(k, 0)))
