import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rarecoal_b200 as R
from rarecoal_b200 import workloads as Wk
nvec, pats, cnt = Wk.c3_problem()
T = R.defaultTimes()
ev = Wk.c4_batch(1)[0]
for _ in range(3):
    t=time.time(); print(R.api.planStats(nvec, 10, pats, T, [R.ModelEvent.from_tuple(e) for e in ev])); print("wall", time.time()-t)
