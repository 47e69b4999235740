"""North-star config 5 end to end: `mcmc` on the 10-population admixture template (3 pulses, growth on two branches, maxAf 4)
with CHAINS lock-step chains for about PROPOSALS proposals in all; the histogram is the model's own expected counts.
    python tools/mcmc_c5.py [PROPOSALS=100000] [CHAINS=128] [NRGPUS=1]
Prints one JSON line (proposals, seconds, proposals/s)."""
import json, os, subprocess, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rarecoal_b200 as R
from rarecoal_b200 import workloads as Wk

proposals = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
chains = int(sys.argv[2]) if len(sys.argv) > 2 else 128
ngpu = int(sys.argv[3]) if len(sys.argv) > 3 else 1
wl = Wk.get("c5")
nvec, pats, _ = wl.problem()
T = R.defaultTimes()
eng = R.Engine(0)
eng.set_problem(nvec, wl.max_af, pats, None, Wk.TOTAL_SITES)
spec = R.ModelSpec(wl.P, T, wl.theta, [1.0] * wl.P, 0.0, False, [R.ModelEvent.from_tuple(e) for e in wl.events()])
_, probs, _ = eng.eval_models([spec], want_probs=True)
counts = np.rint(probs[0] * Wk.TOTAL_SITES).astype(np.int64)
eng.close()
d = tempfile.mkdtemp()
hist, tmpl, prefix = os.path.join(d, "c5.hist.txt"), os.path.join(d, "c5.template.txt"), os.path.join(d, "c5")
with open(hist, "w") as h:
    h.write("NAMES=" + ",".join(f"P{i}" for i in range(wl.P)) + "\n")
    h.write("N=" + ",".join(str(int(v)) for v in nvec) + "\n")
    h.write(f"MAX_M={wl.max_af}\nTOTAL_SITES={Wk.TOTAL_SITES}\n")
    for p, c in zip(pats, counts):
        if c > 0:
            h.write(",".join(str(int(v)) for v in p) + f" {int(c)}\n")
open(tmpl, "w").write(wl.template)
xs = []
for k, v in zip(wl.names, wl.x0):
    xs += ["-X", f"{k}={float(v)!r}"]
cycles = max(1, proposals // (chains * len(wl.names)))
cmd = [sys.executable, "-m", "rarecoal_b200", "mcmc", "-T", tmpl, "-i", hist, "-m", str(wl.max_af), "--cycles", str(cycles),
       "--nrChains", str(chains), "--seed", "7", "-o", prefix] + (["--nrGpus", str(ngpu)] if ngpu > 1 else []) + xs
t0 = time.time()
r = subprocess.run(cmd, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
dt = time.time() - t0
if r.returncode:
    print(r.stdout[-2000:])
    sys.exit(r.returncode)
n = cycles * chains * len(wl.names)
print(json.dumps({"workload": wl.label, "command": "mcmc", "chains": chains, "cycles": cycles, "parameters": len(wl.names), "proposals": n,
                  "seconds_whole_command": round(dt, 2), "proposals_per_s": round(n / dt, 1), "gpus": ngpu,
                  "note": "wall time of the whole command: start-up, histogram, plans, burn-in bookkeeping and output files included"}))
