"""Randomised differential test: random models (random join orders, admixture pulses, size changes, freezes) on random
problem shapes, GPU (through the C ABI) against the CPU oracle, in the three kernel configurations of the parity suite.
    python tools/fuzz_parity.py [cases] [seed]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rarecoal_b200 as R
from rarecoal_b200 import _lib as L
from rarecoal_b200 import workloads as Wk
from oracle import oracle as O


def random_model(rng, P):
    ev = [(0.0, L.RC_EV_POPSIZE, k, 0, float(rng.uniform(0.3, 3.0))) for k in range(P)]
    alive = list(range(P))
    t = 0.0
    while len(alive) > 1 and rng.random() < 0.93:
        t += float(rng.choice([0.0004, 0.001, 0.003, 0.01]))
        r = rng.random()
        if r < 0.6:                                   # join
            k, l = rng.choice(alive, 2, replace=False)
            ev.append((t, L.RC_EV_JOIN, int(k), int(l), 0.0))
            if rng.random() < 0.6:
                ev.append((t, L.RC_EV_POPSIZE, int(k), 0, float(rng.uniform(0.3, 3.0))))
            alive.remove(int(l))
        elif r < 0.8:                                 # admixture pulse
            k, l = rng.choice(alive, 2, replace=False)
            ev.append((t, L.RC_EV_SPLIT, int(k), int(l), float(rng.choice([0.0, 1.0, rng.uniform(0.02, 0.9)]))))
        elif r < 0.9:
            ev.append((t, L.RC_EV_POPSIZE, int(rng.choice(alive)), 0, float(rng.uniform(0.2, 5.0))))
        else:
            k = int(rng.choice(alive))
            ev.append((t, L.RC_EV_FREEZE, k, 0, 1.0))
            ev.append((t + 0.002, L.RC_EV_FREEZE, k, 0, 0.0))
    return ev


def run(n_cases=40, seed=1, verbose=True):
    rng = np.random.default_rng(seed)
    times = R.defaultTimes()
    engs = []
    for keyed, pm in ((1.0, 1.0), (1.0, 0.0), (0.0, 1.0)):      # pattern kernels; row / state kernels; self-contained kernels
        e = R.Engine(0)
        e.set_option("keyed", keyed)
        e.set_option("pattern_mode", pm)
        engs.append(e)
    e = R.Engine(0)                                   # a trajectory-table budget so small that calls are cut into chunks and
    e.set_option("ktab_total_mb", 1.0)                # models fall back to the self-contained kernels
    e.set_option("ktab_model_mb", 0.5)
    engs.append(e)
    e = R.Engine(0)                                   # pattern kernels on bounding boxes (opt-in), last epoch not fused
    e.set_option("pattern_embed", 1.0)
    e.set_option("fuse_last", 0.0)
    engs.append(e)
    worst = 0.0
    for case in range(n_cases):
        if os.environ.get("FUZZ_BIG"):                 # many branches, many alleles: the multi-branch shapes of the pattern kernels
            P = int(rng.integers(5, 9))
            max_af = int(rng.integers(8, 11))
        else:
            P = int(rng.integers(1, 7))
            max_af = int(rng.integers(2, 11 if P <= 4 else 8))
        nvec = [int(rng.integers(max_af, 60)) for _ in range(P)]
        pats = R.computeStandardOrder(nvec, max_af)
        if len(pats) > 600:
            pats = pats[np.sort(rng.choice(len(pats), 600, replace=False))]
        cnt = Wk.synthetic_counts(pats, total_sites=10 ** 8, seed=case)
        # a batch of models with different event structures, thetas and discovery rates in one call
        T = times[: int(rng.choice([400, 900, len(times)]))]
        specs, want = [], []
        for _ in range(int(rng.integers(1, 5))):
            ev = random_model(rng, P)
            no_sc = bool(rng.random() < 0.15)
            theta = float(rng.choice([0.0005, 0.001, 0.002]))
            disc = [float(rng.choice([1.0, rng.uniform(0.5, 1.0)])) for _ in range(P)]
            specs.append(R.ModelSpec(P, T, theta, disc, 0.0, no_sc, [R.ModelEvent.from_tuple(x) for x in ev]))
            try:
                want.append(O.logl(P, max_af, T, theta, disc, no_sc, ev, nvec, pats, cnt, 10 ** 8))
            except O.OracleError:                   # validateModel rejects it: the engine must reject it too
                want.append(None)
        for e in engs:
            e.set_problem(nvec, max_af, pats, cnt, 10 ** 8)
            logl, probs, status = e.eval_models(specs, want_probs=True)
            steps = 0
            for m, w in enumerate(want):
                if w is None:
                    assert status[m] != 0, (case, m)
                    continue
                ll_o, probs_o, w_o = w
                steps += w_o
                assert status[m] == 0, (case, m, status)
                nz = probs_o != 0
                assert np.array_equal(probs[m][~nz], probs_o[~nz]), (case, m)
                err = float(np.max(np.abs(probs[m][nz] - probs_o[nz]) / np.abs(probs_o[nz]))) if nz.any() else 0.0
                worst = max(worst, err)
                assert err < 1e-10, (case, m, P, max_af, err)
                if np.isfinite(ll_o):
                    assert abs(logl[m] - ll_o) <= 1e-9 * abs(ll_o), (case, m, logl[m], ll_o)
                else:                               # a pattern with count > 0 is unreachable under the model: log 0
                    assert logl[m] == ll_o or (np.isnan(logl[m]) and np.isnan(ll_o)), (case, m, logl[m], ll_o)
            assert e.stats()["state_steps"] == steps, (case, e.stats()["state_steps"], steps)
        ev, no_sc = specs, any(w is None for w in want)
        if verbose:
            print(f"case {case:3d}: P={P} maxAf={max_af} patterns={len(pats):4d} models={len(ev)} rejected={int(no_sc)} steps={len(T)} ok", flush=True)
    for e in engs:
        e.close()
    return worst


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    w = run(n, int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    print("all", n, "cases agree; worst relative error of a pattern probability", w)
