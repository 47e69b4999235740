"""Batched callers of the likelihood path: `maxl` and `mcmc`.

The reference evaluates one likelihood at a time (GSL NMSimplex2 in Maxl.hs:76-84, a single Metropolis chain in
Mcmc.hs:115-175).  Here every optimiser / sampler step hands a BATCH of parameter vectors to one rc_eval launch:
  * maxl: Nelder-Mead on the reference's start simplex (x0 and x0 + max(1e-4, |0.01 x_j|) e_j, Maxl.hs:84), size
    tolerance 1e-8, restarts (Maxl.hs:86-97).  The p worst vertices are reflected through the centroid of the others at
    once and reflection / expansion / outside / inside contraction of each are evaluated speculatively in one launch
    (p = 1 is the classical method with speculative evaluation); a shrink evaluates all moved vertices in one launch.
  * mcmc: the reference's single-site Metropolis sampler with adaptive widths and burn-in detection, run as `nrChains`
    independent chains whose proposals of one step form the batch.
Trajectories are not bit-comparable with GSL / System.Random (parity unpinned there); objective values
are the same function (minFunc, MaxUtils.hs:63-76) evaluated on the GPU.
Output files keep the reference's names and layouts (Maxl.hs:99-116, Mcmc.hs:211-244, MaxUtils.hs:125-195).
"""
import math
import sys
from decimal import Decimal

import numpy as np

from . import frontend as F
from . import _lib as L
from .api import RarecoalException, _events_array

PENALTY = 1.0e20          # MaxUtils.hs:50-51


def hs_show(x):
    """Haskell's `show` for Double (scientific outside [0.1, 1e7), at least one fractional digit)."""
    x = float(x)
    if x != x:
        return "NaN"
    if math.isinf(x):
        return "Infinity" if x > 0 else "-Infinity"
    if x == 0.0:
        return "-0.0" if math.copysign(1.0, x) < 0 else "0.0"
    sign = "-" if x < 0 else ""
    _, digits, exp = Decimal(repr(abs(x))).as_tuple()
    ds = "".join(map(str, digits)).rstrip("0") or "0"
    e = len(digits) + exp                      # |x| = 0.d1d2... * 10^e
    if 0.1 <= abs(x) < 1.0e7:
        if e <= 0:
            return sign + "0." + "0" * (-e) + ds
        ip = ds[:e].ljust(e, "0")
        return sign + ip + "." + (ds[e:] or "0")
    return sign + ds[0] + "." + (ds[1:] or "0") + "e" + str(e - 1)


class Objective:
    """minFunc (MaxUtils.hs:63-76) for batches: -logl + siteRed * regPenalty, any Left / NaN / inf -> 1e20."""

    def __init__(self, opts, mt, hp, eng):
        self.opts, self.mt, self.hp, self.eng = opts, mt, hp, eng
        self.names = F.getParamNames(mt)
        self.P = len(mt.mtBranchNames)
        self.times = F.getTimeSteps(opts.optN0, opts.optLinGen, opts.optTMax)
        eng.set_times(self.times)
        self.n_evals = 0
        self.n_launches = 0

    def __call__(self, X):
        X = np.atleast_2d(np.asarray(X, dtype=float))
        lib = L.load()
        evs, off, ok, pen = [], [0], [], []
        for x in X:
            try:
                ev = F.instantiateEvents(self.mt, list(zip(self.names, x.tolist())))
                arr = _events_array(ev)
                pen.append(lib.rc_reg_penalty(self.P, self.opts.optRegPenalty, arr, len(ev)))
                evs.extend(ev)
                ok.append(True)
            except RarecoalException:
                pen.append(0.0)
                ok.append(False)
            off.append(len(evs))
        B = len(X)
        scores = np.full(B, PENALTY)
        if any(ok):
            logl, _, status = self.eng.eval_raw(_events_array(evs), np.array(off, dtype=np.int32),
                                                np.full(B, self.opts.optTheta), np.ones(B * self.P),
                                                self.opts.optNoShortcut, self.hp.siteRed)
            self.n_launches += 1
            for b in range(B):
                if ok[b] and status[b] == 0 and np.isfinite(logl[b]) and np.isfinite(pen[b]):
                    scores[b] = -logl[b] + self.hp.siteRed * pen[b]
        self.n_evals += B
        return scores


# ---- maxl ----------------------------------------------------------------------------------------------------------
def _simplex_size(S):
    """GSL nmsimplex2 'size': root mean square distance of the vertices from their centre."""
    c = S.mean(axis=0)
    return float(np.sqrt(np.mean(np.sum((S - c) ** 2, axis=1))))


def minimize_batched(f, x0, max_cycles, par=1, tol=1.0e-8):
    """Nelder-Mead with batched speculative evaluation; returns (x, trace rows [iter, f, size, x...])."""
    n = len(x0)
    steps = np.maximum(1.0e-4, np.abs(0.01 * x0))                     # Maxl.hs:84
    S = np.vstack([x0] + [x0 + steps[j] * np.eye(n)[j] for j in range(n)])
    fv = f(S)                                                         # n+1 models, one launch
    par = max(1, min(par, n))
    trace = []
    for it in range(1, max_cycles + 1):
        order = np.argsort(fv, kind="stable")
        S, fv = S[order], fv[order]
        size = _simplex_size(S)
        trace.append([float(it), float(fv[0]), size] + S[0].tolist())
        if size < tol:
            break
        keep = n + 1 - par
        c = S[:keep].mean(axis=0)
        worst = S[keep:]
        cand = np.vstack([c + (c - worst), c + 2.0 * (c - worst), c + 0.5 * (c - worst), c - 0.5 * (c - worst)])
        fc = f(cand).reshape(4, par)                                  # reflection, expansion, outside, inside
        shrink = False
        for i in range(par):
            fr, fe, foc, fic = fc[:, i]
            xr, xe, xoc, xic = (cand[q * par + i] for q in range(4))
            f_best, f_second, f_worst = fv[0], fv[keep - 1], fv[keep + i]
            if fr < f_best:
                S[keep + i], fv[keep + i] = (xe, fe) if fe < fr else (xr, fr)
            elif fr < f_second:
                S[keep + i], fv[keep + i] = xr, fr
            elif fr < f_worst and foc <= fr:
                S[keep + i], fv[keep + i] = xoc, foc
            elif fr >= f_worst and fic < f_worst:
                S[keep + i], fv[keep + i] = xic, fic
            else:
                shrink = True
        if shrink:
            best = S[np.argmin(fv)].copy()
            idx = [j for j in range(n + 1) if not np.array_equal(S[j], best)]
            S[idx] = best + 0.5 * (S[idx] - best)
            fv[idx] = f(S[idx])                                       # all moved vertices, one launch
    j = int(np.argmin(fv))
    return S[j].copy(), trace


def run_maxl(a, opts, mt, params, hp, eng):
    """Maxl.hs:40-74."""
    params = F.fillParameterDictWithDefaults(mt, params)
    F.instantiateModel(opts, mt, params)                              # fail early on a bad template / constraint
    f = Objective(opts, mt, hp, eng)
    x = np.array([v for _, v in params])
    if f(x)[0] >= PENALTY:
        raise RarecoalException("the initial point has no finite likelihood")
    trace = []
    for _ in range(max(1, a.nrRestarts)):                             # minimizeWithRestarts, Maxl.hs:86-97
        print("minimizing from point " + str(x.tolist()), file=sys.stderr)
        x, tr = minimize_batched(f, x, a.maxCycles, a.parallel)
        trace += tr
    score = float(f(x)[0])
    names = f.names
    with open(a.prefix + ".paramEstimates.txt", "w") as h:            # reportMaxResult, Maxl.hs:99-102
        h.write("Score\t" + hs_show(score) + "\n")
        for p, v in zip(names, x):
            h.write(p + "\t" + hs_show(v) + "\n")
    with open(a.prefix + ".trace.txt", "w") as h:                     # reportTraceSimplex, Maxl.hs:104-108
        h.write("\t".join(["Nr", "-Log-Likelihood", "Simplex size"] + names) + "\n")
        for row in trace:
            h.write("\t".join(hs_show(v) for v in row) + "\n")
    _write_fit_tables(a.prefix, opts, mt, list(zip(names, x.tolist())), hp, eng)
    print(f"maxl: score {score!r} after {f.n_evals} likelihoods in {f.n_launches} launches", file=sys.stderr)
    return x, score


# ---- mcmc ----------------------------------------------------------------------------------------------------------
def _burnin(vals):
    """computeBurninCycles, Mcmc.hs:110-113."""
    bound = min(vals) + 10.0
    return next(i for i, v in enumerate(vals) if v < bound)


def run_mcmc(a, opts, mt, params, hp, eng):
    """Mcmc.hs:60-97 with `nrChains` chains advanced in lock-step (one proposal of every chain per launch)."""
    params = F.fillParameterDictWithDefaults(mt, params)
    F.instantiateModel(opts, mt, params)
    f = Objective(opts, mt, hp, eng)
    names = f.names
    x0 = np.array([v for _, v in params])
    v0 = float(f(x0)[0])
    if v0 >= PENALTY:
        raise RarecoalException("the initial point has no finite likelihood")
    C, k = max(1, a.nrChains), len(x0)
    rng = [np.random.default_rng([a.seed, c]) if a.seed else np.random.default_rng() for c in range(C)]
    point = np.tile(x0, (C, 1))
    value = np.full(C, v0)
    width = np.tile(np.maximum(1.0e-8, np.abs(x0 / 100.0)), (C, 1))   # Mcmc.hs:73
    success = np.full((C, k), 0.44)
    states = [[] for _ in range(C)]          # per chain: (value, point, widths, success) after every cycle
    scores = [[] for _ in range(C)]
    cycles = 0

    def not_done():
        if cycles < a.cycles:
            return True
        return any(cycles - _burnin(scores[c]) <= a.cycles for c in range(C))     # mcmcNotDone, Mcmc.hs:100-108

    while not_done():
        orders = [rng[c].permutation(k) for c in range(C)]                         # shuffled sweep, Mcmc.hs:123
        for pos in range(k):
            idx = np.array([orders[c][pos] for c in range(C)])
            prop = point.copy()
            for c in range(C):                                                     # propose, Mcmc.hs:155-164
                d = width[c, idx[c]]
                prop[c, idx[c]] += rng[c].uniform(-d, d)
            newv = f(prop)                                                         # C likelihoods, one launch
            for c in range(C):                                                     # isSuccessFul / accept / reject
                i = idx[c]
                ok = newv[c] < value[c] or rng[c].uniform(0.0, 1.0) < math.exp(min(0.0, value[c] - newv[c]))
                if ok:
                    point[c], value[c] = prop[c], newv[c]
                    success[c, i] = success[c, i] * 0.975 + 0.025
                else:
                    success[c, i] = success[c, i] * 0.975
        cycles += 1
        for c in range(C):
            scores[c].append(float(value[c]))
            states[c].append((float(value[c]), point[c].copy(), width[c].copy(), success[c].copy()))
        if cycles % 10 == 0:                                                       # Mcmc.hs:128-131,192-200
            print(f"Cycle={cycles}\tValue={hs_show(value[0])}\t" +
                  "\t".join(f"{n}={hs_show(v)}" for n, v in zip(names, point[0])), file=sys.stderr)
            low, high = success < 0.29, success > 0.59
            width[low] /= 1.5
            width[high] *= 1.5
    # ---- reportPosteriorStats (Mcmc.hs:211-233), chains pooled after their own burn-in
    burn = [_burnin(scores[c]) for c in range(C)]
    pooled = [s for c in range(C) for s in states[c][burn[c]:]]
    pv = [s[0] for s in pooled]
    mi = int(np.argmin(pv))

    def order_stats(vals):
        n = float(len(vals))
        sv = sorted(vals)
        return [vals[mi]] + [sv[int(math.floor(q * n))] for q in (0.025, 0.5, 0.975)]

    with open(a.prefix + ".paramEstimates.txt", "w") as h:
        h.write(f"# Nr Burnin Cycles: {burn[0] if C == 1 else sum(burn)}\n")
        h.write(f"# Nr Main Cycles: {len(pooled)}\n")
        h.write("Param\tMaxL\tLowerCI\tMedian\tUpperCI\n")
        h.write("\t".join(["Score"] + [hs_show(v) for v in order_stats(pv)]) + "\n")
        for i, nme in enumerate(names):
            h.write("\t".join([nme] + [hs_show(v) for v in order_stats([s[1][i] for s in pooled])]) + "\n")
    with open(a.prefix + ".trace.txt", "w") as h:                                  # reportTrace, Mcmc.hs:235-244
        h.write("\t".join(["Score"] + names + [n + "_delta" for n in names] + [n + "_success" for n in names]) + "\n")
        for c in range(C):
            for s in states[c]:
                h.write("\t".join(hs_show(v) for v in [s[0]] + s[1].tolist() + s[2].tolist() + s[3].tolist()) + "\n")
    best = pooled[mi][1]
    _write_fit_tables(a.prefix, opts, mt, list(zip(names, best.tolist())), hp, eng)
    print(f"mcmc: {cycles} cycles x {C} chains, {f.n_evals} likelihoods in {f.n_launches} launches", file=sys.stderr)
    return best, float(pv[mi])


# ---- fit tables (MaxUtils.hs:125-195; patterns and names in model-branch order, no jackknife columns) ----------------
def _write_fit_tables(prefix, opts, mt, params, hp, eng):
    spec = F.instantiateModel(opts, mt, params)
    spectrum = eng.computeFrequencySpectrum(spec)
    with open(prefix + ".frequencyFitTable.txt", "w") as h:
        h.write("Pattern\tCount\tFreq\tTheoryFreq\trelDev%\n")
        for e in spectrum:
            dev = "n/a" if e.spCount == 0 else str(int(round(100.0 * (e.spTheory - e.spFreq) / e.spFreq)))
            h.write("\t".join(["(" + ",".join(map(str, e.spPattern)) + ")", str(e.spCount), hs_show(e.spFreq),
                               hs_show(e.spTheory), dev]) + "\n")
    nvec = [int(v) for v in hp.nVec]
    names = list(mt.mtBranchNames)
    n = len(nvec)

    def summaries(key):
        single = [0.0] * n
        sharing = [0.0] * (n * (n + 1) // 2)
        for e in spectrum:
            p, val = e.spPattern, key(e)
            if sum(p) == 1:
                single[next(i for i, v in enumerate(p) if v > 0)] += val
            q = 0
            for i in range(n):
                for j in range(i, n):
                    den = nvec[i] * (nvec[i] - 1) if i == j else nvec[i] * nvec[j]
                    num = p[i] * (p[i] - 1) if i == j else p[i] * p[j]
                    sharing[q] += val * num / den if den else 0.0
                    q += 1
        return single + sharing

    real, fit = summaries(lambda e: e.spFreq), summaries(lambda e: e.spTheory)
    labels = [nm + "(singletons)" for nm in names] + [names[i] + "/" + names[j] for i in range(n) for j in range(i, n)]
    with open(prefix + ".summaryFitTable.txt", "w") as h:
        h.write("Populations\tAlleleSharing\tPredicted\trelDev%\n")
        for lab, r, t in zip(labels, real, fit):
            dev = str(int(round(100.0 * (t - r) / r))) if r > 0 else "n/a"
            h.write("\t".join([lab, hs_show(r), hs_show(t), dev]) + "\n")


# ---- CLI -----------------------------------------------------------------------------------------------------------
def add_parsers(sub, _general, _model, _hist, _opts, _template, _params, _load):
    def wrap(fn):
        def go(a):
            mt = _template(a)
            hp, eng = _load(a, mt)
            fn(a, _opts(a), mt, _params(a), hp, eng)
        return go

    p = sub.add_parser("maxl", help="Maximize the likelihood of the model given the data set")
    _general(p); _model(p); _hist(p)
    p.add_argument("--maxCycles", type=int, default=10000)
    p.add_argument("--nrRestarts", type=int, default=5)
    p.add_argument("-o", "--prefix", required=True)
    p.add_argument("--parallel", type=int, default=4,
                   help="vertices reflected per iteration; each costs 4 speculative likelihoods in the same launch")
    p.set_defaults(fn=wrap(run_maxl))
    p = sub.add_parser("mcmc", help="Run MCMC on the model and the data")
    _general(p); _model(p); _hist(p)
    p.add_argument("--cycles", type=int, default=1000)
    p.add_argument("-o", "--prefix", required=True)
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--nrChains", type=int, default=1, help="independent chains advanced in one launch per step")
    p.set_defaults(fn=wrap(run_mcmc))
