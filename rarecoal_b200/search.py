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
import ctypes as C
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
        self._names_c = self._text_c = self._ev = None
        self.gpu_ms = self.host_ms = self.plan_ms = 0.0
        self.plan_misses = 0

    def __call__(self, X):
        X = np.ascontiguousarray(np.atleast_2d(np.asarray(X, dtype=float)))
        lib = L.load()
        B, k = X.shape
        # one call instantiates the template for the whole batch (parsed once) and returns the events packed for rc_eval, the
        # per-model status (a violated constraint, ModelTemplate.hs:229-241, is a Left: the 1e20 penalty) and the regularisation
        # penalties
        if self._names_c is None:
            self._names_c = (C.c_char_p * max(1, k))(*[n.encode() for n in self.names])
            self._text_c = self.mt.text.encode()
        cap = B * (2 * len(self.mt.mtEvents) + 1)
        if self._ev is None or len(self._ev) < cap:
            self._ev = (L.rc_event * max(1, cap))()
        off = np.zeros(B + 1, dtype=np.int32)
        st = np.zeros(B, dtype=np.int32)
        pen = np.zeros(B)
        err = C.create_string_buffer(256)
        n = lib.rc_template_instantiate_batch(self._text_c, k, self._names_c, B, X.ctypes.data_as(C.POINTER(C.c_double)),
                                              float(self.opts.optRegPenalty), self._ev, cap, off.ctypes.data_as(C.POINTER(C.c_int32)),
                                              st.ctypes.data_as(C.POINTER(C.c_int32)), pen.ctypes.data_as(C.POINTER(C.c_double)), err, 256)
        if n < 0:
            raise RarecoalException(err.value.decode() or "template instantiation failed")
        ok = st == 0
        scores = np.full(B, PENALTY)
        if ok.any():
            logl, _, status = self.eng.eval_raw(self._ev, off, np.full(B, self.opts.optTheta), np.ones(B * self.P),
                                                self.opts.optNoShortcut, self.hp.siteRed)
            self.n_launches += 1
            st_ = self.eng.stats()
            self.gpu_ms += st_["total_ms"]
            self.host_ms += st_["host_ms"]
            self.plan_ms += st_["plan_build_ms"]
            self.plan_misses += st_["plan_misses"]
            good = ok & (status == 0) & np.isfinite(logl) & np.isfinite(pen)
            scores[good] = -logl[good] + self.hp.siteRed * pen[good]
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
    # dimension-dependent coefficients (Gao & Han 2012): with the classical 2 / 0.5 / 0.5 the simplex collapses long before
    # the optimum in 15+ dimensions; the reference's GSL simplex has the same weakness and relies on its restarts
    k_exp, k_con, k_shr = 1.0 + 2.0 / n, 0.75 - 0.5 / n, 1.0 - 1.0 / n
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
        cand = np.vstack([c + (c - worst), c + k_exp * (c - worst), c + k_con * (c - worst), c - k_con * (c - worst)])
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
            S[idx] = best + k_shr * (S[idx] - best)
            fv[idx] = f(S[idx])                                       # all moved vertices, one launch
    j = int(np.argmin(fv))
    return S[j].copy(), trace


def polish_batched(f, x, fx, max_iter=400, rel_tol=1.0e-7):
    """Batched coordinate pattern search, run after the simplex: the likelihood is piecewise constant in every time
    parameter (events fire on the grid, Core.hs:120-135), which stalls a collapsing simplex on a ridge.  Every iteration is
    ONE launch: x +- delta_j e_j for every parameter (the central-difference batch of SURVEY.md 8d, C4) plus a few combinations
    of the moves that improved in the previous iteration; the best point is taken, deltas halve when nothing improves."""
    n = len(x)
    delta = np.maximum(1.0e-4, np.abs(0.01 * x))
    combos = []
    trace = []
    for it in range(max_iter):
        if np.all(delta <= np.maximum(1.0e-12, rel_tol * np.abs(x))):
            break
        eye = np.eye(n) * delta
        cand = np.vstack([x + eye, x - eye] + combos)
        fc = f(cand)
        gain_p, gain_m = fx - fc[:n], fx - fc[n:2 * n]
        j = int(np.argmin(fc))
        move = np.where(gain_p > np.maximum(gain_m, 0.0), delta, np.where(gain_m > 0.0, -delta, 0.0))
        if fc[j] < fx:
            x, fx = cand[j].copy(), float(fc[j])
            # next launch also tries all improving moves at once, at three scales
            combos = [x + sc * move for sc in (1.0, 0.5, 0.25)] if np.count_nonzero(move) > 1 else []
            grow = np.zeros(n, dtype=bool)
            if j < 2 * n:
                grow[j % n] = True
            delta = np.where(grow, delta * 1.5, delta)
        else:
            combos = []
            delta = delta * 0.5
        trace.append([float(it + 1), fx, float(np.max(delta / np.maximum(np.abs(x), 1e-300)))] + x.tolist())
    return x, fx, trace


def bfgs_batched(f, x, fx, max_iter=200, fd_rel=0.01):
    """Quasi-Newton descent on batches: the central-difference gradient is ONE launch of 2n models (44
    models for the 22 parameters of the 8-population template), the line search along the BFGS direction a second launch
    of a dozen step lengths.  Works in relative coordinates z = x / scale; the finite-difference step (1 % by default, the
    simplex's own start step, Maxl.hs:84) spans a few cells of the time grid, on which the likelihood is piecewise constant
    in every event time (Core.hs:120-135).  Follows the narrow diagonal valleys (size / time trade-offs) in which both the
    simplex and a coordinate search stall."""
    n = len(x)
    scale = np.maximum(np.abs(x), 1.0e-4)
    H = np.eye(n)
    alphas = np.array([4.0, 2.0, 1.0, 0.5, 0.25, 0.1, 0.05, 0.02, 0.01, 0.003, 0.001])
    trace = []
    g_prev, z_prev = None, None
    h = fd_rel
    for it in range(max_iter):
        z = x / scale
        E = np.eye(n) * h
        fc = f(np.vstack([(z + E) * scale, (z - E) * scale]))
        g = (fc[:n] - fc[n:]) / (2.0 * h)
        # the probes themselves are candidates
        jb = int(np.argmin(fc))
        if g_prev is not None:
            sk, yk = z - z_prev, g - g_prev
            sy = float(sk @ yk)
            if sy > 1.0e-12 * np.linalg.norm(sk) * np.linalg.norm(yk):
                rho = 1.0 / sy
                V = np.eye(n) - rho * np.outer(sk, yk)
                H = V @ H @ V.T + rho * np.outer(sk, sk)
        p = -H @ g
        if g @ p >= 0.0:                       # not a descent direction: restart from steepest descent
            H = np.eye(n)
            p = -g
        p = p / max(1.0, np.max(np.abs(p)) / 0.2)          # no coordinate moves by more than 20 % at unit step
        cand = (z + np.outer(alphas, p)) * scale
        fl = f(cand)
        ib = int(np.argmin(fl))
        g_prev, z_prev = g, z
        f_new, x_new = (float(fl[ib]), cand[ib]) if fl[ib] <= fc[jb] else (float(fc[jb]), np.vstack([(z + E) * scale, (z - E) * scale])[jb])
        if f_new < fx - 1.0e-9 * abs(fx) * 1.0e-3:
            x, fx = x_new.copy(), f_new
        else:
            if h <= 1.0e-4:
                break
            h *= 0.5                            # nothing improved: look closer, forget the curvature
            H = np.eye(n)
            g_prev = None
        trace.append([float(it + 1), fx, h] + x.tolist())
    return x, fx, trace


def random_polish_batched(f, x, fx, seed=1, batch=64, max_launches=300, rel_tol=1.0e-6):
    """Last stage of the batched search: `batch` random perturbations of one to three parameters per launch, the best one
    kept; the width shrinks when a launch brings nothing and grows when it does.  The event times make the objective a
    staircase (Core.hs:120-135), and moving a time to the next grid cell only pays off together with a size or rate change —
    moves that neither finite differences nor single-coordinate probes see, and that a sampler finds by chance."""
    rng = np.random.default_rng(seed)
    n = len(x)
    w = 0.01
    trace = []
    for it in range(max_launches):
        if w < rel_tol:
            break
        cand = np.tile(x, (batch, 1))
        for c in range(batch):
            js = rng.choice(n, size=int(rng.integers(1, 4)), replace=False)
            cand[c, js] *= 1.0 + rng.uniform(-w, w, size=len(js))
        fc = f(cand)
        j = int(np.argmin(fc))
        if fc[j] < fx:
            x, fx = cand[j].copy(), float(fc[j])
            w = min(0.05, w * 1.3)
        else:
            w *= 0.7
        trace.append([float(it + 1), fx, w] + x.tolist())
    return x, fx, trace


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
        if not a.noPolish:
            x, fx, tr = bfgs_batched(f, x, float(f(x)[0]))
            trace += tr
            x, fx, tr = polish_batched(f, x, fx)
            trace += tr
            x, fx, tr = random_polish_batched(f, x, fx, seed=len(trace))
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
    print(f"maxl: score {score!r} after {f.n_evals} likelihoods in {f.n_launches} launches (GPU {f.gpu_ms / 1e3:.2f} s, host side of the "
          f"calls {f.host_ms / 1e3:.2f} s of which {f.plan_misses} transition plans built in {f.plan_ms / 1e3:.2f} s)", file=sys.stderr)
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


# ---- find ----------------------------------------------------------------------------------------------------------
def find_grid(events, nrPops, l, branchAge, deltaTime, maxTime):
    """allParamPairs of Find.hs:49-58: every (branch, join time) at which the free branch `l` could join the tree.
    `events` are (t, type, k, l, v) tuples in model order."""
    joins = [(t, k, frm) for (t, ty, k, frm, _) in events if ty == L.RC_EV_JOIN]

    def is_empty(branch):                                                       # isEmptyBranch, Find.hs:79-82
        return any(frm == branch and t < branchAge for t, _, frm in joins)

    def join_times(k):                                                          # getJoinTimes, Find.hs:84-89
        times, i = [], 1.0
        while i * deltaTime + branchAge <= maxTime:                             # map ((+branchAge) . (*deltaT)) [1.0, 2.0 ..]
            times.append(i * deltaTime + branchAge)
            i += 1.0
        leave = [t for t, _, frm in joins if frm == k]
        return times if not leave else [t for t in times if t < leave[0]]

    return [(k, t) for k in range(nrPops) if k != l and not is_empty(k) for t in join_times(k)]


def run_find(a, opts, mt, params, hp, eng):
    """Find.hs:29-63: where does the free branch join the tree?  The reference evaluates the (branch, time) grid one
    likelihood after the other (Find.hs:59-61); here the whole grid is ONE batch of models for rc_eval."""
    spec = F.instantiateModel(opts, mt, params)
    if a.queryName not in mt.mtBranchNames:
        raise RarecoalException("could not find branch name " + a.queryName)    # findQueryIndex, Find.hs:73-76
    l = mt.mtBranchNames.index(a.queryName)
    ev = [e.as_tuple() for e in spec.mEvents]
    if any(ty == L.RC_EV_JOIN and l in (k, frm) for (_, ty, k, frm, _) in ev):  # hasFreeBranch, Find.hs:66-69
        from .api import RarecoalModelException
        raise RarecoalModelException(f"model must have free branch {l}")
    if a.branchAge > 0.0:                                                       # an ancient sample: frozen until its age
        ev = [(0.0, L.RC_EV_FREEZE, l, 0, 1.0), (a.branchAge, L.RC_EV_FREEZE, l, 0, 0.0)] + ev
    pairs = find_grid(ev, spec.mNrPops, l, a.branchAge, a.deltaTime, a.maxTime)
    if not pairs:
        raise RarecoalException("find: empty (branch, time) grid")
    eng.set_times(spec.mTimeSteps)
    P = spec.mNrPops
    lls = np.zeros(len(pairs))
    step = max(1, a.batch)
    launches = 0
    for i0 in range(0, len(pairs), step):
        chunk = pairs[i0:i0 + step]
        evs, off = [], [0]
        for k, t in chunk:
            evs += [(t, L.RC_EV_JOIN, k, l, 0.0)] + ev                          # newE : e, Find.hs:94-96
            off.append(len(evs))
        logl, _, status = eng.eval_raw(_events_array(evs), np.array(off, dtype=np.int32), np.full(len(chunk), spec.mTheta),
                                       np.ones(len(chunk) * P), spec.mNoShortcut, hp.siteRed)
        launches += 1
        if status.any():                                                        # tryEither: the first failure is fatal
            from .api import _raise
            b = int(np.nonzero(status)[0][0])
            _raise(int(status[b]), f"find: model for branch {chunk[b][0]}, time {chunk[b][1]} is invalid")
        lls[i0:i0 + len(chunk)] = logl
    for (k, t), ll in zip(pairs, lls):
        print(f"branch={k}, time={hs_show(t)}, ll={hs_show(ll)}", file=sys.stderr)
    with open(a.evalFile, "w") as h:                                            # writeResult, Find.hs:98-103
        h.write("Branch\tTime\tLikelihood\n")
        for (k, t), ll in zip(pairs, lls):
            h.write(f"{k}\t{hs_show(t)}\t{hs_show(ll)}\n")
    # maximumBy keeps the LAST of equal maxima
    best = max(range(len(pairs)), key=lambda i: (lls[i], i))
    print(f"highest likelihood point:\nbranch {pairs[best][0]}\ntime {hs_show(pairs[best][1])}\nlog-likelihood {hs_show(lls[best])}")
    print(f"find: {len(pairs)} likelihoods in {launches} launches", file=sys.stderr)
    return pairs, lls


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
    p.add_argument("--noPolish", action="store_true", help="skip the batched quasi-Newton and coordinate searches that follow every simplex run")
    p.add_argument("--parallel", type=int, default=4,
                   help="vertices reflected per iteration; each costs 4 speculative likelihoods in the same launch")
    p.set_defaults(fn=wrap(run_maxl))
    p = sub.add_parser("find", help="Explore where a branch joins the tree")
    _general(p); _model(p); _hist(p)
    p.add_argument("-n", "--queryName", required=True, help="model branch that does not merge onto the tree at any time")
    p.add_argument("-f", "--evalFile", required=True, help="file to write the list of computed likelihoods to")
    p.add_argument("-b", "--branchAge", type=float, default=0.0, help="sampling age of the query sample")
    p.add_argument("--deltaTime", type=float, default=0.0005)
    p.add_argument("--maxTime", type=float, default=0.025)
    p.add_argument("--batch", type=int, default=512, help="grid points evaluated per launch")
    p.set_defaults(fn=wrap(run_find))
    p = sub.add_parser("mcmc", help="Run MCMC on the model and the data")
    _general(p); _model(p); _hist(p)
    p.add_argument("--cycles", type=int, default=1000)
    p.add_argument("-o", "--prefix", required=True)
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--nrChains", type=int, default=1, help="independent chains advanced in one launch per step")
    p.set_defaults(fn=wrap(run_mcmc))
