"""Synthetic workloads named in BASELINE.json / SURVEY.md §8d (C1 ... C5) and helpers to build them.

Events are plain tuples (t, type, k, l, v) as rc_event.  This module is pure numpy and has no package-relative import at
module level, so bench.py's reference arm can load it by file path without importing rarecoal_b200 (whose library it must
not load); pattern enumeration is passed in (`all_configs(maxAf, nVec)`: the library's or the oracle's, they are equal).
"""
import json
import os

import numpy as np

EV_JOIN, EV_SPLIT, EV_POPSIZE, EV_FREEZE, EV_GROWTH = 0, 1, 2, 3, 4
TOTAL_SITES = 2170616682     # exampleData/fourPop.histogram.txt:4
_GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _default_all_configs():
    from .api import computeAllConfigs
    return computeAllConfigs


def standard_order(nvec, max_af, all_configs=None):
    """computeStandardOrder (Utils.hs:169-177) with minAf = 1 and no conditioning: computeAllConfigs itself."""
    return np.asarray((all_configs or _default_all_configs())(max_af, nvec), dtype=np.int32)


def k_event(t, to, frm, size):
    """Template command `K t to from size` = Join then SetPopSize at the same time (ModelTemplate.hs:197-203)."""
    return [(t, EV_JOIN, to, frm, 0.0), (t, EV_POPSIZE, to, 0, size)]


def synthetic_counts(patterns, total_sites=TOTAL_SITES, seed=20240601):
    """counts_i ~ Poisson(TOTAL * q_i), q_i ∝ 1/(af_i^2 * #patterns(af_i)), sum q = 2.5e-3."""
    af = patterns.sum(axis=1)
    n_af = np.bincount(af)
    q = 1.0 / (af.astype(float) ** 2 * n_af[af])
    q *= 2.5e-3 / q.sum()
    rng = np.random.default_rng(seed)
    return rng.poisson(total_sites * q).astype(np.int64)


def desugar_growth_py(times, events):
    """Pure-numpy statement of the SetGrowthRate sugar (include/rarecoal_b200.h, RC_EV_GROWTH): used to feed the oracle and
    to pin rc_desugar_growth in the tests.  Returns the explicit event list, stably sorted by time."""
    times = np.asarray(times, dtype=float)
    ev = sorted(events, key=lambda e: e[0])          # stable
    size, grow, out, gen = {}, {}, [], []

    def flush(k, t_stop):
        if k not in grow:
            return
        t0, n0, r = grow.pop(k)
        i = int(np.searchsorted(times, t0, side="right"))
        while i < len(times) and times[i] < t_stop:
            gen.append((float(times[i]), EV_POPSIZE, k, 0, float(n0 * np.exp(-(r * (times[i] - t0))))))
            i += 1
        size[k] = float(n0 * np.exp(-(r * (t_stop - t0)))) if np.isfinite(t_stop) else size.get(k, 1.0)

    for (t, ty, k, l, v) in ev:
        if ty == EV_GROWTH:
            flush(k, t)
            if v == 0.0:
                gen.append((t, EV_POPSIZE, k, 0, size.get(k, 1.0)))
            else:
                grow[k] = (t, size.get(k, 1.0), v)
            continue
        if ty == EV_POPSIZE:
            flush(k, t)
            size[k] = v
        if ty == EV_JOIN:
            flush(l, t)
        out.append((t, ty, k, l, v))
    for k in sorted(grow):
        flush(k, float("inf"))
    return sorted(out + gen, key=lambda e: e[0])


class Workload:
    """One benchmark / parity configuration: the histogram side (nVec, maxAf, patterns, counts, total sites), the model
    (parameter names, start vector, events as a function of the parameter vector) and its label."""

    def __init__(self, key, label, P, max_af, nvec, params, events_from_params, theta=0.0005, template=None,
                 counts=None, total_sites=TOTAL_SITES, patterns=None):
        self.key, self.label, self.P, self.max_af, self.nvec = key, label, P, max_af, list(nvec)
        self.names = [n for n, _ in params]
        self.x0 = np.array([v for _, v in params], dtype=float)
        self.events_from_params, self.theta, self.template = events_from_params, theta, template
        self._counts, self.total_sites, self._patterns = counts, total_sites, patterns

    def problem(self, all_configs=None):
        pats = self._patterns if self._patterns is not None else standard_order(self.nvec, self.max_af, all_configs)
        cnt = self._counts(pats) if callable(self._counts) else (self._counts if self._counts is not None else synthetic_counts(pats))
        return self.nvec, pats, cnt

    def events(self, x=None):
        return self.events_from_params(self.x0 if x is None else x)

    def proposal_batches(self, n_batches, batch, seed=7, rel=0.01):
        """Single-coordinate perturbations of x0 by a uniform factor in [1-rel, 1+rel] (an mcmc proposal / a simplex or
        finite-difference vertex changes one coordinate), `batch` models per batch."""
        rng = np.random.default_rng(seed)
        out = []
        for _ in range(n_batches):
            evs = []
            for _ in range(batch):
                x = self.x0.copy()
                j = rng.integers(0, len(x))
                x[j] *= 1.0 + rng.uniform(-rel, rel)
                evs.append(self.events_from_params(x))
            out.append(evs)
        return out


# ---- C1 / C2: the reference's own example (exampleData/fourPop.*), maxAf 4 / 10 ---------------------------------------
def _fourpop():
    return json.load(open(os.path.join(_GOLDEN, "fourpop.json")))


def fourpop_events_from_params(x, names=None):
    """instantiateModel (ModelTemplate.hs:176-213) of exampleData/fourPop.template.txt, written out: template order, K =
    Join then SetPopSize.  tests/test_frontend_cpu.py checks it against the template instantiated by the library."""
    f = _fourpop()
    d = dict(zip(names or list(f["params"].keys()), [float(v) for v in x]))
    E, S, I, F = 0, 1, 2, 3          # EUR, SEA, SIB, FAM
    ev = [(0.0, EV_POPSIZE, E, 0, d["p_EUR"]), (0.0, EV_POPSIZE, S, 0, d["p_SEA"]), (0.0, EV_POPSIZE, I, 0, d["p_SIB"]),
          (0.0, EV_POPSIZE, F, 0, d["p_FAM"])]
    ev += k_event(d["t_EUR_SIB"], E, I, d["p_EUR_SIB"]) + k_event(d["t_SEA_FAM"], S, F, d["p_SEA_FAM"])
    ev += k_event(d["t_EUR_SEA"], E, S, d["p_EUR_SEA"])
    ev += [(d["tAdm_FAM_SIB"], EV_SPLIT, F, I, d["adm_FAM_SIB"]), (d["tAdm_FAM_SIB"], EV_POPSIZE, F, 0, d["pAdm_FAM_SIB"]),
           (d["tAdm_SEA_SIB"], EV_SPLIT, S, I, d["adm_SEA_SIB"]), (d["tAdm_SEA_SIB"], EV_POPSIZE, S, 0, d["pAdm_SEA_SIB"]),
           (0.00022, EV_SPLIT, E, F, d["adm_EUR_FAM"])]
    return ev


def fourpop_workload(max_af, synthetic=False):
    f = _fourpop()
    params = list(f["params"].items())
    table = {tuple(p): c for p, c in zip(f["patterns"], f["counts"])}

    def real_counts(pats):
        return np.array([table.get(tuple(int(v) for v in p), 0) for p in pats], dtype=np.int64)

    key = "c1" if max_af == 4 else "c2"
    label = (f"{key.upper()}: exampleData/fourPop (4 pops, N={f['nVec']}, Join + 3 Split pulses), maxAf={max_af}, "
             f"{'synthetic' if synthetic else 'real'} counts")
    return Workload(key, label, 4, max_af, f["nVec"], params, fourpop_events_from_params, f["theta"],
                    template=open(os.path.join(_GOLDEN, "fourPop.template.txt")).read(),
                    counts=None if synthetic else real_counts, total_sites=f["totalSites"])


# ---- C3 / C4: synthetic 8-population balanced tree, maxAf 10 --------------------------------------------------------
C3_JOINS = [(0, 1), (2, 3), (4, 5), (6, 7), (0, 2), (4, 6), (0, 4)]
C3_TEMPLATE = ("BRANCHES = [P0, P1, P2, P3, P4, P5, P6, P7]\nEVENTS = [\n" +
               "".join(f"    P 0.0 P{i} <p_P{i}>;\n" for i in range(8)) +
               ";\n".join(f"    K <t_{a}{b}> P{a} P{b} <p_{a}{b}>" for a, b in C3_JOINS) + "\n]\n")


def c3_events(times=(0.0025, 0.003, 0.0035, 0.004, 0.008, 0.009, 0.02),
              leaf_sizes=None, join_sizes=(0.8, 1.2, 0.9, 1.1, 0.7, 1.3, 1.0)):
    """8-population balanced tree: 8 `P 0.0`, 7 `K` joins, 22 parameters."""
    if leaf_sizes is None:
        leaf_sizes = [1.0 + 0.25 * i for i in range(8)]
    ev = [(0.0, EV_POPSIZE, i, 0, leaf_sizes[i]) for i in range(8)]
    for (to, frm), t, p in zip(C3_JOINS, times, join_sizes):
        ev += k_event(t, to, frm, p)
    return ev


def c3_params():
    return np.array([1.0 + 0.25 * i for i in range(8)] + [0.0025, 0.003, 0.0035, 0.004, 0.008, 0.009, 0.02]
                    + [0.8, 1.2, 0.9, 1.1, 0.7, 1.3, 1.0])


def c3_param_names():
    """Optimiser-vector order of C3_TEMPLATE (getParamNames: time slot before value slot within a command)."""
    names = [f"p_P{i}" for i in range(8)]
    for a, b in C3_JOINS:
        names += [f"t_{a}{b}", f"p_{a}{b}"]
    return names


def c3_events_from_params(x):
    return c3_events(times=x[8:15], leaf_sizes=x[0:8], join_sizes=x[15:22])


def c3_problem(n_pops=8, max_af=10, n_per_pop=100, all_configs=None):
    nvec = [n_per_pop] * n_pops
    pats = standard_order(nvec, max_af, all_configs)
    return nvec, pats, synthetic_counts(pats)


def c3_workload(max_af=10):
    names = [f"p_P{i}" for i in range(8)] + [f"t_{a}{b}" for a, b in C3_JOINS] + [f"p_{a}{b}" for a, b in C3_JOINS]
    return Workload("c3", f"C3: synthetic 8-pop balanced tree, maxAf={max_af}, nVec=[100]*8, 3043-step grid", 8, max_af,
                    [100] * 8, list(zip(names, c3_params())), c3_events_from_params, template=C3_TEMPLATE)


def c4_batch(n_models):
    """C4: x0 with one coordinate perturbed by max(1e-4, |0.01 x_j|) per model (Maxl.hs:84), cycling over
    coordinates with alternating sign (initial simplex: 23 models; central differences: 44)."""
    x0 = c3_params()
    out = []
    for b in range(n_models):
        x = x0.copy()
        if b > 0:
            j = (b - 1) % len(x0)
            sgn = 1.0 if ((b - 1) // len(x0)) % 2 == 0 else -1.0
            x[j] += sgn * max(1e-4, abs(0.01 * x[j]))
        out.append(c3_events_from_params(x))
    return out


# ---- C5: synthetic 10-population admixture model, maxAf 4 ---------------------------------------------
# 10 `P 0.0`, 9 `K` joins (caterpillar into P0, t = 0.002 j), 3 `S` pulses (adm 0.05 / 0.1 / 0.2 at t = 0.001 / 0.005 /
# 0.011), each followed by `P` at the same time, and growth on two branches (`G`, SetGrowthRate: P0 until its first join
# resets the size, P9 until it joins at 0.018) — 36 parameters.
C5_PULSES = [(0.001, 1, 2, 0.05), (0.005, 4, 5, 0.1), (0.011, 7, 8, 0.2)]
C5_GROWTH = [(0, 60.0), (9, 25.0)]


def c5_template(n_pops=10):
    lines = [f"    P 0.0 P{i} <p_P{i}>" for i in range(n_pops)]
    lines += [f"    G 0.0 P{k} <g_P{k}>" for k, _ in C5_GROWTH]
    lines += [f"    K <t_0{j}> P0 P{j} <p_0{j}>" for j in range(1, n_pops)]
    for q, (t, to, frm, _) in enumerate(C5_PULSES):
        lines += [f"    S <tAdm_{q}> P{to} P{frm} <adm_{q}>", f"    P <tAdm_{q}> P{to} <pAdm_{q}>"]
    return ("BRANCHES = [" + ", ".join(f"P{i}" for i in range(n_pops)) + "]\nEVENTS = [\n" + ";\n".join(lines) + "\n]\n")


def c5_param_dict(n_pops=10):
    """Parameters in the optimiser-vector order of c5_template (getParamNames, ModelTemplate.hs:147-174)."""
    d = [(f"p_P{i}", 1.0 + 0.1 * i) for i in range(n_pops)]
    d += [(f"g_P{k}", r) for k, r in C5_GROWTH]
    for j in range(1, n_pops):
        d += [(f"t_0{j}", 0.002 * j), (f"p_0{j}", 1.0 + 0.05 * j)]
    for q, (t, _, _, adm) in enumerate(C5_PULSES):
        d += [(f"adm_{q}", adm), (f"tAdm_{q}", t), (f"pAdm_{q}", 1.5)]      # a repeated name sits at its LAST appearance
    return d


def c5_events_from_params(x, n_pops=10):
    """instantiateModel of c5_template written out (template order); growth events stay SetGrowthRate (type 4)."""
    d = dict(zip([n for n, _ in c5_param_dict(n_pops)], [float(v) for v in x]))
    ev = [(0.0, EV_POPSIZE, i, 0, d[f"p_P{i}"]) for i in range(n_pops)]
    ev += [(0.0, EV_GROWTH, k, 0, d[f"g_P{k}"]) for k, _ in C5_GROWTH]
    for j in range(1, n_pops):
        ev += k_event(d[f"t_0{j}"], 0, j, d[f"p_0{j}"])
    for q, (_, to, frm, _) in enumerate(C5_PULSES):
        ev += [(d[f"tAdm_{q}"], EV_SPLIT, to, frm, d[f"adm_{q}"]), (d[f"tAdm_{q}"], EV_POPSIZE, to, 0, d[f"pAdm_{q}"])]
    return ev


def c5_events(n_pops=10):
    return c5_events_from_params([v for _, v in c5_param_dict(n_pops)], n_pops)


def c5_workload(max_af=4):
    return Workload("c5", f"C5: synthetic 10-pop caterpillar, 3 admixture pulses, growth on 2 branches, maxAf={max_af}, "
                          "nVec=[100]*10, 3043-step grid", 10, max_af, [100] * 10, c5_param_dict(), c5_events_from_params,
                    template=c5_template())


def get(key):
    return {"c1": lambda: fourpop_workload(4), "c2": lambda: fourpop_workload(10), "c2s": lambda: fourpop_workload(10, True),
            "c3": c3_workload, "c5": c5_workload}[key]()
