"""Synthetic workloads named in BASELINE.json / SURVEY.md §8d (C3, C4, C5) and helpers to build them.
Events are plain tuples (t, type, k, l, v) as rc_event."""
import numpy as np

from . import _lib as L
from .api import computeStandardOrder

TOTAL_SITES = 2170616682     # exampleData/fourPop.histogram.txt:4


def k_event(t, to, frm, size):
    """Template command `K t to from size` = Join then SetPopSize at the same time (ModelTemplate.hs:197-203)."""
    return [(t, L.RC_EV_JOIN, to, frm, 0.0), (t, L.RC_EV_POPSIZE, to, 0, size)]


def c3_events(times=(0.0025, 0.003, 0.0035, 0.004, 0.008, 0.009, 0.02),
              leaf_sizes=None, join_sizes=(0.8, 1.2, 0.9, 1.1, 0.7, 1.3, 1.0)):
    """8-population balanced tree: 8 `P 0.0`, 7 `K` joins, 22 parameters."""
    if leaf_sizes is None:
        leaf_sizes = [1.0 + 0.25 * i for i in range(8)]
    ev = [(0.0, L.RC_EV_POPSIZE, i, 0, leaf_sizes[i]) for i in range(8)]
    joins = [(0, 1), (2, 3), (4, 5), (6, 7), (0, 2), (4, 6), (0, 4)]
    for (to, frm), t, p in zip(joins, times, join_sizes):
        ev += k_event(t, to, frm, p)
    return ev


def c3_params():
    return np.array([1.0 + 0.25 * i for i in range(8)] + [0.0025, 0.003, 0.0035, 0.004, 0.008, 0.009, 0.02]
                    + [0.8, 1.2, 0.9, 1.1, 0.7, 1.3, 1.0])


def c3_events_from_params(x):
    return c3_events(times=x[8:15], leaf_sizes=x[0:8], join_sizes=x[15:22])


def synthetic_counts(patterns, total_sites=TOTAL_SITES, seed=20240601):
    """counts_i ~ Poisson(TOTAL * q_i), q_i ∝ 1/(af_i^2 * #patterns(af_i)), sum q = 2.5e-3."""
    af = patterns.sum(axis=1)
    n_af = np.bincount(af)
    q = 1.0 / (af.astype(float) ** 2 * n_af[af])
    q *= 2.5e-3 / q.sum()
    rng = np.random.default_rng(seed)
    return rng.poisson(total_sites * q).astype(np.int64)


def c3_problem(n_pops=8, max_af=10, n_per_pop=100):
    nvec = [n_per_pop] * n_pops
    pats = computeStandardOrder(nvec, max_af)
    return nvec, pats, synthetic_counts(pats)


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


def c5_events(n_pops=10, times=None):
    """C5: 10-population caterpillar with 3 admixture pulses and size changes."""
    ev = [(0.0, L.RC_EV_POPSIZE, i, 0, 1.0 + 0.1 * i) for i in range(n_pops)]
    for j in range(1, n_pops):
        ev += k_event(0.002 * j, 0, j, 1.0 + 0.05 * j)
    for (t, to, frm, adm) in ((0.001, 1, 2, 0.05), (0.005, 4, 5, 0.1), (0.011, 7, 8, 0.2)):
        ev.append((t, L.RC_EV_SPLIT, to, frm, adm))
        ev.append((t, L.RC_EV_POPSIZE, to, 0, 1.5))
    return ev
