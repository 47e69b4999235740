"""CPU tests of the product's host logic: the C-ABI library loads and exports every declared symbol,
the host-side mirrors agree with the oracle, and the planner's structure (live-set sizes through
Join/Split epochs, shortcut steps) reproduces the oracle's executed state-step counts.  No device work."""
import ctypes
import os
import re

import numpy as np
import pytest

import rarecoal_b200 as R
from rarecoal_b200 import _lib as L
from rarecoal_b200 import workloads as Wk
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "rarecoal_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(rc_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    lib = ctypes.CDLL(L.SO)
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)
    L.load()


def test_no_device_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(R.RarecoalDeviceError):
        R.Engine(0)


def test_time_grid_and_enumeration_match_oracle():
    assert np.array_equal(R.defaultTimes(), O.time_steps())
    assert np.array_equal(R.getTimeSteps(10000, 50, 5.0), O.time_steps(10000, 50, 5.0))
    for P, max_af, nvec in ((4, 4, [66, 44, 44, 28]), (5, 5, [20, 20, 2, 20, 20]), (3, 7, [2, 1, 9]), (8, 3, [100] * 8)):
        assert np.array_equal(R.computeAllConfigs(max_af, nvec), O.all_configs(max_af, nvec))
    for n, k in ((100, 3), (28, 10), (5, 0), (3, 5)):
        assert R.choose(n, k) == O.choose(n, k)


def test_standard_order_filters():
    # Utils.hs:169-177
    nvec = [3, 3, 3]
    full = R.computeStandardOrder(nvec, 4)
    assert len(full) == len(O.all_configs(4, nvec))
    assert (R.computeStandardOrder(nvec, 4, minAf=2).sum(axis=1) >= 2).all()
    assert (R.computeStandardOrder(nvec, 4, conditionOn=[1])[:, 1] > 0).all()
    ex = R.computeStandardOrder(nvec, 4, excludePatterns=[[1, 0, 0]])
    assert len(ex) == len(full) - 1


def test_state_space_mirror():
    # StateSpace/Test.hs:35-58,82-95
    ss, so = R.JointStateSpace(5, 4), O.StateSpace(5, 4)
    assert ss._jsNrStates == so.nr_states == 125
    for i in range(125):
        x = ss._jsIdToState(i)
        assert np.array_equal(x, so.id_to_state(i))
        assert ss._jsStateToId(x) == i
        assert np.array_equal(ss._jsX1up(i), so.x1up(i))
    for k in range(5):
        assert ss._jsX1(k) == so.x1(k)
    seeds = [ss._jsStateToId(s) for s in ([2, 2, 0, 0, 0], [0, 2, 2, 0, 0])]
    assert ss.fillUpStateSpace(seeds).tolist() == so.fill_up(seeds).tolist()
    assert ss._jsStateToId([5, 0, 0, 0, 0]) == -1


def test_validate_and_regularization_match_oracle(fourpop):
    spec = R.ModelSpec(4, R.defaultTimes(), 0.0005, [1.0] * 4, 10.0, False,
                       [R.ModelEvent.from_tuple(e) for e in fourpop["events"]])
    R.validateModel(spec)
    assert R.getRegularizationPenalty(spec) == O.reg_penalty(4, 10.0, fourpop["events"])
    bad = R.ModelSpec(2, R.defaultTimes(), 0.0005, [1.0, 1.0], 0.0, False,
                      [R.ModelEvent(0.1, R.Join(0, 1)), R.ModelEvent(0.2, R.SetPopSize(1, 2.0))])
    with pytest.raises(R.RarecoalModelException):
        R.validateModel(bad)
    with pytest.raises(R.RarecoalModelException):
        R.validateModel(R.ModelSpec(2, [], 0.1, [1.0, 1.5], 0.0, False, []))
    # branch indices outside [0, nrPops) must not touch the size vector (the reference's V.! / V.// fail cleanly)
    for ev in (R.SetPopSize(50000000, 2.0), R.SetPopSize(-1, 2.0), R.Join(0, 7), R.Join(9, 1)):
        with pytest.raises(R.RarecoalModelException):
            R.getRegularizationPenalty(R.ModelSpec(2, R.defaultTimes(), 0.0005, [1.0, 1.0], 10.0, False,
                                                   [R.ModelEvent(0.1, ev)]))


@pytest.mark.parametrize("max_af,want", [(4, 224261), (10, 26379903)])
def test_planner_state_steps_fourpop(fourpop, max_af, want):
    # executed state-steps of C1 / C2 — exercises Join, Split (with the b>0 filter) and shortcut
    std = R.computeStandardOrder(fourpop["nVec"], max_af)
    st = R.planStats(fourpop["nVec"], max_af, std, R.defaultTimes(), fourpop["events"])
    assert st["state_steps"] == want
    assert st["n_epochs"] == 7


def test_planner_state_steps_c3():
    nvec, pats, _ = Wk.c3_problem()
    assert len(pats) == 43757
    st = R.planStats(nvec, 10, pats, R.defaultTimes(), Wk.c3_events())
    assert st == dict(state_steps=298329136, slots0=545327, max_live=36, n_epochs=8, n_bins=st["n_bins"], checksum=st["checksum"])


@pytest.mark.parametrize("events,no_shortcut,T", [
    ([], False, 3043), ([], True, 300),
    ([(0.0, 1, 0, 1, 0.3), (0.00001, 1, 2, 0, 0.5), (0.02, 0, 0, 1, 0), (0.03, 0, 0, 2, 0)], False, 3043),
    ([(0.00006, 1, 0, 1, 0.3), (0.02, 0, 0, 1, 0), (0.03, 0, 0, 2, 0)], False, 3043),
    ([(0.002, 1, 0, 1, 0.0), (0.003, 1, 2, 1, 1.0), (0.01, 0, 0, 1, 0), (0.03, 0, 0, 2, 0)], False, 3043),
    ([(0.0, 3, 2, 0, 1.0), (0.00004, 1, 0, 2, 0.5), (0.01, 0, 0, 1, 0)], False, 500),
    ([(0.01, 0, 0, 1, 0), (25.0, 0, 0, 2, 0)], False, 400),
])
def test_planner_state_steps_match_oracle_on_edge_models(events, no_shortcut, T):
    nvec = [12, 12, 12]
    pats = R.computeStandardOrder(nvec, 5)
    times = R.defaultTimes()[:T]
    _, _, w = O.logl(3, 5, times, 0.001, [1.0] * 3, no_shortcut, events, nvec, pats, np.zeros(len(pats), dtype=np.int64), 0)
    st = R.planStats(nvec, 5, pats, times, events, no_shortcut)
    assert st["state_steps"] == w


def test_plan_self_check_on_every_workload_shape(fourpop):
    # rc_plan_stats also runs the planner's bounds / capacity self-check (every index a kernel dereferences, every
    # bin against the shared-memory limits of its kernel): trees (state and row mode), admixture, big-tier patterns
    T = R.defaultTimes()
    nv10 = [100] * 10
    assert R.planStats(nv10, 4, R.computeStandardOrder(nv10, 4), T, Wk.c5_events())["n_epochs"] == 13   # 9 joins + 3 pulses
    nvec, pats, _ = Wk.c3_problem(max_af=6)
    assert R.planStats(nvec, 6, pats, T, Wk.c3_events())["max_live"] <= 16
    std8 = R.computeStandardOrder(fourpop["nVec"], 8)
    assert R.planStats(fourpop["nVec"], 8, std8, T, fourpop["events"])["max_live"] > 256      # big tier engaged
    assert R.planStats([5, 5], 3, R.computeStandardOrder([5, 5], 3), T[:50], [], True)["n_epochs"] == 1


def test_plan_does_not_depend_on_the_host_threads_that_build_it(fourpop):
    # Every phase of the planner runs on a thread pool (per pattern, per epoch, per bin group, results put one behind the
    # other afterwards).  The plan — every array a kernel reads, rc_plan_last_checksum — must be the same with 1, 3 and 8
    # threads and from run to run: C3 (tree, pattern mode, direct hand-over), fourPop at maxAf 10 (Split pulses: rows, states,
    # big tier) and C5 (13 epochs).
    import subprocess, sys
    code = (
        "import sys, json; sys.path.insert(0, %r)\n"
        "import rarecoal_b200 as R\nfrom rarecoal_b200 import workloads as Wk\n"
        "T = R.defaultTimes(); out = []\n"
        "nvec, pats, _ = Wk.c3_problem(); out.append(R.planStats(nvec, 10, pats, T, Wk.c3_events())['checksum'])\n"
        "wl = Wk.get('c2'); nvec, pats, _ = wl.problem(); ev = wl.proposal_batches(1, 1, seed=3)[0][0]\n"
        "out.append(R.planStats(nvec, wl.max_af, pats, T, ev)['checksum'])\n"
        "nv = [100] * 10; out.append(R.planStats(nv, 4, R.computeStandardOrder(nv, 4), T, Wk.c5_events())['checksum'])\n"
        "print(json.dumps(out))\n" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sums = []
    for threads in ("1", "3", "8", "8"):
        env = dict(os.environ, RC_PLAN_THREADS=threads)
        r = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        sums.append(r.stdout.strip().splitlines()[-1])
    assert len(set(sums)) == 1, sums
    assert all(v != 0 for v in __import__("json").loads(sums[0]))


def test_c_abi_rejects_bad_arguments_without_device():
    lib = L.load()
    assert lib.rc_all_configs(0, 3, None, None, 0) == L.RC_ERR_ARG
    assert lib.rc_ss_create(3, 0) is None
    assert lib.rc_last_error(None) == b"null context"
    assert lib.rc_version().startswith(b"rarecoal_b200")


def test_growth_rate_desugaring_matches_the_numpy_statement():
    # SetGrowthRate (RC_EV_GROWTH) is sugar: rc_desugar_growth must produce exactly the per-grid-step SetPopSize list of the
    # independent numpy statement (workloads.desugar_growth_py), which is what the oracle is given in the GPU tests
    T = R.defaultTimes()
    cases = [Wk.c5_events(),
             [(0.0, 2, 0, 0, 2.0), (0.0, 4, 0, 0, 30.0), (0.003, 4, 0, 0, -10.0), (0.006, 4, 0, 0, 0.0), (0.01, 0, 1, 0, 0.0)],
             [(0.001, 4, 1, 0, 5.0), (0.002, 0, 0, 1, 0.0)],                       # growth ended by the Join that removes the branch
             [(0.0, 4, 0, 0, 1.0)]]                                                # growth to the end of the grid
    for ev in cases:
        got = R.desugarGrowth(T[:800], ev)
        want = Wk.desugar_growth_py(T[:800], ev)
        assert len(got) == len(want)
        for a, b in zip(got, want):
            assert a[:4] == tuple(b[:4]) and abs(a[4] - b[4]) <= 4e-16 * abs(b[4])
        assert all(e[1] != 4 for e in got)
    assert len(R.desugarGrowth(T[:800], cases[3])) == 800 - 0    # one SetPopSize per grid point after t0 = 0
    # validateModel: bad branch / non-finite rate
    for bad in (R.SetGrowthRate(5, 1.0), R.SetGrowthRate(0, float("nan"))):
        with pytest.raises(R.RarecoalModelException):
            R.validateModel(R.ModelSpec(2, T, 0.0005, [1.0, 1.0], 0.0, False, [R.ModelEvent(0.1, bad)]))
    # a grown branch used after it was joined away is still an illegal join (StateSpace.hs:176-180)
    with pytest.raises(R.RarecoalModelException):
        R.validateModel(R.ModelSpec(2, T, 0.0005, [1.0, 1.0], 0.0, False,
                                    [R.ModelEvent(0.1, R.Join(0, 1)), R.ModelEvent(0.2, R.SetGrowthRate(1, 2.0))]))
    # no penalty term and no tracked-size update from a growth phase
    spec = lambda evs: R.ModelSpec(2, T, 0.0005, [1.0, 1.0], 10.0, False, [R.ModelEvent.from_tuple(e) for e in evs])
    assert R.getRegularizationPenalty(spec([(0.0, 4, 0, 0, 3.0), (0.1, 0, 0, 1, 0.0)])) == \
        R.getRegularizationPenalty(spec([(0.1, 0, 0, 1, 0.0)]))
