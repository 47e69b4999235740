"""Pins the CPU oracle against every known answer the reference ships for the hot path
(SURVEY.md §8c).  CPU only."""
import numpy as np
import pytest

from oracle import oracle as O


def test_time_grid(default_times):
    # Utils.hs:61-73: getTimeSteps 20000 400 20.0 -> nr_steps-1 = 3043 points, ends below tMax
    assert len(default_times) == 3043
    assert default_times[0] == pytest.approx(2.5003e-5, rel=1e-4)
    assert default_times[-1] == pytest.approx(19.9501, rel=1e-5)
    assert np.all(np.diff(default_times) > 0)


def test_getprob_known_answers(fivepop, default_times):
    # Core/Test.hs:167-184 (ref tolerance 1e-8; ours 1e-10)
    for e in fivepop["kat"]:
        ss = O.StateSpace(5, max(4, sum(e["config"])))
        p = O.get_prob(ss, default_times, fivepop["theta"], fivepop["disc"], False, fivepop["events"],
                       fivepop["nVec"], e["config"])
        assert abs(p - e["prob"]) / e["prob"] < 1e-10, e


def _std_counts(f, max_af):
    std = O.all_configs(max_af, f["nVec"])
    d = {tuple(p): c for p, c in zip(f["patterns"], f["counts"])}
    return std, [d.get(tuple(p), 0) for p in std]


def test_fourpop_whole_path_golden(fourpop, default_times):
    # exampleData/fourPop.m4.mcmc.paramEstimates.txt:4: Score(MaxL) = -logl (regularization 0)
    std, counts = _std_counts(fourpop, 4)
    assert len(std) == 69
    ll, probs, w = O.logl(4, 4, default_times, fourpop["theta"], [1.0] * 4, False, fourpop["events"],
                          fourpop["nVec"], std, counts, fourpop["totalSites"])
    assert abs(-ll - fourpop["score_m4_maxl"]) / fourpop["score_m4_maxl"] < 1e-12
    assert w == 224261                      # SURVEY.md §8d state-steps, C1
    assert np.all(probs > 0)


def test_fourpop_m10_state_steps(fourpop, default_times):
    std, counts = _std_counts(fourpop, 10)
    assert len(std) == 1000
    ll, probs, w = O.logl(4, 10, default_times, fourpop["theta"], [1.0] * 4, False, fourpop["events"],
                          fourpop["nVec"], std, counts, fourpop["totalSites"])
    assert w == 26379903                    # SURVEY.md §8d state-steps, C2
    assert ll == pytest.approx(-76350242.12181702, rel=1e-12)


def test_fillup_known_answer(fivepop):
    # StateSpace/Test.hs:82-95
    fx = fivepop["fillup"]
    ss = O.StateSpace(fx["nrPops"], fx["maxAf"])
    assert ss.nr_states == 125
    got = ss.fill_up([ss.state_to_id(s) for s in fx["seeds"]])
    want = sorted(ss.state_to_id(s) for s in fx["expected"])
    assert sorted(got.tolist()) == want


def test_state_space_round_trips():
    # StateSpace/Test.hs:35-58
    ss = O.StateSpace(5, 4)
    for i in range(ss.nr_states):
        assert ss.state_to_id(ss.id_to_state(i)) == i
    for k in range(5):
        e = [0] * 5; e[k] = 1
        assert ss.x1(k) == ss.state_to_id(e)
    # x1up: -1 exactly when the sum would exceed maxAf (StateSpace.hs:77-79)
    for i in range(ss.nr_states):
        x = ss.id_to_state(i)
        up = ss.x1up(i)
        for k in range(5):
            if x.sum() + 1 > 4:
                assert up[k] == -1
            else:
                y = x.copy(); y[k] += 1
                assert up[k] == ss.state_to_id(y)


def test_all_configs_matches_crude():
    # StateSpace/Test.hs:66-80: computeAllConfigs == computeAllConfigsCrude as sets, nVec=[20,20,2,20,20], maxAf 5
    nvec = [20, 20, 2, 20, 20]
    new = {tuple(p) for p in O.all_configs(5, nvec)}
    crude = set()
    import itertools
    for v in itertools.product(range(6), repeat=5):
        if 0 < sum(v) <= 5 and all(a <= b for a, b in zip(v, nvec)):
            crude.add(v)
    assert new == crude
    assert len(O.all_configs(5, nvec)) == len(new)


def test_enumeration_order():
    # Utils.hs:100-127 order
    got = O.all_configs(2, [2, 2, 2, 2])[:6].tolist()
    assert got == [[0, 0, 0, 1], [0, 0, 1, 0], [0, 1, 0, 0], [1, 0, 0, 0], [0, 0, 0, 2], [0, 0, 1, 1]]


def test_join_split_conservation():
    # Core/Test.hs:35-127: popJoin/popSplit conserve sum(a), sum(b); Split with m=1 == Join
    rng = np.random.default_rng(1)
    ss = O.StateSpace(5, 4)
    for _ in range(50):
        k, l = rng.choice(5, 2, replace=False)
        a = rng.integers(0, 100, 5).astype(float)
        assert abs(O.pop_join_a(a, k, l).sum() - a.sum()) < 1e-12
        m = rng.random()
        assert abs(O.pop_split_a(a, k, l, m).sum() - a.sum()) < 1e-12
        assert np.array_equal(O.pop_split_a(a, k, l, 1.0), O.pop_join_a(a, k, l))
        seed = int(rng.integers(0, ss.nr_states))
        live = ss.fill_up([seed])
        b = np.zeros(ss.nr_states); b[live] = rng.random(len(live))
        bj, lj = ss.pop_join_b(b, live, k, l)
        assert abs(bj.sum() - b.sum()) < 1e-12
        bs, ls = ss.pop_split_b(b, live, k, l, m)
        assert abs(bs.sum() - b.sum()) < 1e-12
        b1, l1 = ss.pop_split_b(b, live, k, l, 1.0)
        assert np.array_equal(b1, bj)
        assert set(np.nonzero(bj)[0]) <= set(lj.tolist())


def test_validate_model():
    # StateSpace.hs:163-197
    ok = [(0.1, O.EV_JOIN, 0, 1, 0.0), (0.2, O.EV_POPSIZE, 0, 0, 2.0)]
    O.validate_model(2, [1.0, 1.0], ok)
    for bad in ([(-0.1, O.EV_JOIN, 0, 1, 0.0)],
                [(0.1, O.EV_JOIN, 0, 1, 0.0), (0.2, O.EV_POPSIZE, 1, 0, 2.0)],   # branch used after join
                [(0.1, O.EV_JOIN, 0, 0, 0.0)],
                [(0.1, O.EV_POPSIZE, 0, 0, 0.0)],
                [(0.1, O.EV_SPLIT, 0, 1, 1.5)],
                [(0.1, O.EV_JOIN, 0, 2, 0.0)]):
        with pytest.raises(O.OracleError) as ei:
            O.validate_model(2, [1.0, 1.0], bad)
        assert ei.value.code == 1
    with pytest.raises(O.OracleError):
        O.validate_model(2, [1.0, 0.0], ok)


def test_reg_penalty(fourpop):
    # StateSpace.hs:199-221; with reg=10 the fourPop MaxL column gives +2327.892...
    pen = O.reg_penalty(4, 10.0, fourpop["events"])
    assert pen == pytest.approx(2327.892, rel=1e-5)
