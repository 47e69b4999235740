"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the reference's
known answers.  Tolerances are BASELINE.json's: per-pattern probabilities 1e-10 relative, total
log-likelihood 1e-9 relative."""
import numpy as np
import pytest

import rarecoal_b200 as R
from rarecoal_b200 import workloads as Wk
from rarecoal_b200 import _lib as L
from oracle import oracle as O

pytestmark = pytest.mark.gpu

PROB_RTOL = 1e-10
LOGL_RTOL = 1e-9


@pytest.fixture(scope="module", params=["keyed", "keyed-rows", "selfcontained"])
def eng(request):
    # every fast-tier code path: trajectory-key kernels with one pattern per thread (default), the same with pattern mode
    # off (row / state kernels), and the self-contained kernel
    e = R.Engine(0)
    e.set_option("keyed", 0.0 if request.param == "selfcontained" else 1.0)
    e.set_option("pattern_mode", 0.0 if request.param == "keyed-rows" else 1.0)
    yield e
    e.close()


def _spec(P, times, events, theta=0.0005, disc=None, no_shortcut=False):
    return R.ModelSpec(P, times, theta, disc or [1.0] * P, 0.0, no_shortcut,
                       [R.ModelEvent.from_tuple(e) for e in events])


def _counts(f, std):
    d = {tuple(p): c for p, c in zip(f["patterns"], f["counts"])}
    return np.array([d.get(tuple(int(v) for v in p), 0) for p in std], dtype=np.int64)


def _relerr(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    zero = b == 0
    assert np.array_equal(a[zero], b[zero])        # exact zeros (unreachable patterns) must be exact zeros
    if zero.all():
        return 0.0
    return np.max(np.abs(a[~zero] - b[~zero]) / np.abs(b[~zero]))


def test_getprob_known_answers(eng, fivepop, default_times):
    # Core/Test.hs:167-184
    spec = _spec(5, default_times, fivepop["events"], fivepop["theta"])
    for e in fivepop["kat"]:
        maxaf = max(4, sum(e["config"]))
        p = eng.getProb(spec, R.JointStateSpace(5, maxaf), fivepop["nVec"], e["config"])
        assert abs(p - e["prob"]) / e["prob"] < PROB_RTOL, (e, p)


@pytest.mark.parametrize("max_af", [4, 10])
def test_fourpop_vs_oracle_and_golden(eng, fourpop, default_times, max_af):
    # config 1 (maxAf=4, whole-path golden) and config 2's real-data twin (maxAf=10): Join+Split+SetPopSize
    std = R.computeStandardOrder(fourpop["nVec"], max_af)
    cnt = _counts(fourpop, std)
    eng.set_problem(fourpop["nVec"], max_af, std, cnt, fourpop["totalSites"])
    spec = _spec(4, default_times, fourpop["events"], fourpop["theta"])
    logl, probs, status = eng.eval_models([spec], want_probs=True)
    assert status[0] == 0
    ll_o, probs_o, w_o = O.logl(4, max_af, default_times, fourpop["theta"], [1.0] * 4, False, fourpop["events"],
                                fourpop["nVec"], std, cnt, fourpop["totalSites"])
    assert _relerr(probs[0], probs_o) < PROB_RTOL
    assert abs(logl[0] - ll_o) / abs(ll_o) < LOGL_RTOL
    assert eng.stats()["state_steps"] == w_o
    if max_af == 4:
        assert abs(-logl[0] - fourpop["score_m4_maxl"]) / fourpop["score_m4_maxl"] < LOGL_RTOL
    # reference-named entry points agree with the batch call
    assert eng.computeLogLikelihood(spec) == logl[0]
    spec_list = eng.computeFrequencySpectrum(spec)
    assert R.computeLogLikelihoodFromSpec(fourpop["totalSites"], 1.0, spec_list) == pytest.approx(logl[0], rel=1e-12)


def test_c3_eight_pop_tree_all_patterns_vs_oracle(eng, default_times):
    # config 3 — the configuration that carries the headline number — in full: every one of the 43 757 pattern
    # probabilities and the log-likelihood against the oracle (MaxUtils.hs:85-119)
    nvec, pats, cnt = Wk.c3_problem()
    eng.set_problem(nvec, 10, pats, cnt, Wk.TOTAL_SITES)
    ev = Wk.c3_events()
    spec = _spec(8, default_times, ev)
    logl, probs, status = eng.eval_models([spec], want_probs=True)
    assert status[0] == 0 and np.isfinite(logl[0])
    st = eng.stats()
    assert st["state_steps"] == 298329136
    ll_o, probs_o, w_o = O.logl(8, 10, default_times, 0.0005, [1.0] * 8, False, ev, nvec, pats, cnt, Wk.TOTAL_SITES)
    assert w_o == st["state_steps"]
    assert _relerr(probs[0], probs_o) < PROB_RTOL
    assert abs(logl[0] - ll_o) / abs(ll_o) < LOGL_RTOL
    # size-independent properties at full size
    p = probs[0]
    assert np.all(p > 0) and p.sum() < 1.0
    # batching is consistent
    logl2, probs2, _ = eng.eval_models([spec, spec, spec], want_probs=True)
    assert np.array_equal(probs2[0], p) and np.array_equal(probs2[2], p) and np.all(logl2 == logl[0])


def test_c4_batched_parameter_vectors(eng, default_times):
    # config 4 at its real size (maxAf 10, the 44 central-difference vectors of one maxl step in ONE launch): equals
    # one-at-a-time evaluation bit for bit, and the oracle in full for three of the models
    nvec, pats, cnt = Wk.c3_problem()
    eng.set_problem(nvec, 10, pats, cnt, Wk.TOTAL_SITES)
    batch = Wk.c4_batch(44)
    specs = [_spec(8, default_times, ev) for ev in batch]
    logl, probs, status = eng.eval_models(specs, want_probs=True)
    assert not status.any()
    for b in (0, 7, 43):
        l1, p1, _ = eng.eval_models([specs[b]], want_probs=True)
        assert np.array_equal(p1[0], probs[b]) and l1[0] == logl[b]
    for b in (1, 12, 36):
        ll_o, probs_o, _ = O.logl(8, 10, default_times, 0.0005, [1.0] * 8, False, batch[b], nvec, pats, cnt,
                                  Wk.TOTAL_SITES)
        assert _relerr(probs[b], probs_o) < PROB_RTOL
        assert abs(logl[b] - ll_o) / abs(ll_o) < LOGL_RTOL
    assert len(set(logl.tolist())) > 20     # the perturbations do change the likelihood


def test_c5_admixture_ten_pop_vs_oracle(eng, default_times):
    # config 5: 10 pops, maxAf 4, Join + Split pulses + size changes + SetGrowthRate on two branches.  The library gets the
    # SetGrowthRate events; the oracle (which has the reference's four event types only) gets the per-step SetPopSize
    # list of the independent numpy statement of the sugar.
    wl = Wk.c5_workload()
    nvec, pats, cnt = wl.problem()
    eng.set_problem(nvec, 4, pats, cnt, Wk.TOTAL_SITES)
    ev = wl.events()
    assert sum(1 for e in ev if e[1] == L.RC_EV_GROWTH) == 2
    ev_o = Wk.desugar_growth_py(default_times, ev)
    assert len(ev_o) > len(ev) + 300
    batch = [ev] + wl.proposal_batches(1, 3)[0]
    logl, probs, status = eng.eval_models([_spec(10, default_times, e) for e in batch], want_probs=True)
    assert not status.any()
    for b in (0, 2):
        e_o = Wk.desugar_growth_py(default_times, batch[b])
        ll_o, probs_o, w_o = O.logl(10, 4, default_times, 0.0005, [1.0] * 10, False, e_o, nvec, pats, cnt, Wk.TOTAL_SITES)
        assert _relerr(probs[b], probs_o) < PROB_RTOL
        assert abs(logl[b] - ll_o) / abs(ll_o) < LOGL_RTOL
    assert eng.stats()["state_steps"] > 0


def _small_problem(eng, P=3, max_af=5, n=12):
    nvec = [n] * P
    pats = R.computeStandardOrder(nvec, max_af)
    cnt = Wk.synthetic_counts(pats, total_sites=10 ** 7, seed=1)
    eng.set_problem(nvec, max_af, pats, cnt, 10 ** 7)
    return nvec, pats, cnt


@pytest.mark.parametrize("name,events,no_shortcut,disc", [
    ("no_events", [], False, None),
    ("no_events_noshortcut", [], True, None),
    ("tree_noshortcut", [(0.01, 0, 0, 1, 0), (0.03, 0, 0, 2, 0)], True, None),
    ("freeze", [(0.0, 3, 2, 0, 1.0), (0.004, 3, 2, 0, 0.0), (0.01, 0, 0, 1, 0), (0.03, 0, 0, 2, 0)], False, None),
    ("freeze_forever", [(0.0, 3, 2, 0, 1.0), (0.01, 0, 0, 1, 0)], False, None),
    ("split_m0_m1", [(0.002, 1, 0, 1, 0.0), (0.003, 1, 2, 1, 1.0), (0.01, 0, 0, 1, 0), (0.03, 0, 0, 2, 0)], False, None),
    ("early_split_before_first_step", [(0.0, 1, 0, 1, 0.3), (0.00001, 1, 2, 0, 0.5), (0.02, 0, 0, 1, 0), (0.03, 0, 0, 2, 0)],
     False, None),
    ("split_after_two_steps", [(0.00006, 1, 0, 1, 0.3), (0.02, 0, 0, 1, 0), (0.03, 0, 0, 2, 0)], False, None),
    ("same_time_events", [(0.01, 0, 0, 1, 0), (0.01, 2, 0, 0, 3.0), (0.01, 1, 0, 2, 0.25), (0.01, 2, 0, 0, 0.5),
                          (0.05, 0, 0, 2, 0)], False, None),
    ("event_beyond_grid", [(0.01, 0, 0, 1, 0), (25.0, 0, 0, 2, 0)], False, None),
    ("popsize_tiny_large", [(0.0, 2, 0, 0, 0.01), (0.0, 2, 1, 0, 50.0), (0.02, 0, 0, 1, 0), (0.04, 0, 0, 2, 0)], False, None),
    ("discovery_rates", [(0.01, 0, 0, 1, 0), (0.03, 0, 0, 2, 0)], False, [0.5, 1.0, 0.25]),
    ("unjoined_branch", [(0.01, 0, 0, 1, 0)], False, None),
])
def test_event_semantics_vs_oracle(eng, default_times, name, events, no_shortcut, disc):
    # strict t < nextTime firing, stable ties, freeze, split filter, shortcut on/off
    nvec, pats, cnt = _small_problem(eng)
    times = default_times if not no_shortcut and name != "unjoined_branch" and name != "freeze_forever" \
        and name != "event_beyond_grid" else default_times[:400]
    disc = disc or [1.0] * 3
    logl, probs, status = eng.eval_models([_spec(3, times, events, 0.001, disc, no_shortcut)], want_probs=True)
    assert status[0] == 0
    ll_o, probs_o, w_o = O.logl(3, 5, times, 0.001, disc, no_shortcut, events, nvec, pats, cnt, 10 ** 7)
    assert _relerr(probs[0], probs_o) < PROB_RTOL, name
    if np.isfinite(ll_o):
        assert abs(logl[0] - ll_o) / abs(ll_o) < LOGL_RTOL
    else:
        assert np.isnan(logl[0]) == np.isnan(ll_o) and (np.isnan(ll_o) or logl[0] == ll_o)
    assert eng.stats()["state_steps"] == w_o


def test_growth_rate_event_is_per_step_popsize(eng, default_times):
    # SetGrowthRate has no reference semantics; it is sugar for one SetPopSize per grid step, which the
    # oracle pins: explicit per-step list == the event == the oracle on the list
    nvec, pats, cnt = _small_problem(eng)
    times = default_times[:600]
    base = [(0.0, 2, 0, 0, 1.3), (0.01, 0, 0, 1, 0), (0.01, 2, 0, 0, 0.7), (0.02, 0, 0, 2, 0)]
    grow = base + [(0.0, 4, 0, 0, 40.0), (0.0, 4, 2, 0, -15.0), (0.004, 4, 2, 0, 0.0)]   # branch 0 until its K, branch 2 until 0.004
    explicit = Wk.desugar_growth_py(times, grow)
    assert all(e[1] != 4 for e in explicit) and len(explicit) > 400
    logl, probs, status = eng.eval_models([_spec(3, times, grow, 0.001), _spec(3, times, explicit, 0.001),
                                           _spec(3, times, base, 0.001)], want_probs=True)
    assert not status.any()
    ll_o, probs_o, _ = O.logl(3, 5, times, 0.001, [1.0] * 3, False, explicit, nvec, pats, cnt, 10 ** 7)
    assert _relerr(probs[0], probs_o) < PROB_RTOL
    assert abs(logl[0] - ll_o) / abs(ll_o) < LOGL_RTOL
    assert _relerr(probs[0], probs[1]) < 1e-13          # exp() on the host vs numpy: at most an ulp in a size
    assert logl[2] != logl[0]


def test_model_errors_and_statuses(eng, default_times):
    nvec, pats, cnt = _small_problem(eng)
    good = _spec(3, default_times, [(0.01, 0, 0, 1, 0), (0.03, 0, 0, 2, 0)])
    bad_reuse = _spec(3, default_times, [(0.01, 0, 0, 1, 0), (0.02, 2, 1, 0, 2.0)])     # branch used after join
    bad_neg = _spec(3, default_times, [(-0.01, 0, 0, 1, 0)])
    bad_disc = _spec(3, default_times, [], disc=[1.0, 0.0, 1.0])
    logl, probs, status = eng.eval_models([good, bad_reuse, bad_neg, good, bad_disc], want_probs=True)
    assert status.tolist() == [0, 1, 1, 0, 1]
    assert np.isfinite(logl[0]) and logl[0] == logl[3]
    assert np.isnan(logl[1]) and np.isnan(logl[2]) and np.isnan(logl[4])
    with pytest.raises(R.RarecoalModelException):
        eng.computeLogLikelihood(bad_reuse)
    with pytest.raises(R.RarecoalCompException):
        eng.getProb(good, R.JointStateSpace(3, 5), [12, 12], [1, 0])       # Core.hs:42
    # pattern with m_k > n_k: combinatorial factor 0 -> RarecoalCompException (Core.hs:53-56)
    eng.set_problem([2, 2, 2], 5, [[3, 0, 0], [1, 0, 0]], [1, 1], 100)
    _, _, st = eng.eval_models([good])
    assert st[0] == L.RC_ERR_COMP
    with pytest.raises(ValueError):
        eng.set_problem([2, 2, 2], 5, [[0, 0, 0]], [1], 100)           # zero state is not in the state space


def test_sharded_partials_sum_to_whole(default_times, fourpop):
    # multi-GPU path, ranks emulated as contexts on one device: partial sums add up to the single-GPU result
    import torch
    std = R.computeStandardOrder(fourpop["nVec"], 10)
    cnt = _counts(fourpop, std)
    spec = _spec(4, default_times, fourpop["events"], fourpop["theta"])
    whole = R.Engine(0)
    whole.set_problem(fourpop["nVec"], 10, std, cnt, fourpop["totalSites"])
    logl, probs, _ = whole.eval_models([spec, spec], want_probs=True)
    world = 3
    acc = torch.zeros(4, dtype=torch.float64, device="cuda")
    accp = torch.zeros(2 * len(std), dtype=torch.float64, device="cuda")
    engines = []
    for r in range(world):
        e = R.Engine(0)
        e.set_problem(fourpop["nVec"], 10, std, cnt, fourpop["totalSites"])
        e.set_shard(r, world)
        e.set_times(default_times)
        ev, off, theta, disc = e._pack([spec, spec])
        part = torch.zeros(4, dtype=torch.float64, device="cuda")
        pr = torch.zeros(2 * len(std), dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        st = e.eval_partial(ev, off, theta, disc, False, part.data_ptr(), pr.data_ptr(), None)
        torch.cuda.synchronize()
        assert not st.any()
        acc += part
        accp += pr
        engines.append(e)
    torch.cuda.synchronize()
    out = engines[0].finalize(2, acc.data_ptr())
    assert abs(out[0] - logl[0]) / abs(logl[0]) < 1e-13
    assert np.array_equal(accp.cpu().numpy().reshape(2, -1), probs)
    for e in engines:
        e.close()
    whole.close()


def test_large_batches_are_chunked_by_table_budget(default_times):
    # a batch whose trajectory tables exceed "ktab_total_mb" runs in chunks and gives the same answers; a model whose own
    # table exceeds the budget runs the self-contained kernels and does not shrink the chunks of the others
    e = R.Engine(0)
    nvec, pats, cnt = Wk.c3_problem(max_af=5)
    e.set_problem(nvec, 5, pats, cnt, Wk.TOTAL_SITES)
    specs = [_spec(8, default_times, ev) for ev in Wk.c4_batch(9)]
    logl, probs, status = e.eval_models(specs, want_probs=True)
    one_chunk = e.stats()
    assert one_chunk["chunks"] == 1 and one_chunk["models_selfcontained"] == 0
    per_model = one_chunk["ktab_mb"] / 9
    e.set_option("ktab_total_mb", 3.5 * per_model)          # three models per chunk
    e.eval_models(specs[:1])                                 # adapts the chunk size to the table of one model
    logl2, probs2, status2 = e.eval_models(specs, want_probs=True)
    st = e.stats()
    assert st["chunks"] == 3 and st["models_selfcontained"] == 0 and st["launches"] > one_chunk["launches"]
    assert st["state_steps"] == one_chunk["state_steps"]
    assert np.array_equal(probs2, probs) and np.array_equal(logl2, logl) and not status2.any()
    # budget below one model's table: every model falls back to the self-contained kernel, in ONE chunk
    e.set_option("ktab_total_mb", 0.5 * per_model)
    logl3, probs3, status3 = e.eval_models(specs, want_probs=True)
    st = e.stats()
    assert st["chunks"] == 1 and st["models_selfcontained"] == 9 and st["state_steps"] == one_chunk["state_steps"]
    # (closed-form trajectories of the keyed tier against the literal recursion of the self-contained one: DESIGN.md §4)
    assert _relerr(probs3, probs) < 5e-12 and np.max(np.abs(logl3 - logl) / np.abs(logl)) < 1e-12 and not status3.any()
    e.close()


@pytest.mark.parametrize("P,max_af,n", [(2, 12, 20), (3, 14, 15), (12, 3, 6), (9, 9, 4), (6, 11, 12), (1, 6, 9)])
def test_tier_selection_on_odd_shapes(eng, default_times, P, max_af, n):
    # shapes that exercise the tier choices: rows longer than the row kernel holds (state mode), more than 8 non-zero
    # branches (general tier), many branches with few alleles, a single population
    nvec = [n] * P
    pats = R.computeStandardOrder(nvec, max_af)
    if len(pats) > 1200:
        pats = pats[np.sort(np.random.default_rng(P * 100 + max_af).choice(len(pats), 1200, replace=False))]
    cnt = Wk.synthetic_counts(pats, total_sites=10 ** 8, seed=3)
    eng.set_problem(nvec, max_af, pats, cnt, 10 ** 8)
    times = default_times[:700]
    ev = [(0.0, 2, k, 0, 0.6 + 0.3 * k) for k in range(P)]
    t = 0.002
    for k in range(P - 1, 0, -1):                 # caterpillar: branch k joins k-1
        ev += Wk.k_event(t, k - 1, k, 0.8 + 0.1 * k)
        t += 0.0015
    logl, probs, status = eng.eval_models([_spec(P, times, ev, 0.001)], want_probs=True)
    assert status[0] == 0
    ll_o, probs_o, w_o = O.logl(P, max_af, times, 0.001, [1.0] * P, False, ev, nvec, pats, cnt, 10 ** 8)
    assert _relerr(probs[0], probs_o) < PROB_RTOL
    assert abs(logl[0] - ll_o) / abs(ll_o) < LOGL_RTOL
    assert eng.stats()["state_steps"] == w_o


def test_batch_mixing_shortcut_and_noshortcut_models(eng, fivepop, default_times):
    # mNoShortcut is a field of every ModelSpec (StateSpace.hs:49) but one flag per rc_eval call: the host layer must not
    # apply the first model's flag to the whole batch
    nvec = fivepop["nVec"]
    pats = R.computeStandardOrder(nvec, 3)
    cnt = Wk.synthetic_counts(pats, total_sites=10 ** 8, seed=5)
    eng.set_problem(nvec, 3, pats, cnt, 10 ** 8)
    times = default_times[:600]
    specs = [_spec(5, times, fivepop["events"], 0.001, no_shortcut=ns) for ns in (True, False, True, False)]
    logl, probs, status = eng.eval_models(specs, want_probs=True)
    assert not status.any()
    total = 0
    for m, ns in enumerate((True, False, True, False)):
        ll_o, probs_o, w_o = O.logl(5, 3, times, 0.001, [1.0] * 5, ns, fivepop["events"], nvec, pats, cnt, 10 ** 8)
        total += w_o
        assert _relerr(probs[m], probs_o) < PROB_RTOL
        assert abs(logl[m] - ll_o) / abs(ll_o) < LOGL_RTOL
    assert logl[0] != logl[1]
    assert eng.stats()["state_steps"] == total


def test_c3_batch_mixing_join_orders_uses_one_plan_per_order(default_times):
    # an optimiser or sampler over the seven join times of the 8-population tree crosses event orders: a batch that mixes
    # four orders is evaluated plan by plan (one transition plan per order, built once and kept in the LRU cache), every
    # model equal to its own single evaluation and — two of them, in full — to the oracle
    nvec, pats, cnt = Wk.c3_problem(max_af=6)
    e = R.Engine(0)
    e.set_problem(nvec, 6, pats, cnt, Wk.TOTAL_SITES)
    base = Wk.c3_params()
    orders = []
    for swap in (None, (8, 9), (9, 10), (12, 13)):           # indices of join times in the parameter vector
        x = base.copy()
        if swap:
            x[swap[0]], x[swap[1]] = x[swap[1]], x[swap[0]]
        orders.append(x)
    batch = [Wk.c3_events_from_params(orders[i % 4] * (1.0 + 0.001 * (i // 4))) for i in range(12)]
    specs = [_spec(8, default_times, ev) for ev in batch]
    logl, probs, status = e.eval_models(specs, want_probs=True)
    st = e.stats()
    assert not status.any() and st["plan_misses"] == 4
    for b in (1, 6, 11):
        l1, p1, _ = e.eval_models([specs[b]], want_probs=True)
        assert np.array_equal(p1[0], probs[b]) and l1[0] == logl[b]
        assert e.stats()["plan_misses"] == 0               # cached
    for b in (2, 3):
        ll_o, probs_o, _ = O.logl(8, 6, default_times, 0.0005, [1.0] * 8, False, batch[b], nvec, pats, cnt, Wk.TOTAL_SITES)
        assert _relerr(probs[b], probs_o) < PROB_RTOL and abs(logl[b] - ll_o) / abs(ll_o) < LOGL_RTOL
    # the cache keeps the most recently used plans when it is full
    e.set_option("plan_cache_max", 2)
    x5 = base.copy()
    x5[10], x5[11] = x5[11], x5[10]
    e.eval_models([_spec(8, default_times, Wk.c3_events_from_params(x5))])       # a fifth order: the least recently used plans go
    st = e.stats()
    assert st["plan_misses"] == 1 and st["plan_evictions"] == 3
    logl2, _, _ = e.eval_models(specs[:4])                                       # rebuilt on demand, same answers
    assert np.array_equal(logl2, logl[:4]) and e.stats()["plan_misses"] >= 3
    e.close()


def test_live_sets_beyond_shared_memory_run_in_global_memory(fourpop, default_times):
    # the reference's dense vectors take a live set of any size (Core.hs:58-70); here the big tier keeps b, the structure
    # words and the accumulators in a global scratch buffer when the largest live set does not fit its shared memory.
    # Forced on the fourPop maxAf=10 model (live sets of up to 668 states after the admixture pulses) with an 8 KB budget,
    # trajectory-key tier off so that the big tier runs these patterns; several models so that the scratch slices are used
    std = R.computeStandardOrder(fourpop["nVec"], 10)
    cnt = _counts(fourpop, std)
    e = R.Engine(0)
    e.set_option("keyed", 0.0)
    e.set_option("big_smem_kb", 8.0)
    e.set_problem(fourpop["nVec"], 10, std, cnt, fourpop["totalSites"])
    spec = _spec(4, default_times, fourpop["events"], fourpop["theta"])
    spec2 = _spec(4, default_times, fourpop["events"], fourpop["theta"] * 1.5)
    logl, probs, status = e.eval_models([spec, spec2, spec], want_probs=True)
    assert not status.any()
    ll_o, probs_o, w_o = O.logl(4, 10, default_times, fourpop["theta"], [1.0] * 4, False, fourpop["events"],
                                fourpop["nVec"], std, cnt, fourpop["totalSites"])
    assert _relerr(probs[0], probs_o) < PROB_RTOL and abs(logl[0] - ll_o) / abs(ll_o) < LOGL_RTOL
    assert np.array_equal(probs[0], probs[2]) and _relerr(probs[1], probs_o * 1.5) < PROB_RTOL
    # the same engine with the default budget gives the same numbers (same kernel, shared memory)
    e.set_option("big_smem_kb", 0.0)
    logl_s, probs_s, _ = e.eval_models([spec], want_probs=True)
    assert np.array_equal(probs_s[0], probs[0]) and logl_s[0] == logl[0]
    e.close()


@pytest.mark.parametrize("key", ["c2", "c5"])
def test_bounding_box_pattern_mode_after_splits(default_times, key):
    # "pattern_embed": the live sets a Split leaves (unions of boxes by support) run one pattern per thread on their bounding
    # boxes, the states that are not live carried along as exact zeros (rc_plan.cpp, pat_box).  Off by default (slower than
    # the row kernel on these workloads); same answers as the default path and as the oracle
    wl = Wk.get(key)
    nvec, pats, cnt = wl.problem()
    e = R.Engine(0)
    e.set_problem(nvec, wl.max_af, pats, cnt, wl.total_sites)
    evs = wl.proposal_batches(1, 5, seed=11)[0]
    specs = [_spec(wl.P, default_times, ev, wl.theta) for ev in evs]
    logl0, probs0, st0 = e.eval_models(specs, want_probs=True)
    steps0 = e.stats()["state_steps"]
    e.set_option("pattern_embed", 1.0)
    logl1, probs1, st1 = e.eval_models(specs, want_probs=True)
    assert not st0.any() and not st1.any() and e.stats()["state_steps"] == steps0
    assert _relerr(probs1, probs0) < PROB_RTOL and np.max(np.abs(logl1 - logl0) / np.abs(logl0)) < LOGL_RTOL
    ev0 = Wk.desugar_growth_py(default_times, evs[0])
    ll_o, probs_o, _ = O.logl(wl.P, wl.max_af, default_times, wl.theta, [1.0] * wl.P, False, ev0, nvec, pats, cnt, wl.total_sites)
    assert _relerr(probs1[0], probs_o) < PROB_RTOL and abs(logl1[0] - ll_o) / abs(ll_o) < LOGL_RTOL
    e.close()


def test_last_epoch_finished_by_the_launches_before_it(default_times):
    # a tree model's last epoch starts with the shortcut and executes no grid step: the launches of the last but one epoch
    # apply the Join, the shortcut and the epilogue themselves and the last epoch is not launched ("fuse_last", default on).
    # Same numbers bit for bit, fewer launches; a model whose last epoch does run steps (noShortcut) is not touched
    nvec, pats, cnt = Wk.c3_problem(max_af=6)
    e = R.Engine(0)
    e.set_option("graphs", 0.0)
    e.set_problem(nvec, 6, pats, cnt, Wk.TOTAL_SITES)
    specs = [_spec(8, default_times, ev) for ev in Wk.c4_batch(5)]
    logl1, probs1, st1 = e.eval_models(specs, want_probs=True)
    n1 = e.stats()["launches"]
    e.set_option("fuse_last", 0.0)
    logl0, probs0, st0 = e.eval_models(specs, want_probs=True)
    n0 = e.stats()["launches"]
    assert not st0.any() and not st1.any()
    assert np.array_equal(probs0, probs1) and np.array_equal(logl0, logl1) and n1 < n0
    e.set_option("fuse_last", 1.0)
    ns = [_spec(8, default_times, ev, no_shortcut=True) for ev in Wk.c4_batch(2)]
    la, pa, _ = e.eval_models(ns, want_probs=True)
    e.set_option("fuse_last", 0.0)
    lb, pb, _ = e.eval_models(ns, want_probs=True)
    assert np.array_equal(pa, pb) and np.array_equal(la, lb)
    e.close()


@pytest.mark.parametrize("key", ["c3", "c2"])
def test_closed_form_trajectories_against_the_literal_recursion(default_times, key):
    # The keyed tier takes a(t) inside an epoch from u = 1/a - 1 and the cumulative product of exp(-dt/2N) (rc_cum_kernel,
    # rc_traj_kernel: no sequential chain); the self-contained tier ("keyed" 0) iterates updateA literally (Core.hs:239-240)
    # as the oracle does.  Same mathematics, different rounding: the probabilities agree far inside the 1e-10 of the suite —
    # with and without the shortcut (3043 steps), with frozen branches and size changes in the batch.
    if key == "c3":
        nvec, pats, cnt = Wk.c3_problem(max_af=6)
        P, max_af, sites = 8, 6, Wk.TOTAL_SITES
        batch = Wk.c4_batch(5)
        batch[1] = batch[1] + [(0.0011, L.RC_EV_FREEZE, 2, 0, 1.0), (0.0061, L.RC_EV_FREEZE, 2, 0, 0.0), (0.005, L.RC_EV_POPSIZE, 0, 0, 0.05)]
        specs = [_spec(8, default_times, ev, no_shortcut=(i == 2)) for i, ev in enumerate(batch)]
    else:
        wl = Wk.get("c2")
        nvec, pats, cnt = wl.problem()
        P, max_af, sites = wl.P, wl.max_af, wl.total_sites
        specs = [_spec(P, default_times, ev, wl.theta) for ev in wl.proposal_batches(1, 3, seed=3)[0]]
    out = []
    for keyed in (1.0, 0.0):
        e = R.Engine(0)
        e.set_option("keyed", keyed)
        e.set_problem(nvec, max_af, pats, cnt, sites)
        logl, probs, st = e.eval_models(specs, want_probs=True)
        assert not st.any()
        assert (e.stats()["models_selfcontained"] == 0) == (keyed == 1.0)
        out.append((logl, probs))
        e.close()
    assert _relerr(out[0][1], out[1][1]) < 5e-12
    assert np.max(np.abs(out[0][0] - out[1][0]) / np.abs(out[1][0])) < 1e-12
