"""CPU tests of the formats either side of the path and of the batched search drivers'
host logic (f-3).  No device work."""
import os

import numpy as np
import pytest

import rarecoal_b200 as R
from rarecoal_b200 import frontend as F
from rarecoal_b200 import search as S

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TEMPLATE = os.path.join(GOLDEN, "fourPop.template.txt")
HIST = os.path.join(GOLDEN, "fourPop.histogram.txt")
PARAMS = os.path.join(GOLDEN, "fourPop.m4.mcmc.paramEstimates.txt")


def test_template_ast_matches_reference_test():
    # src/RarecoalLib/ModelTemplate/Test.hs:10-35: the literal AST of exampleData/fourPop.template.txt
    V, X = F.ParamVariable, F.ParamFixed
    want = F.ModelTemplate(
        ["EUR", "SEA", "SIB", "FAM"],
        [F.MTPopSizeChange(X(0.0), "EUR", V("p_EUR")), F.MTPopSizeChange(X(0.0), "SEA", V("p_SEA")),
         F.MTPopSizeChange(X(0.0), "SIB", V("p_SIB")), F.MTPopSizeChange(X(0.0), "FAM", V("p_FAM")),
         F.MTJoinPopSizeChange(V("t_EUR_SIB"), "EUR", "SIB", V("p_EUR_SIB")),
         F.MTJoinPopSizeChange(V("t_SEA_FAM"), "SEA", "FAM", V("p_SEA_FAM")),
         F.MTJoinPopSizeChange(V("t_EUR_SEA"), "EUR", "SEA", V("p_EUR_SEA")),
         F.MTSplit(V("tAdm_FAM_SIB"), "FAM", "SIB", V("adm_FAM_SIB")),
         F.MTPopSizeChange(V("tAdm_FAM_SIB"), "FAM", V("pAdm_FAM_SIB")),
         F.MTSplit(V("tAdm_SEA_SIB"), "SEA", "SIB", V("adm_SEA_SIB")),
         F.MTPopSizeChange(V("tAdm_SEA_SIB"), "SEA", V("pAdm_SEA_SIB")),
         F.MTSplit(X(0.00022), "EUR", "FAM", V("adm_EUR_FAM"))],
        [F.MTConstraint(V("tAdm_FAM_SIB"), X(0.00022), "ConstraintOpGreater")])
    assert F.getModelTemplate(TEMPLATE) == want


def test_param_names_order():
    # ModelTemplate.hs:147-174: template order, a repeated name at its LAST appearance
    mt = F.getModelTemplate(TEMPLATE)
    assert F.getParamNames(mt) == ["p_EUR", "p_SEA", "p_SIB", "p_FAM", "t_EUR_SIB", "p_EUR_SIB", "t_SEA_FAM", "p_SEA_FAM",
                                   "t_EUR_SEA", "p_EUR_SEA", "adm_FAM_SIB", "tAdm_FAM_SIB", "pAdm_FAM_SIB", "adm_SEA_SIB",
                                   "tAdm_SEA_SIB", "pAdm_SEA_SIB", "adm_EUR_FAM"]
    mt2 = F.parseModelTemplate("BRANCHES = [A,B]\nEVENTS = [J <t> A B; F 0.1 A True; K 0.2 A B <p>; F <tf> B False]")
    assert F.getParamNames(mt2) == ["t", "p", "tf"]
    assert mt2.mtEvents[1] == F.MTFreeze(F.ParamFixed(0.1), "A", True) and mt2.mtConstraints == []


def test_instantiate_matches_fixture_and_parameter_files(fourpop):
    mt = F.getModelTemplate(TEMPLATE)
    pd = F.makeParameterDict(PARAMS)                      # mcmc-style file: '#', 3 lines skipped, MaxL column
    assert pd[0][0] == "Score" and len(pd) == 18          # the Score line is read as a harmless parameter
    assert dict(pd)["p_EUR"] == fourpop["params"]["p_EUR"]
    ev = F.instantiateEvents(mt, pd)
    assert [tuple(e) for e in ev] == [tuple(e) for e in fourpop["events"]]          # K = Join then SetPopSize
    # -X settings come first and win (ModelTemplate.hs:250-263)
    assert dict(F.makeParameterDict(PARAMS, [("p_EUR", 3.0)]))["p_EUR"] == 3.0
    spec = F.instantiateModel(F.GeneralOptions(), mt, pd)
    assert spec.mNrPops == 4 and len(spec.mTimeSteps) == 3043 and spec.mDiscoveryRates == [1.0] * 4
    R.validateModel(spec)


def test_maxl_style_parameter_file(tmp_path):
    p = tmp_path / "x.paramEstimates.txt"
    p.write_text("Score\t5.5e7\np_A\t1.5\nt_AB\t2.0e-2\n")
    assert F.makeParameterDict(str(p)) == [("Score", 5.5e7), ("p_A", 1.5), ("t_AB", 0.02)]


def test_template_errors():
    mt = F.getModelTemplate(TEMPLATE)
    pd = dict(F.makeParameterDict(PARAMS))
    with pytest.raises(F.RarecoalModeltemplateException):           # missing parameter (ModelTemplate.hs:226)
        F.instantiateEvents(mt, [(k, v) for k, v in pd.items() if k != "p_SIB"])
    bad = dict(pd); bad["tAdm_FAM_SIB"] = 0.0001                     # constraint C <tAdm_FAM_SIB> > 0.00022
    with pytest.raises(R.RarecoalModelException):
        F.instantiateEvents(mt, list(bad.items()))
    with pytest.raises(F.RarecoalModeltemplateException):
        F.parseModelTemplate("BRANCHES = [A, B]\nEVENTS = [Q 0.1 A B]")
    with pytest.raises(R.RarecoalModelException):                    # unknown branch (ModelTemplate.hs:215-219)
        F.instantiateEvents(F.parseModelTemplate("BRANCHES = [A, B]\nEVENTS = [J 0.1 A C]"), [])
    mt3 = F.parseModelTemplate("BRANCHES = [A, B]\nEVENTS = [P 0.0 A <p_A>; S <t> A B <adm_x>; J <tj> A B]")
    assert F.fillParameterDictWithDefaults(mt3, [("t", 0.1), ("tj", 0.2)]) == [("p_A", 1.0), ("t", 0.1), ("adm_x", 0.05),
                                                                                ("tj", 0.2)]
    with pytest.raises(F.RarecoalModeltemplateException):           # ModelTemplate.hs:278-281
        F.fillParameterDictWithDefaults(mt3, [("t", 0.1)])


def test_histogram_problem_matches_fixture(fourpop):
    hp = F.loadHistogram(F.HistogramOptions(HIST, 1, 4), fourpop["branches"])
    std = R.computeStandardOrder(fourpop["nVec"], 4)
    d = {tuple(p): c for p, c in zip(fourpop["patterns"], fourpop["counts"])}
    assert np.array_equal(hp.patterns, std)
    assert hp.counts.tolist() == [d.get(tuple(int(v) for v in p), 0) for p in std]
    assert hp.nVec.tolist() == fourpop["nVec"] and hp.totalSites == fourpop["totalSites"] and hp.maxAf == 4
    full = F.loadHistogram(F.HistogramOptions(HIST, 1, 0), fourpop["branches"])       # maxAf 0 = the file's MAX_M
    assert full.maxAf == 10 and len(full.patterns) == 1000


def test_histogram_filters_and_branch_mapping(fourpop):
    # Utils.hs:130-177,114-119: minAf, conditionOn, excludePatterns, ghost branch, permuted model branches
    hp = F.loadHistogram(F.HistogramOptions(HIST, 2, 4, [2], [[0, 0, 2, 0]]), fourpop["branches"])
    assert (hp.patterns.sum(axis=1) >= 2).all() and (hp.patterns[:, 2] > 0).all()
    assert [0, 0, 2, 0] not in hp.patterns.tolist()
    perm = ["GHOST", "FAM", "EUR", "SIB", "SEA"]
    hp2 = F.loadHistogram(F.HistogramOptions(HIST, 1, 4), perm)
    base = F.loadHistogram(F.HistogramOptions(HIST, 1, 4), fourpop["branches"])
    assert hp2.nVec.tolist() == [0, 28, 66, 44, 44]
    assert np.array_equal(hp2.patterns[:, [2, 4, 3, 1]], base.patterns) and (hp2.patterns[:, 0] == 0).all()
    assert np.array_equal(hp2.counts, base.counts)
    with pytest.raises(R.RarecoalHistogramException):
        F.loadHistogram(F.HistogramOptions(HIST, 1, 4), ["EUR", "SEA", "SIB"])          # FAM missing in the model
    with pytest.raises(R.RarecoalHistogramException):
        F.loadHistogram(F.HistogramOptions(HIST, 1, 11), fourpop["branches"])           # illegal maxAF
    with pytest.raises(R.RarecoalHistogramException):
        F.loadHistogram(F.HistogramOptions("/nonexistent/file", 1, 4), fourpop["branches"])


def test_histogram_header_lists_are_replaced_by_options(tmp_path):
    # filterConditionOn / filterExcludePatterns (Utils.hs:129-149) replace raConditionOn / raExcludePatterns when the
    # option is given; computeStandardOrder (Utils.hs:169-177) then uses the option's lists only.  elemIndex takes the
    # first of two equal histogram names (Utils.hs:118).
    path = tmp_path / "h.txt"
    path.write_text("NAMES=A,B,A\nN=4,4,4\nMAX_M=3\nTOTAL_SITES=1000\nCONDITION_ON=0\nEXCLUDE_PATTERNS=1,1,0;1,0,0\n"
                    "1,0,0 5\n0,1,0 7\n1,1,0 3\n0,1,1 2\n")
    std = R.computeAllConfigs(3, [4, 4, 4])

    def expect(cond, excl):
        keep = [p for p in std.tolist() if all(p[i] > 0 for i in cond) and p not in excl]
        return [[p[0], p[1]] for p in keep]            # model branches [A, B]: A maps to the FIRST histogram column

    file_lists = F.loadHistogram(F.HistogramOptions(str(path), 1, 3), ["A", "B"])
    assert file_lists.patterns.tolist() == expect([0], [[1, 1, 0], [1, 0, 0]])
    cli_cond = F.loadHistogram(F.HistogramOptions(str(path), 1, 3, [1]), ["A", "B"])
    assert cli_cond.patterns.tolist() == expect([1], [[1, 1, 0], [1, 0, 0]])          # header's [0] is gone
    cli_excl = F.loadHistogram(F.HistogramOptions(str(path), 1, 3, [], [[0, 1, 1]]), ["A", "B"])
    assert cli_excl.patterns.tolist() == expect([0], [[0, 1, 1]])                     # header's exclusions are gone
    assert cli_excl.nVec.tolist() == [4, 4]
    with pytest.raises(R.RarecoalHistogramException):
        F.loadHistogram(F.HistogramOptions(str(path), 1, 3, [], [[0, 1]]), ["A", "B"])  # Utils.hs:143: wrong length
    with pytest.raises(R.RarecoalHistogramException):
        F.loadHistogram(F.HistogramOptions(str(path), 1, 3, [3]), ["A", "B"])           # Utils.hs:133


def test_haskell_show():
    for x, s in ((5.503134217048233e7, "5.503134217048233e7"), (1.9060110218875406, "1.9060110218875406"),
                 (1.5951118620536157e-2, "1.5951118620536157e-2"), (0.1, "0.1"), (1.0, "1.0"), (100.0, "100.0"),
                 (1.0e7, "1.0e7"), (9999999.0, "9999999.0"), (0.05, "5.0e-2"), (-2.5, "-2.5"), (0.0, "0.0"),
                 (1.0e-8, "1.0e-8"), (1e20, "1.0e20")):
        assert S.hs_show(x) == s


@pytest.mark.parametrize("par", [1, 3])
def test_batched_simplex_on_a_known_function(par):
    # host logic of the batched Nelder-Mead: converges on a shifted quadratic and evaluates in batches
    target = np.array([1.5, -0.25, 3.0, 0.7])
    calls = []

    def f(X):
        X = np.atleast_2d(X)
        calls.append(len(X))
        return np.sum((X - target) ** 2 * np.array([1.0, 4.0, 0.5, 2.0]), axis=1) + 7.0

    x = np.array([1.0, 1.0, 1.0, 1.0])
    for _ in range(3):
        x, trace = S.minimize_batched(f, x, 2000, par)
    assert np.allclose(x, target, atol=1e-6)
    assert calls[0] == 5 and max(calls) >= 4 * par            # start simplex in one batch, 4 speculative points per vertex
    assert trace[-1][2] < 1e-8 or len(trace) == 2000


def test_workload_templates_instantiate_to_the_written_out_events(fourpop):
    # bench.py's reference arm builds events without the library; they must be what the template front-end produces
    from rarecoal_b200 import workloads as Wk
    for wl in (Wk.get("c2"), Wk.get("c3"), Wk.get("c5")):
        mt = F.parseModelTemplate(wl.template)
        assert len(mt.mtBranchNames) == wl.P
        x = wl.x0 * (1.0 + 0.001 * np.arange(len(wl.x0)))
        d = list(zip(wl.names, x.tolist()))
        assert sorted(F.getParamNames(mt)) == sorted(wl.names)
        got = F.instantiateEvents(mt, d)
        assert [tuple(e) for e in got] == [tuple(map(float, e)) if False else (float(e[0]), e[1], e[2], e[3], float(e[4]))
                                           for e in wl.events(x)]
    mt = F.parseModelTemplate(Wk.get("c5").template)
    assert F.getParamNames(mt) == [n for n, _ in Wk.c5_param_dict()]           # optimiser-vector order
    assert any(isinstance(e, F.MTGrowthRate) for e in mt.mtEvents)
    assert F.getParamNames(F.parseModelTemplate(Wk.C3_TEMPLATE)) == Wk.c3_param_names()


def test_batched_instantiation_equals_one_at_a_time(fourpop):
    # rc_template_instantiate_batch (what the optimiser's objective calls once per step): same events, offsets, statuses and
    # regularisation penalties as instantiating and penalising model by model; a violated constraint gives a status, no events
    import ctypes as C
    from rarecoal_b200 import _lib as L
    from rarecoal_b200.api import _events_array
    lib = L.load()
    text = open(os.path.join(GOLDEN, "fourPop.template.txt")).read()          # CONSTRAINTS = [C <tAdm_FAM_SIB> > 0.00022]
    mt = F.parseModelTemplate(text)
    names = F.getParamNames(mt)
    x0 = np.array([fourpop["params"][n] for n in names])
    X = np.stack([x0, x0 * 1.01, x0 * 0.99, x0])
    X[3, names.index("tAdm_FAM_SIB")] = 0.0001            # violates the constraint
    B, k = X.shape
    names_c = (C.c_char_p * k)(*[n.encode() for n in names])
    cap = B * (2 * len(mt.mtEvents) + 1)
    ev = (L.rc_event * cap)()
    off, st, pen = np.zeros(B + 1, dtype=np.int32), np.zeros(B, dtype=np.int32), np.zeros(B)
    err = C.create_string_buffer(256)
    n = lib.rc_template_instantiate_batch(text.encode(), k, names_c, B, X.ctypes.data_as(C.POINTER(C.c_double)), 10.0, ev, cap,
                                          off.ctypes.data_as(C.POINTER(C.c_int32)), st.ctypes.data_as(C.POINTER(C.c_int32)),
                                          pen.ctypes.data_as(C.POINTER(C.c_double)), err, 256)
    assert n == off[B] and list(st) == [0, 0, 0, L.RC_ERR_MODEL] and off[3] == off[4]
    for b in range(3):
        one = F.instantiateEvents(mt, list(zip(names, X[b].tolist())))
        got = [(ev[i].t, ev[i].type, ev[i].k, ev[i].l, ev[i].v) for i in range(off[b], off[b + 1])]
        assert got == one
        assert pen[b] == lib.rc_reg_penalty(4, 10.0, _events_array(one), len(one))
    with pytest.raises(F.RarecoalModelException):
        F.instantiateEvents(mt, list(zip(names, X[3].tolist())))
