"""GPU tests of the command surface built on the path: `logl`, `prob`, batched `maxl` and `mcmc`
(python -m rarecoal_b200 ..., option names of src-rarecoal/Main.hs)."""
import io
import os
import sys
from contextlib import redirect_stdout

import numpy as np
import pytest

import rarecoal_b200 as R
from rarecoal_b200 import __main__ as cli
from rarecoal_b200 import frontend as F
from rarecoal_b200 import search as S

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TEMPLATE = os.path.join(GOLDEN, "fourPop.template.txt")
HIST = os.path.join(GOLDEN, "fourPop.histogram.txt")
PARAMS = os.path.join(GOLDEN, "fourPop.m4.mcmc.paramEstimates.txt")


def _run(argv):
    buf = io.StringIO()
    with redirect_stdout(buf):
        cli.main(argv)
    return buf.getvalue()


def test_logl_command_reproduces_the_reference_golden(fourpop):
    # config 1: rarecoal logl -T template -P params -i histogram -m 4 (regularization does not enter logl)
    out = _run(["logl", "-T", TEMPLATE, "-P", PARAMS, "-i", HIST, "-m", "4"])
    logl = float(out.strip())
    assert abs(-logl - fourpop["score_m4_maxl"]) / fourpop["score_m4_maxl"] < 1e-9


def test_prob_command(fivepop):
    tmpl = ("BRANCHES = [A, B, C, D, E]\nEVENTS = [J 0.0025 A B; J 0.006 C D; J 0.0075 C E; J 0.01 A C]")
    out = _run(["prob", "-t", tmpl, "[100,100,100,100,100]", "[0,1,2,0,1]"])
    assert abs(float(out) - 2.009581497800885e-6) / 2.009581497800885e-6 < 1e-10      # Core/Test.hs:177


def test_maxl_improves_a_perturbed_start(tmp_path, fourpop):
    # batched Nelder-Mead: start 3 % off the reference's MaxL point, must come back at least to the golden score
    prefix = str(tmp_path / "fit")
    xs = []
    for k in ("p_EUR", "p_SEA", "p_SIB"):
        xs += ["-X", f"{k}={fourpop['params'][k] * 1.03}"]
    _run(["maxl", "-T", TEMPLATE, "-P", PARAMS, "-i", HIST, "-m", "4", "--regularization", "0", "--maxCycles", "60",
          "--nrRestarts", "1", "--parallel", "4", "-o", prefix] + xs)
    lines = open(prefix + ".paramEstimates.txt").read().split("\n")
    assert lines[0].startswith("Score\t")
    score = float(lines[0].split("\t")[1])
    mt = F.getModelTemplate(TEMPLATE)
    start = dict(F.makeParameterDict(PARAMS, [(k, fourpop["params"][k] * 1.03) for k in ("p_EUR", "p_SEA", "p_SIB")]))
    hp = F.loadHistogram(F.HistogramOptions(HIST, 1, 4), mt.mtBranchNames)
    eng = R.Engine(0)
    eng.set_problem(hp.nVec, hp.maxAf, hp.patterns, hp.counts, hp.totalSites)
    opts = F.GeneralOptions(optRegPenalty=0.0)
    f = S.Objective(opts, mt, hp, eng)
    s0 = f(np.array([start[n] for n in f.names]))[0]
    assert score < s0                                   # better than the perturbed start
    assert [l.split("\t")[0] for l in lines[1:18]] == F.getParamNames(mt)
    # the written file is a valid -P input (maxl style) and reproduces its own score
    back = dict(F.makeParameterDict(prefix + ".paramEstimates.txt"))
    assert f(np.array([back[n] for n in f.names]))[0] == pytest.approx(score, rel=1e-12)
    tr = open(prefix + ".trace.txt").read().split("\n")
    assert tr[0].split("\t")[:3] == ["Nr", "-Log-Likelihood", "Simplex size"]
    assert open(prefix + ".frequencyFitTable.txt").readline().startswith("Pattern\tCount\tFreq\tTheoryFreq")
    assert len(open(prefix + ".summaryFitTable.txt").read().strip().split("\n")) == 1 + 4 + 10
    # constraint violations and invalid models are penalised, not fatal (MaxUtils.hs:50-51)
    bad = np.array([start[n] for n in f.names]); bad[f.names.index("tAdm_FAM_SIB")] = 1e-5
    neg = np.array([start[n] for n in f.names]); neg[f.names.index("p_EUR")] = -1.0
    sc = f(np.vstack([bad, neg, [start[n] for n in f.names]]))
    assert sc[0] == S.PENALTY and sc[1] == S.PENALTY and sc[2] == s0
    eng.close()


def test_mcmc_runs_chains_in_lockstep(tmp_path, fourpop):
    prefix = str(tmp_path / "chain")
    _run(["mcmc", "-T", TEMPLATE, "-P", PARAMS, "-i", HIST, "-m", "4", "--regularization", "0", "--cycles", "4",
          "--nrChains", "6", "--seed", "11", "-o", prefix])
    lines = open(prefix + ".paramEstimates.txt").read().split("\n")
    assert lines[0].startswith("# Nr Burnin Cycles:") and lines[2] == "Param\tMaxL\tLowerCI\tMedian\tUpperCI"
    assert lines[3].startswith("Score\t")
    best = float(lines[3].split("\t")[1])
    assert best <= fourpop["score_m4_maxl"] * (1 + 1e-6)          # started at the MaxL point; cannot get much worse
    # the mcmc output is a valid -P input (mcmc style, MaxL column)
    back = dict(F.makeParameterDict(prefix + ".paramEstimates.txt"))
    assert set(F.getParamNames(F.getModelTemplate(TEMPLATE))) <= set(back)
    tr = open(prefix + ".trace.txt").read().strip().split("\n")
    assert len(tr) >= 1 + 6 * 4 and len(tr[1].split("\t")) == 1 + 3 * 17
