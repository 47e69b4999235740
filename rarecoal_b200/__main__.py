"""Command surface of the accelerated path: `python -m rarecoal_b200 {logl,prob,maxl,mcmc} ...` with the reference's
option names (src-rarecoal/Main.hs:97-258, Logl.hs, Prob.hs, Maxl.hs, Mcmc.hs)."""
import argparse
import ast
import sys
import time

import numpy as np

from . import frontend as F
from .api import Engine, JointStateSpace


def _general(ap):
    ap.add_argument("--theta", type=float, default=0.0005)
    ap.add_argument("--nrThreads", type=int, default=0, help="ignored: the likelihood runs on the GPU")
    ap.add_argument("--noShortcut", action="store_true")
    ap.add_argument("--regularization", type=float, default=10.0)
    ap.add_argument("--n0", type=int, default=20000)
    ap.add_argument("--lingen", type=int, default=400)
    ap.add_argument("--tMax", type=float, default=20.0)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--nrGpus", type=int, default=1,
                    help="GPUs of this machine to use (rc_create_multi): a batch of models is sharded by model, a single "
                         "likelihood by pattern with an NCCL all-reduce of the partial sums")


def _model(ap):
    g = ap.add_mutually_exclusive_group(required=True)
    g.add_argument("-T", "--modelTemplate", dest="templateFile")
    g.add_argument("-t", "--modelTemplateString", dest="templateText")
    ap.add_argument("-P", "--paramsFile")
    ap.add_argument("-X", "--paramSetting", action="append", default=[], metavar="name=value")


def _hist(ap):
    ap.add_argument("-i", "--histogram", required=True)
    ap.add_argument("--minAf", type=int, default=1)
    ap.add_argument("-m", "--maxAf", type=int, default=4)
    ap.add_argument("--conditionOn", type=ast.literal_eval, default=[])
    ap.add_argument("--excludePattern", type=ast.literal_eval, action="append", default=[])
    ap.add_argument("--effectiveSitesReduction", type=float, default=1.0)


def _opts(a):
    return F.GeneralOptions(a.theta, a.nrThreads, a.noShortcut, a.regularization, a.n0, a.lingen, a.tMax)


def _template(a):
    return F.getModelTemplate(a.templateFile) if a.templateFile else F.parseModelTemplate(a.templateText)


def _params(a):
    xs = []
    for s in a.paramSetting:
        k, v = s.split("=", 1)
        xs.append((k, float(v)))
    return F.makeParameterDict(a.paramsFile, xs)


def _load(a, mt):
    ho = F.HistogramOptions(a.histogram, a.minAf, a.maxAf, a.conditionOn, a.excludePattern, a.effectiveSitesReduction)
    hp = F.loadHistogram(ho, mt.mtBranchNames)
    eng = Engine(a.device) if a.nrGpus <= 1 else Engine(nrGpus=a.nrGpus)
    eng.set_problem(hp.nVec, hp.maxAf, hp.patterns, hp.counts, hp.totalSites)
    return hp, eng


def run_logl(a):
    """Logl.hs:20-29."""
    mt = _template(a)
    spec = F.instantiateModel(_opts(a), mt, _params(a))
    hp, eng = _load(a, mt)
    print(repr(float(eng.computeLogLikelihood(spec, hp.siteRed))))


def run_prob(a):
    """Prob.hs:20-29."""
    mt = _template(a)
    spec = F.instantiateModel(_opts(a), mt, _params(a))
    eng = Engine(a.device)
    print(repr(float(eng.getProb(spec, JointStateSpace(spec.mNrPops, max(1, sum(a.kVec))), a.nVec, a.kVec))))


def main(argv=None):
    ap = argparse.ArgumentParser(prog="rarecoal_b200")
    sub = ap.add_subparsers(dest="cmd", required=True)
    p = sub.add_parser("logl", help="Compute the likelihood of the given model for the given data set")
    _general(p); _model(p); _hist(p)
    p.set_defaults(fn=run_logl)
    p = sub.add_parser("prob", help="Compute probability for a given configuration")
    _general(p); _model(p)
    p.add_argument("nVec", type=ast.literal_eval)
    p.add_argument("kVec", type=ast.literal_eval)
    p.set_defaults(fn=run_prob)
    from . import search
    search.add_parsers(sub, _general, _model, _hist, _opts, _template, _params, _load)
    a = ap.parse_args(argv)
    t0 = time.time()
    print(f"rarecoal_b200 starting {a.cmd}", file=sys.stderr)
    a.fn(a)
    print(f"rarecoal_b200 finished in {time.time() - t0:.2f}s", file=sys.stderr)


if __name__ == "__main__":
    main()
