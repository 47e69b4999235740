"""Formats either side of the likelihood path, parsed by librarecoal_b200.so (rc_frontend.cpp):
model templates (ModelTemplate.hs:78-241), parameter files (ModelTemplate.hs:243-281) and rare-allele histograms
(Utils.hs:130-206).  Names follow the reference."""
import ctypes as C
import json
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib as L
from .api import (ModelEvent, ModelSpec, RarecoalException, RarecoalHistogramException, RarecoalModelException,
                  getTimeSteps)


class RarecoalModeltemplateException(RarecoalException):
    """Utils.hs:48."""


@dataclass
class GeneralOptions:
    """Utils.hs:28-36 with the CLI defaults of src-rarecoal/Main.hs:97-132."""
    optTheta: float = 0.0005
    optNrThreads: int = 0
    optNoShortcut: bool = False
    optRegPenalty: float = 10.0
    optN0: int = 20000
    optLinGen: int = 400
    optTMax: float = 20.0


@dataclass(frozen=True)
class ParamFixed:
    value: float


@dataclass(frozen=True)
class ParamVariable:
    name: str


def _spec(d):
    return ParamFixed(d["fixed"]) if "fixed" in d else ParamVariable(d["var"])


@dataclass(frozen=True)
class MTPopSizeChange:
    t: object
    branch: str
    size: object


@dataclass(frozen=True)
class MTJoin:
    t: object
    to: str
    frm: str


@dataclass(frozen=True)
class MTJoinPopSizeChange:
    t: object
    to: str
    frm: str
    size: object


@dataclass(frozen=True)
class MTSplit:
    t: object
    to: str
    frm: str
    rate: object


@dataclass(frozen=True)
class MTFreeze:
    t: object
    branch: str
    flag: bool


@dataclass(frozen=True)
class MTGrowthRate:
    """`G t branch rate` — SetGrowthRate, an extension of the reference grammar (see api.SetGrowthRate)."""
    t: object
    branch: str
    rate: object


@dataclass(frozen=True)
class MTConstraint:
    lhs: object
    rhs: object
    op: str          # "ConstraintOpGreater" / "ConstraintOpSmaller"


@dataclass
class ModelTemplate:
    """ModelTemplate.hs:35-40."""
    mtBranchNames: List[str]
    mtEvents: List[object]
    mtConstraints: List[MTConstraint]
    text: str = field(default="", compare=False, repr=False)
    paramNames: List[str] = field(default_factory=list, compare=False, repr=False)


def parseModelTemplate(text: str) -> ModelTemplate:
    """modelTemplateP (ModelTemplate.hs:78-145); getModelTemplate (ModelByText)."""
    lib = L.load()
    err = C.create_string_buffer(512)
    cap = 1 << 16
    while True:
        out = C.create_string_buffer(cap)
        rc = lib.rc_template_to_json(text.encode(), out, cap, err, 512)
        if rc == 0:
            break
        if rc < 0:
            raise RarecoalModeltemplateException(err.value.decode())
        cap = rc + 16
    d = json.loads(out.value.decode())
    evs = []
    for e in d["events"]:
        t, v = _spec(e["t"]), _spec(e["v"])
        evs.append({"P": lambda: MTPopSizeChange(t, e["a"], v), "J": lambda: MTJoin(t, e["a"], e["b"]),
                    "K": lambda: MTJoinPopSizeChange(t, e["a"], e["b"], v), "S": lambda: MTSplit(t, e["a"], e["b"], v),
                    "F": lambda: MTFreeze(t, e["a"], e["flag"]), "G": lambda: MTGrowthRate(t, e["a"], v)}[e["cmd"]]())
    cons = [MTConstraint(_spec(c["lhs"]), _spec(c["rhs"]),
                         "ConstraintOpGreater" if c["op"] == ">" else "ConstraintOpSmaller") for c in d["constraints"]]
    return ModelTemplate(d["branches"], evs, cons, text, d["params"])


def getModelTemplate(path: str) -> ModelTemplate:
    """getModelTemplate (ModelByFile), ModelTemplate.hs:59-71."""
    return parseModelTemplate(open(path).read())


def getParamNames(mt: ModelTemplate) -> List[str]:
    """ModelTemplate.hs:147-174."""
    return list(mt.paramNames)


def instantiateEvents(mt: ModelTemplate, paramsDict: Sequence[Tuple[str, float]]):
    """Events of instantiateModel as rc_event tuples (t, type, k, l, v)."""
    lib = L.load()
    names = (C.c_char_p * max(1, len(paramsDict)))(*[k.encode() for k, _ in paramsDict])
    vals = np.array([v for _, v in paramsDict] or [0.0], dtype=np.float64)
    cap = 2 * len(mt.mtEvents) + 1
    out = (L.rc_event * cap)()
    err = C.create_string_buffer(512)
    n = lib.rc_template_instantiate(mt.text.encode(), len(paramsDict), names, vals.ctypes.data_as(C.POINTER(C.c_double)),
                                    out, cap, err, 512)
    if n == -L.RC_ERR_MODEL:
        raise RarecoalModelException(err.value.decode())
    if n < 0:
        raise RarecoalModeltemplateException(err.value.decode())
    return [(out[i].t, out[i].type, out[i].k, out[i].l, out[i].v) for i in range(n)]


def instantiateModel(opts: GeneralOptions, mt: ModelTemplate, paramsDict: Sequence[Tuple[str, float]]) -> ModelSpec:
    """ModelTemplate.hs:176-213: discovery rates 1.0, grid from getTimeSteps n0 lingen tMax."""
    ev = instantiateEvents(mt, paramsDict)
    n = len(mt.mtBranchNames)
    return ModelSpec(n, getTimeSteps(opts.optN0, opts.optLinGen, opts.optTMax), opts.optTheta, [1.0] * n,
                     opts.optRegPenalty, opts.optNoShortcut, [ModelEvent.from_tuple(e) for e in ev])


def makeParameterDict(paramsFile: Optional[str], xSettings: Sequence[Tuple[str, float]] = ()) -> List[Tuple[str, float]]:
    """ModelTemplate.hs:243-263: -X settings first, then the file; the first occurrence of a key wins."""
    pairs = list(xSettings)
    if paramsFile:
        text = open(paramsFile).read()
        lib = L.load()
        cap = text.count("\n") + 2
        names = C.create_string_buffer(len(text) + 16)
        vals = np.zeros(cap)
        n = lib.rc_params_parse(text.encode(), names, len(text) + 16, vals.ctypes.data_as(C.POINTER(C.c_double)), cap)
        keys = names.value.decode().split("\n") if n > 0 else []
        pairs += list(zip(keys, vals[:n].tolist()))
    seen, out = set(), []
    for k, v in pairs:
        if k not in seen:
            seen.add(k)
            out.append((k, float(v)))
    return out


def fillParameterDictWithDefaults(mt: ModelTemplate, paramsDict):
    """ModelTemplate.hs:265-281: names starting with 'p' -> 1.0, with 'adm' -> 0.05, anything else is an error."""
    d = dict(paramsDict)
    out = []
    for p in getParamNames(mt):
        if p in d:
            out.append((p, d[p]))
        elif p.startswith("p"):
            out.append((p, 1.0))
        elif p.startswith("adm"):
            out.append((p, 0.05))
        else:
            raise RarecoalModeltemplateException(
                f"Don't know how to initialize parameter {p}. Please provide initial value via -X {p}=...")
    return out


@dataclass
class HistogramOptions:
    """Utils.hs:38-45 with the CLI defaults of Main.hs:174-201."""
    optHistPath: str = ""
    optMinAf: int = 1
    optMaxAf: int = 4
    optConditionOn: Sequence[int] = ()
    optExcludePatterns: Sequence[Sequence[int]] = ()
    optSiteReduction: float = 1.0


@dataclass
class HistogramProblem:
    """What computeFrequencySpectrum needs from a RareAlleleHistogram (MaxUtils.hs:97-113), already in model-branch
    order: the arguments of Engine.set_problem."""
    names: List[str]
    nVec: np.ndarray
    maxAf: int
    patterns: np.ndarray
    counts: np.ndarray
    totalSites: int
    siteRed: float


def loadHistogram(histOpts: HistogramOptions, modelBranches: Sequence[str]) -> HistogramProblem:
    """loadHistogram (Utils.hs:179-197) followed by computeStandardOrder and the model-branch mapping."""
    lib = L.load()
    err = C.create_string_buffer(512)
    h = lib.rc_hist_read(histOpts.optHistPath.encode(), err, 512)
    if not h:
        raise RarecoalHistogramException(err.value.decode())
    try:
        nP, nR, mn, mx, tot = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32(), C.c_int64()
        nv = np.zeros(64, dtype=np.int32)
        names = C.create_string_buffer(4096)
        lib.rc_hist_info(h, C.byref(nP), C.byref(nR), C.byref(mn), C.byref(mx), C.byref(tot),
                         nv.ctypes.data_as(C.POINTER(C.c_int32)), names, 4096)
        max_af = histOpts.optMaxAf if histOpts.optMaxAf else mx.value
        nB = len(modelBranches)
        mb = (C.c_char_p * nB)(*[b.encode() for b in modelBranches])
        if any(len(x) != nP.value for x in histOpts.optExcludePatterns):                 # Utils.hs:143
            raise RarecoalHistogramException("illegal excludePattern(s)" + str(list(histOpts.optExcludePatterns)))
        cond = np.array(list(histOpts.optConditionOn) or [0], dtype=np.int32)
        excl = np.array(list(histOpts.optExcludePatterns) or [[0] * nP.value], dtype=np.int32)
        ip = C.POINTER(C.c_int32)
        args = (h, int(histOpts.optMaxAf), int(histOpts.optMinAf), cond.ctypes.data_as(ip), len(histOpts.optConditionOn),
                excl.ctypes.data_as(ip), len(histOpts.optExcludePatterns), nB, mb)
        n = lib.rc_hist_problem(*args, None, None, None, 0, err, 512)
        if n < 0:
            raise RarecoalHistogramException(err.value.decode())
        pats = np.zeros((n, nB), dtype=np.int32)
        cnts = np.zeros(n, dtype=np.int64)
        nvec = np.zeros(nB, dtype=np.int32)
        lib.rc_hist_problem(*args, pats.ctypes.data_as(ip), cnts.ctypes.data_as(C.POINTER(C.c_int64)),
                            nvec.ctypes.data_as(ip), n, err, 512)
        return HistogramProblem(names.value.decode().split(","), nvec, max_af, pats, cnts, tot.value,
                                histOpts.optSiteReduction)
    finally:
        lib.rc_hist_free(h)
