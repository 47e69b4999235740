"""Host-side mirror of the reference's library API for the likelihood path.

Names follow the reference (src/RarecoalLib/*.hs): ModelEvent / Join / Split / SetPopSize / SetFreeze /
ModelSpec (StateSpace.hs:30-51), JointStateSpace (StateSpace.hs:20-28), getProb (Core.hs:38),
computeFrequencySpectrum / computeLogLikelihood / computeLogLikelihoodFromSpec (MaxUtils.hs:78-119),
getTimeSteps / defaultTimes / computeAllConfigs / computeStandardOrder / choose (Utils.hs), and the
exception constructors of Utils.hs:47-52.  All arithmetic of the path runs in librarecoal_b200.so on the
GPU; this module only marshals arguments.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _lib as L


# ---- exceptions (Utils.hs:47-52) -------------------------------------------------------------------
class RarecoalException(Exception):
    pass


class RarecoalHistogramException(RarecoalException):
    pass


class RarecoalModelException(RarecoalException):
    pass


class RarecoalCompException(RarecoalException):
    pass


class RarecoalDeviceError(RuntimeError):
    """CUDA / library failure (no reference counterpart; there is no CPU fallback)."""


def _raise(code, msg):
    if code == L.RC_ERR_MODEL:
        raise RarecoalModelException(msg)
    if code == L.RC_ERR_COMP:
        raise RarecoalCompException(msg)
    if code == L.RC_ERR_ARG:
        raise ValueError(msg)
    raise RarecoalDeviceError(f"rarecoal_b200 error {code}: {msg}")


# ---- model types (StateSpace.hs:30-51) ----------------------------------------------------------------
@dataclass(frozen=True)
class Join:
    k: int
    l: int


@dataclass(frozen=True)
class Split:
    k: int
    l: int
    m: float


@dataclass(frozen=True)
class SetPopSize:
    k: int
    p: float


@dataclass(frozen=True)
class SetFreeze:
    k: int
    b: bool


@dataclass(frozen=True)
class SetGrowthRate:
    """Not a reference constructor (StateSpace.hs:36-40): sugar for one SetPopSize per grid step, see
    include/rarecoal_b200.h (RC_EV_GROWTH) and desugarGrowth."""
    k: int
    r: float


@dataclass(frozen=True)
class ModelEvent:
    meTime: float
    meEventType: object

    def as_tuple(self):
        e = self.meEventType
        if isinstance(e, Join):
            return (self.meTime, L.RC_EV_JOIN, e.k, e.l, 0.0)
        if isinstance(e, Split):
            return (self.meTime, L.RC_EV_SPLIT, e.k, e.l, e.m)
        if isinstance(e, SetPopSize):
            return (self.meTime, L.RC_EV_POPSIZE, e.k, 0, e.p)
        if isinstance(e, SetFreeze):
            return (self.meTime, L.RC_EV_FREEZE, e.k, 0, 1.0 if e.b else 0.0)
        if isinstance(e, SetGrowthRate):
            return (self.meTime, L.RC_EV_GROWTH, e.k, 0, e.r)
        raise TypeError(e)

    @staticmethod
    def from_tuple(t):
        tt, ty, k, l, v = t
        et = {L.RC_EV_JOIN: lambda: Join(int(k), int(l)), L.RC_EV_SPLIT: lambda: Split(int(k), int(l), float(v)),
              L.RC_EV_POPSIZE: lambda: SetPopSize(int(k), float(v)),
              L.RC_EV_FREEZE: lambda: SetFreeze(int(k), bool(v)),
              L.RC_EV_GROWTH: lambda: SetGrowthRate(int(k), float(v))}[int(ty)]()
        return ModelEvent(float(tt), et)


@dataclass
class ModelSpec:
    mNrPops: int
    mTimeSteps: Sequence[float]
    mTheta: float
    mDiscoveryRates: Sequence[float]
    mPopSizeRegularization: float = 0.0
    mNoShortcut: bool = False
    mEvents: List[ModelEvent] = field(default_factory=list)


def _events_array(events):
    arr = (L.rc_event * max(1, len(events)))()
    for i, e in enumerate(events):
        t = e.as_tuple() if isinstance(e, ModelEvent) else e
        arr[i] = L.rc_event(float(t[0]), int(t[1]), int(t[2]), int(t[3]), float(t[4]))
    return arr


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


# ---- small host-side functions ------------------------------------------------------------------------
def getTimeSteps(n0=20000, lingen=400, tMax=20.0):
    """Utils.hs:64-73."""
    lib = L.load()
    n = lib.rc_time_steps(n0, lingen, tMax, None, 0)
    out = np.zeros(n)
    lib.rc_time_steps(n0, lingen, tMax, _dp(out), n)
    return out


def defaultTimes():
    """Utils.hs:61-62."""
    return getTimeSteps(20000, 400, 20.0)


def computeAllConfigs(maxAf, nVec):
    """Utils.hs:100-104."""
    lib = L.load()
    nv = _i32(nVec)
    n = lib.rc_all_configs(maxAf, len(nv), _ip(nv), None, 0)
    out = np.zeros((n, len(nv)), dtype=np.int32)
    lib.rc_all_configs(maxAf, len(nv), _ip(nv), _ip(out), n)
    return out


def computeStandardOrder(nVec, maxAf, minAf=1, conditionOn=(), excludePatterns=()):
    """Utils.hs:169-177 on the histogram fields it reads."""
    pats = computeAllConfigs(maxAf, nVec)
    keep = pats.sum(axis=1) >= minAf
    for i in conditionOn:
        keep &= pats[:, i] > 0
    ex = {tuple(int(v) for v in p) for p in excludePatterns}
    if ex:
        keep &= np.array([tuple(int(v) for v in p) not in ex for p in pats])
    return pats[keep]


def choose(n, k):
    """Utils.hs:208-210."""
    return L.load().rc_choose(int(n), int(k))


def validateModel(modelSpec: ModelSpec):
    """StateSpace.hs:163-197."""
    err = C.create_string_buffer(256)
    ev = _events_array(modelSpec.mEvents)
    rc = L.load().rc_validate_model(modelSpec.mNrPops, _dp(_f64(modelSpec.mDiscoveryRates)), ev,
                                    len(modelSpec.mEvents), err, 256)
    if rc:
        _raise(rc, err.value.decode())


def getRegularizationPenalty(modelSpec: ModelSpec):
    """StateSpace.hs:199-221."""
    ev = _events_array(modelSpec.mEvents)
    v = L.load().rc_reg_penalty(modelSpec.mNrPops, modelSpec.mPopSizeRegularization, ev, len(modelSpec.mEvents))
    if v != v and not any(e.as_tuple()[4] != e.as_tuple()[4] for e in modelSpec.mEvents):
        # a branch index outside [0, nrPops): the reference's V.! / V.// fails here (StateSpace.hs:205-215)
        raise RarecoalModelException("illegal branch index in getRegularizationPenalty")
    return v


def desugarGrowth(times, events):
    """Events (ModelEvents or (t, type, k, l, v) tuples) with every SetGrowthRate rewritten into per-grid-step SetPopSize
    events, stably sorted by time (rc_desugar_growth) — the model the likelihood path actually evaluates."""
    t = _f64(times)
    ev = _events_array(events)
    lib = L.load()
    n = lib.rc_desugar_growth(len(t), _dp(t), ev, len(events), None, 0)
    if n < 0:
        raise ValueError("rc_desugar_growth: bad arguments")
    out = (L.rc_event * max(1, n))()
    lib.rc_desugar_growth(len(t), _dp(t), ev, len(events), out, n)
    return [(out[i].t, out[i].type, out[i].k, out[i].l, out[i].v) for i in range(n)]


def shardPatterns(patterns, rank, world):
    """Pattern indices that Engine.set_shard(rank, world) evaluates on `rank` (host only)."""
    pats = _i32(patterns)
    out = np.zeros(len(pats), dtype=np.int32)
    n = L.load().rc_shard_patterns(pats.shape[1], len(pats), _ip(pats), int(rank), int(world), _ip(out), len(out))
    if n < 0:
        raise ValueError("rc_shard_patterns: bad arguments")
    return out[:n]


def planStats(nVec, maxAf, patterns, times, events, noShortcut=False):
    """Host planner only (no device): dict(state_steps, slots0, max_live, n_epochs, n_bins)."""
    pats = _i32(patterns); nv = _i32(nVec); t = _f64(times)
    w, s0 = C.c_int64(0), C.c_int64(0)
    ml, ne, nb = C.c_int32(0), C.c_int32(0), C.c_int32(0)
    rc = L.load().rc_plan_stats(len(nv), maxAf, len(pats), _ip(pats), _ip(nv), len(t), _dp(t), _events_array(events),
                                len(events), int(noShortcut), C.byref(w), C.byref(s0), C.byref(ml), C.byref(ne),
                                C.byref(nb))
    if rc:
        _raise(rc, "rc_plan_stats failed")
    return dict(state_steps=w.value, slots0=s0.value, max_live=ml.value, n_epochs=ne.value, n_bins=nb.value,
                checksum=int(L.load().rc_plan_last_checksum()))


class JointStateSpace:
    """StateSpace.hs:20-28; makeJointStateSpace nrPop maxAf (StateSpace.hs:73-87)."""

    def __init__(self, nrPop, maxAf):
        self._lib = L.load()
        self._h = self._lib.rc_ss_create(nrPop, maxAf)
        if not self._h:
            raise ValueError("unsupported state space size")
        self._jsNrPop, self._jsMaxAf = nrPop, maxAf
        self._jsNrStates = self._lib.rc_ss_nr_states(self._h)

    def __del__(self):
        try:
            self._lib.rc_ss_destroy(self._h)
        except Exception:
            pass

    def _jsStateToId(self, state):
        return self._lib.rc_ss_state_to_id(self._h, _ip(_i32(state)))

    def _jsIdToState(self, i):
        out = np.zeros(self._jsNrPop, dtype=np.int32)
        if self._lib.rc_ss_id_to_state(self._h, int(i), _ip(out)):
            raise IndexError(i)
        return out

    def _jsX1up(self, i):
        out = np.zeros(self._jsNrPop, dtype=np.int32)
        if self._lib.rc_ss_x1up(self._h, int(i), _ip(out)):
            raise IndexError(i)
        return out

    def _jsX1(self, k):
        return self._lib.rc_ss_x1(self._h, int(k))

    def fillUpStateSpace(self, ids):
        """StateSpace.hs:145-160."""
        ids = _i32(ids)
        out = np.zeros(self._jsNrStates, dtype=np.int32)
        n = self._lib.rc_ss_fill_up(self._h, _ip(ids), len(ids), _ip(out), len(out))
        return out[:n]


makeJointStateSpace = JointStateSpace


# ---- the engine ------------------------------------------------------------------------------------------
@dataclass
class SpectrumEntry:
    """MaxUtils.hs:40-46."""
    spPattern: List[int]
    spCount: int
    spFreq: float
    spJackknifeError: Optional[float]
    spTheory: float


class Engine:
    """One rc_ctx: bound to one GPU (`Engine(device)`), or driving several GPUs of this process (`Engine(devices=[0, 1, ..])`
    / `Engine(nrGpus=N)`, rc_create_multi).  Holds the histogram-side problem (patterns, nVec, counts, total sites) and
    evaluates ModelSpecs against it."""

    def __init__(self, device=0, devices=None, nrGpus=None):
        self._lib = L.load()
        h = C.c_void_p()
        if nrGpus is not None and devices is None and nrGpus > 1:
            devices = list(range(int(nrGpus)))
        if devices is not None and len(devices) > 1:
            dv = _i32(devices)
            rc = self._lib.rc_create_multi(len(dv), _ip(dv), C.byref(h))
            device = int(dv[0])
        else:
            if devices is not None and len(devices) == 1:
                device = int(devices[0])
            rc = self._lib.rc_create(int(device), C.byref(h))
        if rc != L.RC_OK:
            raise RarecoalDeviceError(
                f"rc_create(device={device}, devices={devices}) failed with {rc}: no usable CUDA device. "
                "rarecoal_b200 has no CPU fallback.")
        self._h = h
        self.device = device
        self.nPat = 0
        self.P = 0
        self._times = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.rc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != L.RC_OK:
            _raise(rc, self._lib.rc_last_error(self._h).decode())

    # -- problem ------------------------------------------------------------------------------------
    def set_problem(self, nVec, maxAf, patterns, counts=None, totalSites=0):
        pats = _i32(patterns).reshape(-1, len(nVec))
        nv = _i32(nVec)
        cnt = np.zeros(len(pats), dtype=np.int64) if counts is None else np.ascontiguousarray(counts, dtype=np.int64)
        self._check(self._lib.rc_set_problem(self._h, len(nv), int(maxAf), len(pats), _ip(pats), _ip(nv),
                                             cnt.ctypes.data_as(C.POINTER(C.c_int64)), int(totalSites)))
        self.P, self.maxAf, self.nPat = len(nv), int(maxAf), len(pats)
        self.patterns, self.counts, self.totalSites, self.nVec = pats, cnt, int(totalSites), nv

    @property
    def nrDevices(self):
        return self._lib.rc_device_count(self._h)

    # -- one process per GPU: NCCL communicator inside the library --------------------------------------
    @staticmethod
    def comm_unique_id():
        """128-byte NCCL unique id (rank 0 creates it, the launcher hands it to every rank)."""
        buf = C.create_string_buffer(128)
        rc = L.load().rc_comm_unique_id(buf)
        if rc != L.RC_OK:
            raise RarecoalDeviceError("rc_comm_unique_id failed: NCCL (libnccl.so.2) is not available")
        return buf.raw

    def comm_init(self, nRanks, rank, unique_id):
        """Joins the communicator and fixes this context's pattern shard; eval_* then return the all-reduced result."""
        assert len(unique_id) == 128
        self._check(self._lib.rc_comm_init(self._h, int(nRanks), int(rank), C.c_char_p(unique_id)))

    def comm_allreduce(self, dPtr, n, stream=None):
        self._check(self._lib.rc_comm_allreduce(self._h, C.c_void_p(dPtr), int(n), C.c_void_p(stream) if stream else None))

    def set_shard(self, rank, world):
        self._check(self._lib.rc_set_shard(self._h, int(rank), int(world)))

    def set_times(self, times):
        t = _f64(times)
        self._check(self._lib.rc_set_times(self._h, len(t), _dp(t)))
        self._times = t

    # -- evaluation ---------------------------------------------------------------------------------
    def _pack(self, models):
        P = self.P
        evs, off, theta, disc = [], [0], [], []
        for m in models:
            if m.mNrPops != P or len(m.mDiscoveryRates) != P:
                raise RarecoalCompException("illegal sample configuration given")   # Core.hs:42
            evs.extend(m.mEvents)
            off.append(len(evs))
            theta.append(m.mTheta)
            disc.extend(m.mDiscoveryRates)
        return _events_array(evs), _i32(off), _f64(theta), _f64(disc)

    def _ensure_times(self, models):
        t = _f64(models[0].mTimeSteps)
        if self._times is None or len(t) != len(self._times) or not np.array_equal(t, self._times):
            self.set_times(t)

    def eval_models(self, models: Sequence[ModelSpec], siteRed=1.0, want_probs=False):
        """Batch of ModelSpecs (same grid) -> (logl[B], probs[B][nPat] or None, status[B]).  mNoShortcut is passed per
        model (rc_eval_models)."""
        B = len(models)
        self._ensure_times(models)
        ev, off, theta, disc = self._pack(models)
        ns = _i32([1 if m.mNoShortcut else 0 for m in models])
        logl = np.zeros(B)
        status = np.zeros(B, dtype=np.int32)
        probs = np.zeros((B, self.nPat)) if want_probs else None
        rc = self._lib.rc_eval_models(self._h, B, ev, _ip(off), _dp(theta), _dp(disc), _ip(ns), float(siteRed),
                                      _dp(probs) if want_probs else None, _dp(logl), _ip(status))
        self._check(rc)
        return logl, probs, status

    def eval_raw(self, ev, off, theta, disc, noShortcut=False, siteRed=1.0, want_probs=False):
        """Same as eval_models on pre-marshalled arrays (used by bench.py to keep Python out of the timing)."""
        B = len(theta)
        logl = np.zeros(B)
        status = np.zeros(B, dtype=np.int32)
        probs = np.zeros((B, self.nPat)) if want_probs else None
        rc = self._lib.rc_eval(self._h, B, ev, _ip(off), _dp(theta), _dp(disc), int(noShortcut), float(siteRed),
                               _dp(probs) if want_probs else None, _dp(logl), _ip(status))
        self._check(rc)
        return logl, probs, status

    def eval_partial(self, ev, off, theta, disc, noShortcut, dPartialPtr, dProbsPtr=None, stream=None):
        B = len(theta)
        status = np.zeros(B, dtype=np.int32)
        rc = self._lib.rc_eval_partial(self._h, B, ev, _ip(off), _dp(theta), _dp(disc), int(noShortcut),
                                       C.c_void_p(dPartialPtr), C.c_void_p(dProbsPtr) if dProbsPtr else None,
                                       C.c_void_p(stream) if stream else None, _ip(status))
        self._check(rc)
        return status

    def finalize(self, B, dPartialPtr, siteRed=1.0, stream=None):
        logl = np.zeros(B)
        self._check(self._lib.rc_finalize(self._h, B, C.c_void_p(dPartialPtr), float(siteRed),
                                          C.c_void_p(stream) if stream else None, _dp(logl)))
        return logl

    def set_option(self, name, value):
        """rc_set_option: "keyed", "ktab_model_mb", "ktab_total_mb"."""
        self._check(self._lib.rc_set_option(self._h, name.encode(), float(value)))

    def stats(self):
        st = L.rc_stats()
        self._check(self._lib.rc_last_stats(self._h, C.byref(st)))
        return {f: getattr(st, f) for f, _ in L.rc_stats._fields_}

    def fp64_peak_tflops(self):
        out = C.c_double(0.0)
        self._check(self._lib.rc_fp64_peak(self._h, C.byref(out)))
        return out.value

    # -- reference-named entry points -------------------------------------------------------------------
    def getProb(self, modelSpec: ModelSpec, jointStateSpace, nVec, config):
        """Core.hs:38-56.  `jointStateSpace` supplies maxAf as in the reference signature."""
        if not (len(nVec) == len(config) == modelSpec.mNrPops):
            raise RarecoalCompException("illegal sample configuration given")
        self.set_times(modelSpec.mTimeSteps) if self._times is None or len(self._times) != len(modelSpec.mTimeSteps) \
            or not np.array_equal(self._times, _f64(modelSpec.mTimeSteps)) else None
        max_af = jointStateSpace._jsMaxAf if hasattr(jointStateSpace, "_jsMaxAf") else int(jointStateSpace)
        out = C.c_double(0.0)
        ev = _events_array(modelSpec.mEvents)
        rc = self._lib.rc_get_prob(self._h, modelSpec.mNrPops, max_af, _ip(_i32(nVec)), _ip(_i32(config)), ev,
                                   len(modelSpec.mEvents), modelSpec.mTheta, _dp(_f64(modelSpec.mDiscoveryRates)),
                                   int(modelSpec.mNoShortcut), C.byref(out))
        self._check(rc)
        return out.value

    def computeFrequencySpectrum(self, modelSpec: ModelSpec):
        """MaxUtils.hs:93-119 on the problem set with set_problem (patterns already in model-branch order)."""
        logl, probs, status = self.eval_models([modelSpec], want_probs=True)
        if status[0]:
            _raise(int(status[0]), self._lib.rc_last_error(self._h).decode())
        tot = float(self.totalSites) if self.totalSites else float("nan")
        return [SpectrumEntry([int(v) for v in p], int(c), c / tot, None, float(t))
                for p, c, t in zip(self.patterns, self.counts, probs[0])]

    def computeLogLikelihood(self, modelSpec: ModelSpec, siteRed=1.0):
        """MaxUtils.hs:78-83."""
        logl, _, status = self.eval_models([modelSpec], siteRed=siteRed)
        if status[0]:
            _raise(int(status[0]), self._lib.rc_last_error(self._h).decode())
        return float(logl[0])


def computeLogLikelihoodFromSpec(totalNrSites, siteRed, spectrum):
    """MaxUtils.hs:85-91 on an already computed spectrum (host arithmetic on <= nPat numbers; the engine
    computes the same reduction on the device inside rc_eval)."""
    import math
    ll = 0.0
    sp = 0.0
    sc = 0
    for e in spectrum:
        ll = ll + math.log(e.spTheory) * float(e.spCount)
        sp = sp + e.spTheory
        sc += e.spCount
    return siteRed * (ll + float(totalNrSites - sc) * math.log(1.0 - sp))
