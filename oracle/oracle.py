"""ctypes loader for the CPU oracle (oracle/rarecoal_oracle.cpp).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; the product package rarecoal_b200/ never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "librarecoal_oracle.so")

EV_JOIN, EV_SPLIT, EV_POPSIZE, EV_FREEZE = 0, 1, 2, 3


class Event(C.Structure):
    _fields_ = [("t", C.c_double), ("type", C.c_int32), ("k", C.c_int32), ("l", C.c_int32), ("v", C.c_double)]


def build(force=False):
    src = os.path.join(_HERE, "rarecoal_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        env = dict(os.environ)
        env.pop("CXX", None)
        subprocess.check_call(["make", "-C", _HERE, "CXX=g++"], env=env, stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        dp, ip, llp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_longlong)
        evp = C.POINTER(Event)
        L.orc_time_steps.argtypes = [C.c_int, C.c_int, C.c_double, dp, C.c_int]
        L.orc_all_configs.argtypes = [C.c_int, C.c_int, ip, ip, C.c_int]
        L.orc_choose.argtypes = [C.c_int, C.c_int]; L.orc_choose.restype = C.c_double
        L.orc_choose_cont.argtypes = [C.c_double, C.c_int]; L.orc_choose_cont.restype = C.c_double
        L.orc_ss_create.argtypes = [C.c_int, C.c_int]; L.orc_ss_create.restype = C.c_void_p
        L.orc_ss_destroy.argtypes = [C.c_void_p]
        L.orc_ss_nr_states.argtypes = [C.c_void_p]
        L.orc_ss_state_to_id.argtypes = [C.c_void_p, ip]
        L.orc_ss_id_to_state.argtypes = [C.c_void_p, C.c_int, ip]
        L.orc_ss_x1up.argtypes = [C.c_void_p, C.c_int, ip]
        L.orc_ss_x1.argtypes = [C.c_void_p, C.c_int]
        L.orc_ss_fill_up.argtypes = [C.c_void_p, ip, C.c_int, ip, C.c_int]
        L.orc_validate_model.argtypes = [C.c_int, dp, evp, C.c_int, C.c_char_p, C.c_int]
        L.orc_pop_join_a.argtypes = [dp, C.c_int, C.c_int]
        L.orc_pop_split_a.argtypes = [dp, C.c_int, C.c_int, C.c_double]
        L.orc_pop_join_b.argtypes = [C.c_void_p, dp, ip, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_pop_split_b.argtypes = [C.c_void_p, dp, ip, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]
        L.orc_get_prob.argtypes = [C.c_void_p, C.c_int, dp, C.c_double, dp, C.c_int, evp, C.c_int, ip, ip, dp,
                                   llp, C.c_char_p, C.c_int]
        L.orc_reg_penalty.argtypes = [C.c_int, C.c_double, evp, C.c_int]; L.orc_reg_penalty.restype = C.c_double
        L.orc_logl.argtypes = [C.c_int, C.c_int, C.c_int, dp, C.c_double, dp, C.c_int, evp, C.c_int, ip, C.c_int,
                               ip, llp, C.c_longlong, C.c_double, C.c_int, dp, dp, llp, C.c_char_p, C.c_int]
        _lib = L
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def make_events(events):
    """events: iterable of (t, type, k, l, v)."""
    arr = (Event * max(1, len(events)))()
    for i, (t, ty, k, l, v) in enumerate(events):
        arr[i] = Event(float(t), int(ty), int(k), int(l), float(v))
    return arr


def time_steps(n0=20000, lingen=400, tmax=20.0):
    n = lib().orc_time_steps(n0, lingen, tmax, None, 0)
    out = np.zeros(n)
    lib().orc_time_steps(n0, lingen, tmax, _dp(out), n)
    return out


def all_configs(max_af, nvec):
    nv = _i(nvec)
    n = lib().orc_all_configs(max_af, len(nv), _ip(nv), None, 0)
    out = np.zeros((n, len(nv)), dtype=np.int32)
    lib().orc_all_configs(max_af, len(nv), _ip(nv), _ip(out), n)
    return out


def choose(n, k):
    return lib().orc_choose(n, k)


def choose_cont(n, k):
    return lib().orc_choose_cont(n, k)


class StateSpace:
    """Mirror of JointStateSpace (StateSpace.hs:20-28)."""

    def __init__(self, nr_pop, max_af):
        self.P, self.max_af = nr_pop, max_af
        self.h = lib().orc_ss_create(nr_pop, max_af)
        self.nr_states = lib().orc_ss_nr_states(self.h)

    def __del__(self):
        try:
            lib().orc_ss_destroy(self.h)
        except Exception:
            pass

    def state_to_id(self, x):
        return lib().orc_ss_state_to_id(self.h, _ip(_i(x)))

    def id_to_state(self, i):
        out = np.zeros(self.P, dtype=np.int32)
        lib().orc_ss_id_to_state(self.h, i, _ip(out))
        return out

    def x1up(self, i):
        out = np.zeros(self.P, dtype=np.int32)
        lib().orc_ss_x1up(self.h, i, _ip(out))
        return out

    def x1(self, k):
        return lib().orc_ss_x1(self.h, k)

    def fill_up(self, ids):
        ids = _i(ids)
        out = np.zeros(self.nr_states, dtype=np.int32)
        n = lib().orc_ss_fill_up(self.h, _ip(ids), len(ids), _ip(out), len(out))
        return out[:n]

    def pop_join_b(self, b, live, k, l):
        b = _d(b).copy(); buf = np.zeros(self.nr_states, dtype=np.int32); buf[:len(live)] = live
        n = lib().orc_pop_join_b(self.h, _dp(b), _ip(buf), len(live), len(buf), k, l)
        return b, buf[:n]

    def pop_split_b(self, b, live, k, l, m):
        b = _d(b).copy(); buf = np.zeros(self.nr_states, dtype=np.int32); buf[:len(live)] = live
        n = lib().orc_pop_split_b(self.h, _dp(b), _ip(buf), len(live), len(buf), k, l, m)
        return b, buf[:n]


def pop_join_a(a, k, l):
    a = _d(a).copy(); lib().orc_pop_join_a(_dp(a), k, l); return a


def pop_split_a(a, k, l, m):
    a = _d(a).copy(); lib().orc_pop_split_a(_dp(a), k, l, m); return a


class OracleError(Exception):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code  # 1 = RarecoalModelException, 2 = RarecoalCompException


def validate_model(nr_pop, disc, events):
    err = C.create_string_buffer(256)
    rc = lib().orc_validate_model(nr_pop, _dp(_d(disc)), make_events(events), len(events), err, 256)
    if rc:
        raise OracleError(rc, err.value.decode())


def get_prob(ss, times, theta, disc, no_shortcut, events, nvec, config, want_steps=False):
    """getProb (Core.hs:38-56)."""
    t = _d(times); out = C.c_double(0.0); w = C.c_longlong(0); err = C.create_string_buffer(256)
    rc = lib().orc_get_prob(ss.h, len(t), _dp(t), theta, _dp(_d(disc)), int(no_shortcut), make_events(events),
                            len(events), _ip(_i(nvec)), _ip(_i(config)), C.byref(out), C.byref(w), err, 256)
    if rc:
        raise OracleError(rc, err.value.decode())
    return (out.value, w.value) if want_steps else out.value


def reg_penalty(nr_pop, reg, events):
    return lib().orc_reg_penalty(nr_pop, reg, make_events(events), len(events))


def logl(nr_pop, max_af, times, theta, disc, no_shortcut, events, nvec, patterns, counts, total_sites,
         site_red=1.0, threads=0):
    """computeFrequencySpectrum + computeLogLikelihoodFromSpec (MaxUtils.hs:85-119).
    Returns (logl, probs, state_steps)."""
    t = _d(times); pats = _i(patterns); cnt = np.ascontiguousarray(counts, dtype=np.int64)
    probs = np.zeros(len(pats)); ll = C.c_double(0.0); w = C.c_longlong(0); err = C.create_string_buffer(256)
    rc = lib().orc_logl(nr_pop, max_af, len(t), _dp(t), theta, _dp(_d(disc)), int(no_shortcut), make_events(events),
                        len(events), _ip(_i(nvec)), len(pats), _ip(pats),
                        cnt.ctypes.data_as(C.POINTER(C.c_longlong)), int(total_sites), site_red, threads,
                        _dp(probs), C.byref(ll), C.byref(w), err, 256)
    if rc:
        raise OracleError(rc, err.value.decode())
    return ll.value, probs, w.value
