"""ctypes binding of librarecoal_b200.so (include/rarecoal_b200.h).  Fails loudly when the library is
missing or cannot be loaded: there is no Python / CPU fallback for the compute path."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.environ.get("RARECOAL_B200_LIB", os.path.join(HERE, "librarecoal_b200.so"))   # override: kernel A/B experiments

RC_OK, RC_ERR_MODEL, RC_ERR_COMP, RC_ERR_ARG, RC_ERR_CUDA, RC_ERR_UNSUPPORTED = 0, 1, 2, -1, -2, -3
RC_ERR_TEMPLATE, RC_ERR_HIST = 3, 4
RC_EV_JOIN, RC_EV_SPLIT, RC_EV_POPSIZE, RC_EV_FREEZE, RC_EV_GROWTH = 0, 1, 2, 3, 4


class rc_event(C.Structure):
    _fields_ = [("t", C.c_double), ("type", C.c_int32), ("k", C.c_int32), ("l", C.c_int32), ("v", C.c_double)]


class rc_stats(C.Structure):
    _fields_ = [("state_steps", C.c_int64), ("slots", C.c_int64), ("kernel_ms", C.c_double),
                ("traj_ms", C.c_double), ("tables_ms", C.c_double), ("total_ms", C.c_double), ("launches", C.c_int32),
                ("plan_cache_hit", C.c_int32), ("n_epochs", C.c_int32), ("n_steps", C.c_int32),
                ("fp64_instr", C.c_int64), ("plan_build_ms", C.c_double), ("plan_misses", C.c_int32),
                ("plan_evictions", C.c_int32), ("ktab_mb", C.c_double), ("chunks", C.c_int32),
                ("models_selfcontained", C.c_int32), ("graph_replay", C.c_int32), ("pad_", C.c_int32), ("host_ms", C.c_double)]


_dp, _ip, _lp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
_evp = C.POINTER(rc_event)
_vp = C.c_void_p

# name -> (restype, argtypes); one entry per symbol declared in include/rarecoal_b200.h
SIGNATURES = {
    "rc_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "rc_destroy": (None, [_vp]),
    "rc_last_error": (C.c_char_p, [_vp]),
    "rc_set_problem": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _ip, _ip, _lp, C.c_int64]),
    "rc_set_shard": (C.c_int, [_vp, C.c_int, C.c_int]),
    "rc_set_times": (C.c_int, [_vp, C.c_int, _dp]),
    "rc_time_steps": (C.c_int, [C.c_int, C.c_int, C.c_double, _dp, C.c_int]),
    "rc_eval": (C.c_int, [_vp, C.c_int, _evp, _ip, _dp, _dp, C.c_int, C.c_double, _dp, _dp, _ip]),
    "rc_eval_models": (C.c_int, [_vp, C.c_int, _evp, _ip, _dp, _dp, _ip, C.c_double, _dp, _dp, _ip]),
    "rc_eval_partial": (C.c_int, [_vp, C.c_int, _evp, _ip, _dp, _dp, C.c_int, _vp, _vp, _vp, _ip]),
    "rc_finalize": (C.c_int, [_vp, C.c_int, _vp, C.c_double, _vp, _dp]),
    "rc_create_multi": (C.c_int, [C.c_int, _ip, C.POINTER(_vp)]),
    "rc_device_count": (C.c_int, [_vp]),
    "rc_comm_unique_id": (C.c_int, [_vp]),
    "rc_comm_init": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "rc_comm_allreduce": (C.c_int, [_vp, _vp, C.c_int64, _vp]),
    "rc_get_prob": (C.c_int, [_vp, C.c_int, C.c_int, _ip, _ip, _evp, C.c_int, C.c_double, _dp, C.c_int, _dp]),
    "rc_set_option": (C.c_int, [_vp, C.c_char_p, C.c_double]),
    "rc_last_stats": (C.c_int, [_vp, C.POINTER(rc_stats)]),
    "rc_fp64_peak": (C.c_int, [_vp, _dp]),
    "rc_all_configs": (C.c_int, [C.c_int, C.c_int, _ip, _ip, C.c_int]),
    "rc_validate_model": (C.c_int, [C.c_int, _dp, _evp, C.c_int, C.c_char_p, C.c_int]),
    "rc_reg_penalty": (C.c_double, [C.c_int, C.c_double, _evp, C.c_int]),
    "rc_choose": (C.c_double, [C.c_int, C.c_int]),
    "rc_desugar_growth": (C.c_int, [C.c_int, _dp, _evp, C.c_int, _evp, C.c_int]),
    "rc_shard_patterns": (C.c_int, [C.c_int, C.c_int, _ip, C.c_int, C.c_int, _ip, C.c_int]),
    "rc_plan_stats": (C.c_int, [C.c_int, C.c_int, C.c_int, _ip, _ip, C.c_int, _dp, _evp, C.c_int, C.c_int,
                                _lp, _lp, _ip, _ip, _ip]),
    "rc_plan_last_checksum": (C.c_uint64, []),
    "rc_ss_create": (_vp, [C.c_int, C.c_int]),
    "rc_ss_destroy": (None, [_vp]),
    "rc_ss_nr_states": (C.c_int, [_vp]),
    "rc_ss_state_to_id": (C.c_int, [_vp, _ip]),
    "rc_ss_id_to_state": (C.c_int, [_vp, C.c_int, _ip]),
    "rc_ss_x1up": (C.c_int, [_vp, C.c_int, _ip]),
    "rc_ss_x1": (C.c_int, [_vp, C.c_int]),
    "rc_ss_fill_up": (C.c_int, [_vp, _ip, C.c_int, _ip, C.c_int]),
    "rc_template_to_json": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int]),
    "rc_template_instantiate": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_char_p), _dp, _evp, C.c_int, C.c_char_p, C.c_int]),
    "rc_template_instantiate_batch": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.c_int, _dp, C.c_double, _evp, C.c_int,
                                               C.POINTER(C.c_int32), C.POINTER(C.c_int32), _dp, C.c_char_p, C.c_int]),
    "rc_params_parse": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int, _dp, C.c_int]),
    "rc_hist_read": (_vp, [C.c_char_p, C.c_char_p, C.c_int]),
    "rc_hist_free": (None, [_vp]),
    "rc_hist_info": (C.c_int, [_vp, _ip, _ip, _ip, _ip, _lp, _ip, C.c_char_p, C.c_int]),
    "rc_hist_problem": (C.c_int, [_vp, C.c_int, C.c_int, _ip, C.c_int, _ip, C.c_int, C.c_int, C.POINTER(C.c_char_p), _ip, _lp,
                                  _ip, C.c_int, C.c_char_p, C.c_int]),
    "rc_version": (C.c_char_p, []),
}

_lib = None


def load():
    """Loads the shared library (building it is __graft_entry__.build()'s job, not an import side effect)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO):
        raise ImportError(
            f"{SO} is missing: build it with `python -m rarecoal_b200.build` (nvcc, sm_100a). "
            "rarecoal_b200 has no CPU fallback.")
    lib = C.CDLL(SO)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
