#!/usr/bin/env python3
"""bench.py — log-likelihood evaluations/s of the rarecoal hot path on B200 (BASELINE.json metric).

Headline workload (N = 1 and every N of the scaling run): C3 = BASELINE.json configs[2], synthetic 8-population balanced
tree, maxAf = 10, 43 757 patterns, default 3043-point grid, B = 44 parameter vectors per GPU and step (C4's central
differences: configs[3] is the same launch).  `--workload c1|c2|c5` times the other named configurations; the default
line carries them as `sub_records` (N = 1 only).

A "step" = one batch of `batch * n_gpus` parameter vectors (single-coordinate perturbations of the workload's
parameters, a fresh batch every step) evaluated against the resident histogram.  Weak scaling: per-GPU work is fixed.

  value     logl evals/s, parameter batch marshalled before the timed region, partial sums left on the device
  e2e       same metric through the C-ABI call with HOST buffers: per step the event arrays go host->device and the
            log-likelihoods come back device->host
  roofline  the propagation kernels (rc_propagate_pat_kernel & co, one launch per structural epoch and bin group), bound
            by the FP64 pipe: achieved = FP64 instructions the kernels execute (DMUL / DFMA, counted per shape by the
            planner, 2 flops per pipe slot) / CUDA-event span of those launches; peak = DFMA rate measured on this GPU by
            rc_fp64_peak (MEASURED_PEAKS.json has no FP64 figure); `algorithmic_speedup` keeps SURVEY.md 8d's figure
            (state-steps x (4P+27) flops of the reference's formulation / time / peak), which exceeds 1 because the
            factorised update executes fewer flops than the reference's; traffic = DRAM bytes from the committed ncu capture
  parity    model 0 of the first timed batch against the CPU oracle, all patterns (1e-10 / 1e-9)
  cpu_baseline  the same oracle run, timed: the C++ port of the reference's algorithm with a thread pool over patterns on
            all host cores (the Haskell reference cannot be built in this image)

N > 1 (torchrun, one process per GPU): the headline shards MODELS (no data-path collective); the same line also reports
the north-star split — PATTERNS sharded, partial sums all-reduced by NCCL inside the library (rc_comm_init) — and strong
scaling at a fixed global batch, under `multi_gpu`.

--impl reference times the CPU oracle alone (rank 0 only) and never loads librarecoal_b200.so.
"""
import argparse
import ctypes
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "logl_evals_per_sec"
UNIT = "logl evals/s"
DEFAULT_BATCH = {"c1": 64, "c2": 64, "c3": 44, "c5": 128}
PROB_RTOL, LOGL_RTOL = 1e-10, 1e-9


def _profile_json(name):
    try:
        return json.load(open(os.path.join(ROOT, "profiles", name)))
    except Exception:
        return None


def _hbm_peak():
    """Measured copy bandwidth of this pool's B200s (driver-written MEASURED_PEAKS.json), else the profiling guide's fallback."""
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return 6539.2, "fallback (value measured on this pool earlier)"


def load_workloads():
    """rarecoal_b200/workloads.py by file path: pure numpy, usable without importing the package (reference arm)."""
    spec = importlib.util.spec_from_file_location("rc_workloads", os.path.join(ROOT, "rarecoal_b200", "workloads.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def config_of(wl, batch, gpus):
    """Identical for both arms given the same command line."""
    return {"workload": wl.label, "patterns": None, "batch_per_gpu": batch, "global_batch": batch * gpus,
            "l2": "fresh parameter batch every step; the per-model trajectory tables rewritten every step exceed L2 for the batch"}


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False

    def _run_nvml(self):
        """NVML directly: a sample every 20 ms (nvidia-smi takes ~0.2 s per call, too coarse for a sub-second timed region)."""
        import pynvml as N
        N.nvmlInit()
        h = N.nvmlDeviceGetHandleByIndex(self.index)
        mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
        get_reasons = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = [("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4)]
        while not self.stop_flag:
            r = get_reasons(h)
            self.rows.append([str(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)), str(mx), str(N.nvmlDeviceGetPowerUsage(h) / 1000.0)] +
                             ["Active" if r & b else "Not Active" for _, b in bits])
            time.sleep(0.02)

    def run(self):
        try:
            self._run_nvml()
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                r = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if r.returncode == 0 and r.stdout.strip():
                    self.rows.append([c.strip() for c in r.stdout.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]),
                "power_w_max": max(float(r[2]) for r in self.rows), "samples": len(self.rows), "reasons": reasons}


# ---- the CPU oracle: checker and CPU baseline -------------------------------------------------------------------------
def oracle_eval(Wk, wl, events, problem, threads=0):
    """One full log-likelihood (every pattern) on the host cores.  SetGrowthRate events are handed over as the per-step
    SetPopSize list of the numpy statement of the sugar.  Returns (logl, probs, state_steps, seconds, cores)."""
    from oracle import oracle as O
    nvec, pats, cnt = problem
    times = O.time_steps()
    ev = Wk.desugar_growth_py(times, events)
    t0 = time.perf_counter()
    ll, probs, w = O.logl(wl.P, wl.max_af, times, wl.theta, [1.0] * wl.P, False, ev, nvec, pats, cnt, wl.total_sites,
                          threads=threads)
    dt = time.perf_counter() - t0
    cores = O.lib().orc_max_threads() if threads == 0 else threads
    return ll, probs, w, dt, cores


def run_reference(args):
    """The reference arm: the CPU port of the reference's algorithm on all host cores, all patterns, one model per step.
    Imports oracle/ and the numpy workload definitions only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    Wk = load_workloads()
    from oracle import oracle as O
    wl = Wk.get(args.workload)
    batch = args.batch or DEFAULT_BATCH[wl.key]
    problem = wl.problem(all_configs=O.all_configs)
    models = [m for b in wl.proposal_batches(args.warmup + args.steps, 1) for m in b]
    secs, cores = [], 0
    for i, ev in enumerate(models):
        ll, _, _, dt, cores = oracle_eval(Wk, wl, ev, problem)
        assert np.isfinite(ll)
        if i >= args.warmup:
            secs.append(dt)
    value = len(secs) / float(np.sum(secs))
    cfg = config_of(wl, batch, args.gpus)
    cfg["patterns"] = int(len(problem[1]))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"every step = ONE full log-likelihood (all {len(problem[1])} patterns) of one parameter vector of the "
                                   "workload's proposal stream; the rate does not depend on the batch (models are evaluated one "
                                   "after the other, thread pool over patterns ~ parMap, MaxUtils.hs:108-110). The Haskell "
                                   "reference cannot be built in this image (no ghc): the timed code is oracle/rarecoal_oracle.cpp"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ---- one GPU, one workload ----------------------------------------------------------------------------------------------
def pack_batches(R, batches, P, theta):
    packed = []
    for evs in batches:
        flat = [e for m in evs for e in m]
        off = np.cumsum([0] + [len(m) for m in evs]).astype(np.int32)
        packed.append((R.api._events_array(flat), off, np.full(len(evs), theta), np.ones(len(evs) * P)))
    return packed


def measure(R, Wk, wl, eng, batches, steps, warmup, stream, torch, sync_all, after_step=None, min_seconds=0.0):
    """Device-resident timing of `steps` batches after `warmup`, then the library's own per-kernel CUDA-event times of the
    same batches one by one.  Returns a dict of raw numbers (this rank)."""
    Bl = len(batches[0])
    packed = pack_batches(R, batches, wl.P, wl.theta)
    partial = torch.zeros(2 * Bl, dtype=torch.float64, device="cuda")

    def step(i):
        ev, off, th, dc = packed[i % len(packed)]
        eng.eval_partial(ev, off, th, dc, False, partial.data_ptr(), None, stream.cuda_stream)
        if after_step:
            after_step(partial)

    t_plan0 = time.perf_counter()
    step(0)
    torch.cuda.synchronize()
    first_call_s = time.perf_counter() - t_plan0
    st0 = eng.stats()
    for i in range(1, warmup):
        step(i)
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(warmup, warmup + steps):
        step(i)
    e1.record(stream)
    sync_all()
    dev_ms = e0.elapsed_time(e1)
    # a longer region of the same steps when the requested one is too short for clocks / power to settle
    long_ms, long_steps = None, 0
    if min_seconds > 0 and dev_ms * 1e-3 < min_seconds:
        long_steps = int(np.ceil(min_seconds / (dev_ms * 1e-3 / steps)))
        e0.record(stream)
        for i in range(long_steps):
            step(warmup + i)
        e1.record(stream)
        sync_all()
        long_ms = e0.elapsed_time(e1)
    kernel_ms, traj_ms, launches, w_steps, fp64_instr, misses = [], [], 0, 0, 0, 0
    host_ms, total_ms, tables_ms = [], [], []
    for i in range(warmup, warmup + steps):
        step(i)
        torch.cuda.synchronize()
        st = eng.stats()
        kernel_ms.append(st["kernel_ms"])
        traj_ms.append(st["traj_ms"])
        host_ms.append(st["host_ms"]); total_ms.append(st["total_ms"]); tables_ms.append(st["tables_ms"])
        launches += st["launches"]
        misses += st["plan_misses"]
        w_steps, fp64_instr = st["state_steps"], st["fp64_instr"]
    return {"dev_ms": dev_ms, "long_ms": long_ms, "long_steps": long_steps, "kernel_ms": float(np.mean(kernel_ms)),
            "traj_ms": float(np.mean(traj_ms)), "host_ms": float(np.mean(host_ms)), "gpu_total_ms": float(np.mean(total_ms)),
            "tables_ms": float(np.mean(tables_ms)), "ktab_mb": st["ktab_mb"], "launches": launches, "state_steps": w_steps, "fp64_instr": fp64_instr,
            "plan_build_s": st0["plan_build_ms"] * 1e-3, "first_call_s": first_call_s, "plan_misses_first_call": st0["plan_misses"],
            "plan_misses_timed": misses, "packed": packed}


def single_gpu_record(R, Wk, wl, device, batch, steps, warmup, torch, with_cpu=True, min_seconds=0.0, fp64_peak=None):
    """Everything the bench line says about one workload on one GPU."""
    nvec, pats, cnt = problem = wl.problem()
    eng = R.Engine(device)
    eng.set_problem(nvec, wl.max_af, pats, cnt, wl.total_sites)
    eng.set_times(R.defaultTimes())
    stream = torch.cuda.current_stream()
    n_it = warmup + steps
    batches = wl.proposal_batches(n_it, batch)
    sampler = ClockSampler(device)
    sampler.start()
    m = measure(R, Wk, wl, eng, batches, steps, warmup, stream, torch, torch.cuda.synchronize, min_seconds=min_seconds)
    packed = m.pop("packed")
    # ---- end to end: host buffers in, host log-likelihoods out -----------------------------------------------------
    for i in range(warmup):
        eng.eval_raw(*packed[i], False, 1.0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    last = None
    for i in range(warmup, n_it):
        last, _, status = eng.eval_raw(*packed[i], False, 1.0)
    e2e_s = time.perf_counter() - t0
    sampler.stop_flag = True
    sampler.join(timeout=2)
    assert last is not None and np.all(np.isfinite(last)) and not status.any()
    # ---- single-model latency through the host-buffer call (what `rarecoal logl` sees) -------------------------------
    ev1 = R.api._events_array(batches[0][0])
    off1 = np.array([0, len(batches[0][0])], dtype=np.int32)
    one = (ev1, off1, np.full(1, wl.theta), np.ones(wl.P))
    for _ in range(5):
        eng.eval_raw(*one, False, 1.0)
    t1 = time.perf_counter()
    for _ in range(20):
        eng.eval_raw(*one, False, 1.0)
    lat_ms = (time.perf_counter() - t1) * 50.0
    lat_launches = eng.stats()["launches"]
    # ---- parity of the timed batch: model 0 of the first timed batch vs the oracle, all patterns ------------------------
    ev0 = batches[warmup][0]
    logl_g, probs_g, _ = eng.eval_raw(R.api._events_array(ev0), np.array([0, len(ev0)], dtype=np.int32),
                                      np.full(1, wl.theta), np.ones(wl.P), False, 1.0, want_probs=True)
    rec_parity, rec_cpu = None, None
    if with_cpu:
        ll_o, probs_o, w_o, dt_o, cores = oracle_eval(Wk, wl, ev0, problem)
        nz = probs_o != 0
        max_rel = float(np.max(np.abs(probs_g[0][nz] - probs_o[nz]) / probs_o[nz])) if nz.any() else 0.0
        rel_ll = float(abs(logl_g[0] - ll_o) / abs(ll_o))
        ok = bool(max_rel < PROB_RTOL and rel_ll < LOGL_RTOL and np.array_equal(probs_g[0][~nz], probs_o[~nz]))
        rec_parity = {"max_rel_prob": max_rel, "rel_logl": rel_ll, "patterns": int(len(pats)), "tol_prob": PROB_RTOL,
                      "tol_logl": LOGL_RTOL, "ok": ok, "checked": "model 0 of the first timed batch vs oracle/, every pattern"}
        if not ok:
            raise SystemExit(f"bench.py: parity FAILED on {wl.key}: {rec_parity}")
        rec_cpu = {"value": 1.0 / dt_o, "unit": UNIT, "cores": cores, "kind": "port", "seconds": dt_o,
                   "sample": f"one full log-likelihood: all {len(pats)} patterns of that model ({w_o} state-steps), thread pool over patterns"}
    if fp64_peak is None:
        fp64_peak = eng.fp64_peak_tflops()
    eng.close()
    value = batch * steps / (m["dev_ms"] * 1e-3)
    k_s = m["kernel_ms"] * 1e-3
    executed = 2.0 * m["fp64_instr"] / k_s / 1e12
    algorithmic = m["state_steps"] * (4 * wl.P + 27) / k_s / 1e12
    traffic = _profile_json("r02_traffic.json") if wl.key == "c3" else None
    hbm_peak, hbm_src = _hbm_peak()
    rec = {
        "workload": wl.label, "patterns": int(len(pats)), "batch": batch, "value": value, "unit": UNIT,
        "ms_per_step": m["dev_ms"] / steps, "timed_region_s": m["dev_ms"] * 1e-3,
        "value_long_run": (batch * m["long_steps"] / (m["long_ms"] * 1e-3)) if m["long_ms"] else None,
        "long_run_s": (m["long_ms"] * 1e-3) if m["long_ms"] else None,
        "pattern_probs_per_sec": value * len(pats), "state_steps_per_sec": m["state_steps"] * steps / (m["dev_ms"] * 1e-3),
        "single_logl_latency_ms": lat_ms, "single_logl_launches": lat_launches,
        "e2e": {"value": batch * steps / e2e_s, "unit": UNIT,
                "h2d_bytes_per_step": int(sum(len(x) for x in batches[0]) * ctypes.sizeof(R._lib.rc_event) + batch * (4 + 8 + 8 * wl.P)),
                "d2h_bytes_per_step": 8 * batch},
        "gpu_launches": int(m["launches"]),
        "per_step_ms": {"host_resolve_pack_enqueue": m["host_ms"], "gpu_first_to_last_launch": m["gpu_total_ms"],
                        "tables_kernel": m["tables_ms"], "propagation_launches": m["kernel_ms"], "traj_kernels_side_stream": m["traj_ms"],
                        "trajectory_table_mb": m["ktab_mb"]},
        "plan": {"build_s_first_call": m["plan_build_s"], "first_call_s": m["first_call_s"],
                 "cache_misses_first_call": m["plan_misses_first_call"], "cache_misses_timed": m["plan_misses_timed"]},
        "roofline": {
            "bound": "fp64", "achieved": executed, "peak": fp64_peak, "unit": "TFLOP/s", "frac": executed / fp64_peak if fp64_peak else None,
            "traffic": traffic,
            "kernel": "propagation launches: rc_propagate_pat_kernel<12/25/36>, rc_propagate_rows_kernel, rc_propagate_keyed_kernel",
            "kernel_ms": m["kernel_ms"], "traj_kernels_ms_side_stream": m["traj_ms"],
            "executed": f"{m['fp64_instr']} FP64 instructions (DMUL/DFMA, counted per shape by the planner) x 2 flops per pipe slot, per launch chain of one step",
            "algorithmic": f"{m['state_steps']} state-steps x {4 * wl.P + 27} flops (SURVEY.md 8d, the reference's formulation) per step",
            "algorithmic_tflops": algorithmic, "algorithmic_speedup": algorithmic / fp64_peak if fp64_peak else None,
            "peak_source": "rc_fp64_peak DFMA microbenchmark on this GPU (no FP64 figure in MEASURED_PEAKS.json)",
            "hbm_note": "algorithmic HBM bytes ~24 B x patterns x models per step: not the bound",
            "hbm_view": None if not traffic else {
                "dram_gbs_measured_traffic": traffic.get("dram_bytes_total", 0) / k_s / 1e9, "peak_gbs": hbm_peak, "peak_source": hbm_src,
                "frac": traffic.get("dram_bytes_total", 0) / k_s / 1e9 / hbm_peak}},
        "clocks": sampler.summary(),
    }
    if rec_parity:
        rec["parity"] = rec_parity
    if rec_cpu:
        rec["cpu_baseline"] = rec_cpu
    return rec, fp64_peak


# ---- N GPUs (one process per GPU) -----------------------------------------------------------------------------------
def multi_gpu(R, Wk, wl, args, batch, torch, dist, rank, world, local):
    nvec, pats, cnt = wl.problem()
    stream = torch.cuda.current_stream()
    n_it = args.warmup + args.steps
    B = batch * world
    batches = wl.proposal_batches(n_it, B)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        return float(t.item())

    # ---- (1) models sharded: every rank evaluates its own `batch` models on all patterns, no data-path collective ----
    eng = R.Engine(local)
    eng.set_problem(nvec, wl.max_af, pats, cnt, wl.total_sites)
    eng.set_times(R.defaultTimes())
    mine = [evs[rank * batch:(rank + 1) * batch] for evs in batches]
    sampler = ClockSampler(local)
    sampler.start()
    m = measure(R, Wk, wl, eng, mine, args.steps, args.warmup, stream, torch, barrier)
    packed = m.pop("packed")
    gathered = torch.zeros(B, dtype=torch.float64, device="cuda")
    for i in range(args.warmup):
        eng.eval_raw(*packed[i], False, 1.0)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.warmup, n_it):
        logl, _, _ = eng.eval_raw(*packed[i], False, 1.0)                 # host buffers in, host log-likelihoods out
        dist.all_gather_into_tensor(gathered, torch.from_numpy(logl).cuda())   # results of all ranks to every rank
    torch.cuda.synchronize()
    e2e_ms = allmax((time.perf_counter() - t0) * 1e3)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    dev_ms = allmax(m["dev_ms"])
    k_ms = allmax(m["kernel_ms"])
    w_total = allsum(m["state_steps"])
    fp64_peak = eng.fp64_peak_tflops()
    ref_logl = gathered.cpu().numpy().copy()                              # last batch, every model, from the model-sharded path
    # ---- (2) strong scaling: ONE batch of `batch` models over all ranks (models sharded) -----------------------------
    strong = None
    if batch >= world:
        per = batch // world
        small = [evs[rank * per:(rank + 1) * per] for evs in wl.proposal_batches(n_it, per * world)]
        ms = measure(R, Wk, wl, eng, small, args.steps, args.warmup, stream, torch, barrier)
        strong = {"global_batch": per * world, "value": per * world * args.steps / (allmax(ms["dev_ms"]) * 1e-3),
                  "ms_per_step": allmax(ms["dev_ms"]) / args.steps, "mode": "models sharded"}
    eng.close()
    # ---- (3) the north-star split: patterns sharded, NCCL all-reduce of the 2*B partial sums inside the library --------
    engp = R.Engine(local)
    engp.set_problem(nvec, wl.max_af, pats, cnt, wl.total_sites)
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        uid = torch.from_numpy(np.frombuffer(R.Engine.comm_unique_id(), dtype=np.uint8).copy())
    uid = uid.cuda()
    dist.broadcast(uid, 0)
    engp.comm_init(world, rank, uid.cpu().numpy().tobytes())              # also shards the patterns (rc_set_shard)
    engp.set_times(R.defaultTimes())
    mp = measure(R, Wk, wl, engp, batches, args.steps, args.warmup, stream, torch, barrier,
                 after_step=lambda part: engp.comm_allreduce(part.data_ptr(), part.numel(), stream.cuda_stream))
    packed_p = mp.pop("packed")
    for i in range(args.warmup):
        engp.eval_raw(*packed_p[i], False, 1.0)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.warmup, n_it):
        logl_p, _, _ = engp.eval_raw(*packed_p[i], False, 1.0)            # ONE C call: shard, evaluate, ncclAllReduce, finalize
    torch.cuda.synchronize()
    e2e_p_ms = allmax((time.perf_counter() - t0) * 1e3)
    rel = float(np.max(np.abs(logl_p - ref_logl) / np.abs(ref_logl)))    # the two splits agree on the same batch
    if not rel < LOGL_RTOL:
        raise SystemExit(f"bench.py: pattern-sharded log-likelihoods differ from the model-sharded ones: {rel}")
    strong_p = None
    one = wl.proposal_batches(n_it, batch)
    msp = measure(R, Wk, wl, engp, one, args.steps, args.warmup, stream, torch, barrier,
                  after_step=lambda part: engp.comm_allreduce(part.data_ptr(), part.numel(), stream.cuda_stream))
    msp.pop("packed")
    strong_p = {"global_batch": batch, "value": batch * args.steps / (allmax(msp["dev_ms"]) * 1e-3),
                "ms_per_step": allmax(msp["dev_ms"]) / args.steps, "mode": "patterns sharded + ncclAllReduce"}
    dev_p_ms, k_p_ms, traj_p_ms = allmax(mp["dev_ms"]), allmax(mp["kernel_ms"]), allmax(mp["traj_ms"])
    engp.close()
    if rank != 0:
        return None
    value = B * args.steps / (dev_ms * 1e-3)
    value_p = B * args.steps / (dev_p_ms * 1e-3)
    executed = 2.0 * m["fp64_instr"] / (k_ms * 1e-3) / 1e12
    algorithmic = m["state_steps"] * (4 * wl.P + 27) / (k_ms * 1e-3) / 1e12
    cfg = config_of(wl, batch, world)
    cfg["patterns"] = int(len(pats))
    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": cfg,
        "parallelism": f"models sharded over {world} ranks: {batch} models per rank and step on all patterns, no data-path collective "
                       "(log-likelihoods all-gathered in the e2e path)",
        "e2e": {"value": B * args.steps / (e2e_ms * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": int(world * (sum(len(x) for x in mine[0]) * ctypes.sizeof(R._lib.rc_event) + batch * (4 + 8 + 8 * wl.P))),
                "d2h_bytes_per_step": 8 * B},
        "gpu_launches": int(m["launches"]),
        "pattern_probs_per_sec": value * len(pats), "state_steps_per_sec": w_total * args.steps / (dev_ms * 1e-3),
        "roofline": {"bound": "fp64", "achieved": executed, "peak": fp64_peak, "unit": "TFLOP/s", "frac": executed / fp64_peak,
                     "traffic": _profile_json("r02_traffic.json"), "kernel_ms": k_ms, "algorithmic_tflops": algorithmic,
                     "algorithmic_speedup": algorithmic / fp64_peak,
                     "kernel": "propagation launches of rank 0 (rc_propagate_pat_kernel<12/25/36>)",
                     "peak_source": "rc_fp64_peak DFMA microbenchmark on this GPU"},
        "multi_gpu": {
            "pattern_sharded": {"value": value_p, "vs_model_sharded": value_p / value, "ms_per_step": dev_p_ms / args.steps,
                                "kernel_ms": k_p_ms, "traj_ms": traj_p_ms, "e2e_value": B * args.steps / (e2e_p_ms * 1e-3),
                                "collective": "ncclAllReduce(sum, double, 2*B) issued by the library on its own stream (rc_comm_init)",
                                "max_rel_logl_vs_model_sharded": rel},
            "strong_scaling": [s for s in (strong, strong_p) if s]},
        "clocks": sampler.summary(),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", choices=["c1", "c2", "c3", "c5"], default="c3")
    ap.add_argument("--batch", type=int, default=0, help="models per GPU per step (default: 44 for c3 = central differences, C4)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="N = 1: skip the sub-records of the other named workloads")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 1)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import rarecoal_b200 as R
    from rarecoal_b200 import workloads as Wk

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; rarecoal_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    wl = Wk.get(args.workload)
    batch = args.batch or DEFAULT_BATCH[wl.key]
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        line = multi_gpu(R, Wk, wl, args, batch, torch, dist, rank, world, local)
        if rank == 0:
            print(json.dumps(line))
        dist.barrier()
        dist.destroy_process_group()
        return

    rec, peak = single_gpu_record(R, Wk, wl, local, batch, args.steps, args.warmup, torch, with_cpu=not args.no_cpu_baseline,
                                  min_seconds=1.5)
    cfg = config_of(wl, batch, 1)
    cfg["patterns"] = rec["patterns"]
    line = {"metric": METRIC, "value": rec["value"], "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg,
            "parallelism": f"one GPU: {batch} models per step in one launch chain on all patterns"}
    for k in ("e2e", "gpu_launches", "roofline", "cpu_baseline", "parity", "clocks", "plan", "timed_region_s", "value_long_run",
              "long_run_s", "per_step_ms", "pattern_probs_per_sec", "state_steps_per_sec", "single_logl_latency_ms", "single_logl_launches"):
        if k in rec:
            line[k] = rec[k]
    if not args.no_sub and args.workload == "c3":
        # the other named configurations, shorter runs: C2 = configs[1] (4 pops, maxAf 10, admixture: the reference's own
        # example), C5 = configs[4] (10 pops, admixture + growth, mcmc proposal batches), C1 = configs[0]
        subs = {}
        for key in ("c2", "c5", "c1"):
            sub, _ = single_gpu_record(R, Wk, Wk.get(key), local, DEFAULT_BATCH[key], max(5, min(args.steps, 20)), max(2, args.warmup),
                                       torch, with_cpu=not args.no_cpu_baseline, fp64_peak=peak)
            sub.pop("clocks", None)
            subs[key] = sub
        line["sub_records"] = subs
    print(json.dumps(line))


if __name__ == "__main__":
    main()
