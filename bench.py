#!/usr/bin/env python3
"""bench.py — log-likelihood evaluations/s on the C3 workload (BASELINE.json config[2]):
synthetic 8-population balanced tree, maxAf=10, 43 757 patterns, default 3043-point grid.

A "step" = one batch of B_total = batch * n_gpus parameter vectors (C4-style single-coordinate perturbations of
the C3 parameters; a fresh batch every step) evaluated against the resident histogram.  Weak scaling: per-GPU work
is fixed as N grows.  Default sharding is by MODEL (independent objects: every rank evaluates its own `batch` models
on all patterns, no data-path collective; the log-likelihoods are all-gathered in the e2e path).  `--shard patterns`
shards the patterns instead and all-reduces the two partial sums per model over NCCL.

  value  logl evals/s, parameter batch marshalled before the timed region, partials left on the device
  e2e    same metric through the C-ABI call with HOST buffers: per step the event arrays go host->device and
         the B log-likelihoods come back device->host
  roofline  propagation kernels (one launch per structural epoch), FP64 FMA pipe: achieved = state-steps * (4P+27)
         flops / CUDA-event time of those launches, peak = DFMA rate measured on this GPU by rc_fp64_peak
         (MEASURED_PEAKS.json has no FP64 figure); traffic = DRAM bytes of the same launches from the committed ncu capture
  cpu_baseline  the CPU oracle (port of the reference's algorithm; the Haskell reference cannot be built here)
         on a bounded pattern sample with all host cores, scaled to the full pattern set by state-steps

--impl reference times that CPU oracle alone (rank 0 only).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

P_C3, MAXAF_C3 = 8, 10
FLOPS_PER_STATE_STEP = 4 * P_C3 + 27
def _traffic():
    """dram read+write bytes of one rc_propagate_keyed_kernel launch from the committed ncu capture."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r01_keyed_traffic.json")))
    except Exception:
        return None


TRAFFIC = _traffic()


def _hbm_peak():
    """Measured copy bandwidth of this pool's B200s (driver-written MEASURED_PEAKS.json), else the profiling guide's fallback."""
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return 6539.2, "fallback (value measured on this pool earlier in the round)"


METRIC = "logl_evals_per_sec"
UNIT = "logl evals/s"
WORKLOAD = "C3: synthetic 8-pop balanced tree, maxAf=10, 43757 patterns, nVec=[100]*8, 3043-step grid"


def make_batches(n_batches, batch, seed=7):
    """Pre-generated parameter batches: each model = C3 parameters with one coordinate perturbed by a random
    factor in [0.99, 1.01] (never reorders events), as a maxl simplex / mcmc proposal batch would."""
    from rarecoal_b200 import workloads as Wk
    rng = np.random.default_rng(seed)
    x0 = Wk.c3_params()
    out = []
    for _ in range(n_batches):
        evs = []
        for _ in range(batch):
            x = x0.copy()
            j = rng.integers(0, len(x0))
            x[j] *= 1.0 + rng.uniform(-0.01, 0.01)
            evs.append(Wk.c3_events_from_params(x))
        out.append(evs)
    return out


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


def cpu_baseline_run(sample_every, threads=0):
    """Times the CPU oracle on every `sample_every`-th pattern of C3; returns (evals/s scaled to the full
    pattern set, cores, seconds, sample description)."""
    import rarecoal_b200 as R
    from rarecoal_b200 import workloads as Wk
    from oracle import oracle as O
    nvec, pats, cnt = Wk.c3_problem()
    times = O.time_steps()
    ev = Wk.c3_events()
    idx = np.arange(0, len(pats), sample_every)
    t0 = time.perf_counter()
    _, _, w = O.logl(P_C3, MAXAF_C3, times, 0.0005, [1.0] * P_C3, False, ev, nvec, pats[idx],
                     np.zeros(len(idx), dtype=np.int64), 0, threads=threads)
    dt = time.perf_counter() - t0
    w_full = R.planStats(nvec, MAXAF_C3, pats, times, ev)["state_steps"]
    full_seconds = dt * w_full / w
    cores = O.lib().orc_max_threads() if threads == 0 else threads
    return 1.0 / full_seconds, cores, dt, f"{len(idx)} of {len(pats)} patterns (every {sample_every}th), scaled by state-steps {w}/{w_full}"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    sample_every = 2
    for i in range(args.warmup + args.steps):
        v, cores, dt, sample = cpu_baseline_run(sample_every)
        if i >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = 1e3 / value
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": args.batch, "global_batch": args.batch * args.gpus,
                   "parallelism": "host thread pool over patterns, one model after the other (the rate does not depend on the batch); "
                                  "every step times every second pattern of one model and scales by executed state-steps"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample + "; the Haskell reference cannot be built in this image (no ghc), "
                                            "the timed code is the C++ oracle port with a thread pool over patterns"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--batch", type=int, default=44, help="models per GPU per step (44 = central differences, C4)")
    ap.add_argument("--shard", choices=["models", "patterns"], default="models",
                    help="N>1: models = every rank evaluates its own models on all patterns (no data-path collective, "
                         "log-likelihoods gathered at the end); patterns = every rank evaluates a pattern shard of all "
                         "models and the partial sums are all-reduced")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
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
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    nvec, pats, cnt = Wk.c3_problem()
    times = R.defaultTimes()
    eng = R.Engine(local)
    eng.set_problem(nvec, MAXAF_C3, pats, cnt, Wk.TOTAL_SITES)
    by_pattern = args.shard == "patterns" and world > 1
    if by_pattern:
        eng.set_shard(rank, world)
    eng.set_times(times)
    B = args.batch * world                     # models per step over all ranks (weak scaling)
    n_it = args.warmup + args.steps
    batches = make_batches(n_it, B)
    if not by_pattern:                         # this rank's slice of every batch
        batches = [evs[rank * args.batch:(rank + 1) * args.batch] for evs in batches]
    Bl = len(batches[0])                       # models this rank evaluates per step
    packed = []
    for evs in batches:
        flat = [e for m in evs for e in m]
        off = np.cumsum([0] + [len(m) for m in evs]).astype(np.int32)
        packed.append((R.api._events_array(flat), off, np.full(Bl, 0.0005), np.ones(Bl * P_C3)))
    h2d = world * (sum(len(m) for m in batches[0]) * ctypes.sizeof(R._lib.rc_event) + Bl * (4 + 8 + 8 * P_C3))
    d2h = 8 * B

    stream = torch.cuda.current_stream()
    partial = torch.zeros(2 * Bl, dtype=torch.float64, device="cuda")
    gathered = torch.zeros(B, dtype=torch.float64, device="cuda")

    def step_device(i):
        ev, off, th, dc = packed[i]
        eng.eval_partial(ev, off, th, dc, False, partial.data_ptr(), None, stream.cuda_stream)
        if by_pattern:
            dist.all_reduce(partial)

    def step_e2e(i):
        ev, off, th, dc = packed[i]
        if not by_pattern:
            logl, _, _ = eng.eval_raw(ev, off, th, dc, False, 1.0)       # host buffers in, host log-likelihoods out
            if world > 1:                                                # results of all ranks to every rank
                dist.all_gather_into_tensor(gathered, torch.from_numpy(logl).cuda())
            return logl
        eng.eval_partial(ev, off, th, dc, False, partial.data_ptr(), None, stream.cuda_stream)
        dist.all_reduce(partial)
        return eng.finalize(Bl, partial.data_ptr(), 1.0, stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fp64_peak = eng.fp64_peak_tflops()

    # ---- device-resident timing -------------------------------------------------------------------------------
    for i in range(args.warmup):
        step_device(i)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms, traj_ms, launches, w_steps = [], [], 0, 0
    barrier()
    e0.record(stream)
    for i in range(args.warmup, n_it):
        step_device(i)
    e1.record(stream)
    barrier()
    dev_ms = e0.elapsed_time(e1)
    # per-kernel CUDA-event times: re-run the timed batches one by one and read the library's own events
    for i in range(args.warmup, n_it):
        step_device(i)
        torch.cuda.synchronize()
        st = eng.stats()
        # span of the propagation launches on the main stream; the trajectory kernels run beside them on a side stream
        # (each epoch's launch waits for its own trajectory), so their time is not subtracted — the conservative reading
        kernel_ms.append(st["kernel_ms"])
        traj_ms.append(st["traj_ms"])
        launches += st["launches"] + (1 if by_pattern else 0)
        w_steps = st["state_steps"]
    # ---- end-to-end timing (host buffers in, host logl out) ---------------------------------------------------
    for i in range(args.warmup):
        step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    last = None
    for i in range(args.warmup, n_it):
        last = step_e2e(i)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    assert last is not None and np.all(np.isfinite(last))
    # single-model latency through the host-buffer call (what `rarecoal logl` would see), rank 0 / 1 GPU only
    lat_ms = None
    if world == 1:
        ev1 = R.api._events_array(batches[0][0])
        off1 = np.array([0, len(batches[0][0])], dtype=np.int32)
        for _ in range(3):
            eng.eval_raw(ev1, off1, np.full(1, 0.0005), np.ones(P_C3), False, 1.0)
        t1 = time.perf_counter()
        for _ in range(10):
            eng.eval_raw(ev1, off1, np.full(1, 0.0005), np.ones(P_C3), False, 1.0)
        lat_ms = (time.perf_counter() - t1) * 100.0

    t = torch.tensor([dev_ms, e2e_s * 1e3, float(np.mean(kernel_ms))], dtype=torch.float64, device="cuda")
    traj_mean = float(np.mean(traj_ms))
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, k_ms = (float(v) for v in t.cpu())
    w_total = torch.tensor([float(w_steps)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(w_total)
    if rank == 0:
        value = B * args.steps / (dev_ms * 1e-3)
        e2e_value = B * args.steps / (e2e_ms * 1e-3)
        achieved = w_steps * FLOPS_PER_STATE_STEP / (k_ms * 1e-3) / 1e12     # this rank's kernel
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": args.batch, "global_batch": B,
                       "parallelism": (f"patterns sharded over {world} ranks, NCCL all-reduce of 2*B doubles per step" if by_pattern
                                       else f"models sharded over {world} rank(s): {args.batch} models per rank and step on all patterns, "
                                            "no data-path collective (log-likelihoods all-gathered in the e2e path)"),
                       "l2": "fresh parameter batch each step; the trajectory tables (~60 MB per model, larger than L2 for the "
                             "batch) are rewritten by rc_traj_kernel every step",
                       "single_logl_latency_ms": lat_ms,
                       "pattern_probs_per_sec": value * len(pats),
                       "state_steps_per_sec": float(w_total.item()) * args.steps / (dev_ms * 1e-3)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp64_peak if fp64_peak else None, "traffic": TRAFFIC,
                         "kernel": (TRAFFIC or {}).get("kernels", "propagation launches (rc_propagate_pat_kernel / rc_propagate_rows_kernel)"),
                         "pipes": (TRAFFIC or {}).get("pipes_pct_of_peak_time_weighted"),
                         "frac_note": "achieved counts the reference's 4P+27 flops per state-step (SURVEY.md 8d); the factorised update "
                                      "executes fewer DP instructions, so frac can exceed 1 - the pipes measured by ncu are in `pipes`",
                         "kernel_ms": k_ms, "traj_kernels_ms_side_stream": traj_mean,
                         "algorithmic": f"{w_steps} state-steps x {FLOPS_PER_STATE_STEP} flops per launch (rank 0)",
                         "peak_source": "rc_fp64_peak DFMA microbenchmark on this GPU (no FP64 figure in MEASURED_PEAKS.json)",
                         "hbm_note": "algorithmic HBM bytes ~24 B x patterns x models per launch: not the bound",
                         "hbm_view": (lambda pk: {"dram_gbs_measured_traffic": (TRAFFIC or {}).get("dram_bytes_total", 0) / (k_ms * 1e-3) / 1e9,
                                                  "peak_gbs": pk[0], "peak_source": pk[1],
                                                  "frac": (TRAFFIC or {}).get("dram_bytes_total", 0) / (k_ms * 1e-3) / 1e9 / pk[0]})(_hbm_peak())},
            "clocks": sampler.summary(),
        }
        if not args.no_cpu_baseline:
            v, cores, dt, sample = cpu_baseline_run(1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                                    "seconds": dt}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()
