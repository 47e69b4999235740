"""GPU tests of the multi-GPU path that lives INSIDE the library (include/rarecoal_b200.h, "multi-GPU"):
rc_create_multi (one process, N devices) and rc_comm_init (one process per GPU, NCCL communicator in the library).
The reference's counterpart is the parMap fan-out inside computeFrequencySpectrum (MaxUtils.hs:108-110).

The model-sharded mode of rc_create_multi is exercised on ONE device as well (two per-device contexts on device 0: no
collective is involved); everything that needs NCCL needs two GPUs and is skipped otherwise (`gpurun --gpus 2`)."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import rarecoal_b200 as R
from rarecoal_b200 import workloads as Wk
from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _specs(wl, times, n, seed=11):
    evs = [wl.events()] + wl.proposal_batches(1, n - 1, seed=seed)[0]
    return [R.ModelSpec(wl.P, times, wl.theta, [1.0] * wl.P, 0.0, False, [R.ModelEvent.from_tuple(e) for e in ev]) for ev in evs]


def _single(wl, specs):
    nvec, pats, cnt = wl.problem()
    e = R.Engine(0)
    e.set_problem(nvec, wl.max_af, pats, cnt, wl.total_sites)
    out = e.eval_models(specs, want_probs=True)
    steps = e.stats()["state_steps"]
    e.close()
    return out, steps


def test_multi_context_models_sharded_on_one_device(default_times):
    # two per-device contexts behind one rc_ctx, both on device 0: the batch is cut into contiguous model ranges, every
    # range evaluated by its own context on its own host thread, results written straight into the caller's arrays
    wl = Wk.get("c2")
    specs = _specs(wl, default_times, 7)
    (logl, probs, status), steps = _single(wl, specs)
    nvec, pats, cnt = wl.problem()
    m = R.Engine(devices=[0, 0])
    assert m.nrDevices == 2
    m.set_option("multi_shard", 1)                # both contexts sit on ONE device here: NCCL would refuse the pattern split
    m.set_problem(nvec, wl.max_af, pats, cnt, wl.total_sites)
    logl2, probs2, status2 = m.eval_models(specs, want_probs=True)
    assert np.array_equal(logl2, logl) and np.array_equal(probs2, probs) and np.array_equal(status2, status)
    assert m.stats()["state_steps"] == steps
    # one failing model does not fail its range; an odd split (3 models over 2 contexts) works
    bad = R.ModelSpec(wl.P, default_times, wl.theta, [1.0] * wl.P, 0.0, False, [R.ModelEvent(-1.0, R.Join(0, 1))])
    l3, _, s3 = m.eval_models([specs[0], bad, specs[2]])
    assert s3.tolist() == [0, 1, 0] and l3[0] == logl[0] and np.isnan(l3[1]) and l3[2] == logl[2]
    # reference-named entry points work on the multi-device context too
    assert m.computeLogLikelihood(specs[0]) == logl[0]
    assert m.getProb(specs[0], R.JointStateSpace(wl.P, wl.max_af), nvec, pats[5]) == pytest.approx(probs[0][5], rel=1e-13)
    with pytest.raises(ValueError):
        m.set_shard(0, 2)                         # a multi-device context shards by itself
    m.close()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs (NCCL refuses two ranks on one device)")
@pytest.mark.parametrize("key,n_models", [("c2", 5), ("c3", 3)])
def test_multi_context_both_splits_match_one_gpu(default_times, key, n_models):
    # rc_create_multi over all GPUs of the box: models sharded and patterns sharded (ncclAllReduce inside the library)
    # must both reproduce the one-GPU answer; the oracle pins model 0
    wl = Wk.get(key)
    specs = _specs(wl, default_times, n_models)
    (logl, probs, status), steps = _single(wl, specs)
    nvec, pats, cnt = wl.problem()
    n = _n_gpus()
    m = R.Engine(nrGpus=n)
    m.set_problem(nvec, wl.max_af, pats, cnt, wl.total_sites)
    for mode in (1, 2, 0):
        m.set_option("multi_shard", mode)
        l2, p2, s2 = m.eval_models(specs, want_probs=True)
        assert not s2.any()
        assert m.stats()["state_steps"] == steps
        if mode == 1:
            assert np.array_equal(l2, logl) and np.array_equal(p2, probs)
        else:
            assert np.array_equal(p2, probs)                 # every pattern is computed by exactly one device
            assert np.max(np.abs(l2 - logl) / np.abs(logl)) < 1e-13
        l1, _, _ = m.eval_models(specs[:1])                   # B = 1 < N: auto = patterns sharded
        assert abs(l1[0] - logl[0]) / abs(logl[0]) < 1e-13
    ll_o, probs_o, _ = O.logl(wl.P, wl.max_af, default_times, wl.theta, [1.0] * wl.P, False, wl.events(), nvec, pats, cnt,
                              wl.total_sites)
    assert np.max(np.abs(p2[0] - probs_o) / probs_o) < 1e-10 and abs(l2[0] - ll_o) / abs(ll_o) < 1e-9
    m.close()


_RANK_SCRIPT = r"""
import json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import rarecoal_b200 as R
from rarecoal_b200 import workloads as Wk
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
wl = Wk.get(sys.argv[2])
nvec, pats, cnt = wl.problem()
times = R.defaultTimes()
evs = [wl.events()] + wl.proposal_batches(1, 2, seed=11)[0]
specs = [R.ModelSpec(wl.P, times, wl.theta, [1.0] * wl.P, 0.0, False, [R.ModelEvent.from_tuple(e) for e in ev]) for ev in evs]
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid = torch.from_numpy(np.frombuffer(R.Engine.comm_unique_id(), dtype=np.uint8).copy()).cuda()
dist.broadcast(uid, 0)                                   # plumbing only: the data-path collective is the library's
eng = R.Engine(rank)
eng.comm_init(world, rank, uid.cpu().numpy().tobytes())
eng.set_problem(nvec, wl.max_af, pats, cnt, wl.total_sites)
logl, probs, status = eng.eval_models(specs, want_probs=True)   # ONE C call per rank: shard, evaluate, ncclAllReduce, finalize
json.dump({"logl": logl.tolist(), "probs0": probs[0].tolist(), "status": status.tolist(), "steps": eng.stats()["state_steps"]},
          open(sys.argv[3] + f".{rank}", "w"))
eng.close()
dist.barrier()
dist.destroy_process_group()
"""


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("key", ["c2", "c3"])
def test_two_ranks_nccl_pattern_shards_match_one_gpu(tmp_path, default_times, key):
    # the north-star split on hardware: two processes, one GPU each, patterns sharded, partial sums all-reduced by NCCL
    # inside rc_eval — every rank must return the single-GPU log-likelihoods and probabilities
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "rank.py"
    script.write_text(_RANK_SCRIPT)
    out = str(tmp_path / "res")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), str(script), ROOT, key, out],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    wl = Wk.get(key)
    specs = _specs(wl, default_times, 3)
    (logl, probs, status), steps = _single(wl, specs)
    total_steps = 0
    for rank in (0, 1):
        res = json.load(open(f"{out}.{rank}"))
        assert res["status"] == [0, 0, 0]
        assert np.max(np.abs(np.array(res["logl"]) - logl) / np.abs(logl)) < 1e-13
        assert np.array_equal(np.array(res["probs0"]), probs[0])
        total_steps += res["steps"]
    assert total_steps == steps                    # the shards are a disjoint cover of the work
