"""The N>1 path on CPU: two `gloo` ranks each take the pattern shard rc_set_shard would give them, compute their
partial sums (sum cnt_i log p_i, sum p_i) with the CPU oracle standing in for the kernels, all-reduce them exactly as
bench.py / the multi-GPU host code does, and must reproduce the single-rank log-likelihood."""
import json
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import rarecoal_b200 as R
from oracle import oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    f = json.load(open(os.path.join(GOLDEN, "fourpop.json")))
    times = R.defaultTimes()
    std = R.computeStandardOrder(f["nVec"], 4)
    d = {tuple(p): c for p, c in zip(f["patterns"], f["counts"])}
    cnt = np.array([d.get(tuple(int(v) for v in p), 0) for p in std], dtype=np.int64)
    mine = R.shardPatterns(std, rank, world)
    _, probs, _ = O.logl(4, 4, times, f["theta"], [1.0] * 4, False, f["events"], f["nVec"], std[mine], cnt[mine], 0)
    partial = torch.tensor([float(np.sum(np.log(probs) * cnt[mine])), float(probs.sum())], dtype=torch.float64)
    owned = torch.zeros(len(std), dtype=torch.float64)
    owned[torch.from_numpy(mine.astype(np.int64))] = 1.0
    dist.all_reduce(partial)          # the only data-path collective: 2 doubles per model
    dist.all_reduce(owned)
    other = f["totalSites"] - int(cnt.sum())
    logl = partial[0].item() + other * np.log(1.0 - partial[1].item())
    if rank == 0:
        json.dump({"logl": logl, "owned_min": owned.min().item(), "owned_max": owned.max().item(), "n0": len(mine)},
                  open(out_path, "w"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_partials(tmp_path):
    out = str(tmp_path / "r0.json")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = json.load(open(out))
    f = json.load(open(os.path.join(GOLDEN, "fourpop.json")))
    assert res["owned_min"] == 1.0 and res["owned_max"] == 1.0          # disjoint cover
    assert abs(-res["logl"] - f["score_m4_maxl"]) / f["score_m4_maxl"] < 1e-12
    assert 30 <= res["n0"] <= 39                                          # 69 patterns dealt evenly


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shards_are_a_balanced_disjoint_cover(world):
    nvec = [100] * 8
    pats = R.computeStandardOrder(nvec, 5)
    seen = np.zeros(len(pats), dtype=int)
    work = np.prod(np.maximum(pats, 1), axis=1)
    loads = []
    for r in range(world):
        mine = R.shardPatterns(pats, r, world)
        assert np.all(np.diff(mine) > 0)
        seen[mine] += 1
        loads.append(work[mine].sum())
    assert np.all(seen == 1)
    assert max(loads) <= 1.02 * min(loads) + 40          # work-sorted round-robin keeps ranks even
