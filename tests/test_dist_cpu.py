"""CPU, world_size 2, gloo: the multi-GPU host logic (contiguous sample shards, one allgather of the placement
records, uneven / empty last shards).  The per-shard placements are produced by the oracle port here (there is
no CPU product path); what is under test is usher_b200/dist.py."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

import common
from usher_b200 import capi, dist as ud


def _worker(rank, world, port, path, n_take, q):
    import torch.distributed as dist
    from oracle import port as oport
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = common.load(path)
    s_ptr = g["s_ptr"][: n_take + 1]
    calls = g["calls"][: int(s_ptr[-1])]
    lo, hi, per, sp, sc = ud.shard_batch(s_ptr, calls, rank, world)
    local = np.zeros(hi - lo, capi.PLACEMENT_DTYPE)
    if hi > lo:
        pt = oport.PortTree(g["parent"], g["row_ptr"], g["muts"])
        o = pt.search(sp, sc, want_set=False)
        local["score"], local["best_node"], local["best_j"] = o["score"], o["best_dfs"], o["best_j"]
        local["num_best"], local["has_unique"] = o["num_best"], o["has_unique"]
        pt.close()
    full = ud.allgather_records(local, n_take, world)
    if rank == 0:
        q.put({k: full[k].tolist() for k in ("score", "best_node", "best_j", "num_best", "has_unique")})
    dist.destroy_process_group()


@pytest.mark.parametrize("n_take", [40, 37, 1])
def test_sharded_allgather_matches_single_batch(n_take):
    path = common.GOLDEN + "/random_05.npz"
    g = common.load(path)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + n_take
    procs = [ctx.Process(target=_worker, args=(r, 2, port, path, n_take, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert got["score"] == g["exp_score"][:n_take].tolist()
    assert got["best_node"] == g["exp_best_dfs"][:n_take].tolist()
    assert got["best_j"] == g["exp_best_j"][:n_take].tolist()
    assert got["num_best"] == g["exp_num_best"][:n_take].tolist()
    assert got["has_unique"] == g["exp_has_unique"][:n_take].tolist()


def test_shard_ranges_cover_the_batch():
    for n in (0, 1, 31, 32, 33, 1000, 10_000):
        for w in (1, 2, 4, 8):
            spans = [ud.shard_range(n, r, w)[:2] for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
