"""N>1 host logic on CPU: contiguous cell-balanced shards + the final gather over gloo (world_size 2).
The per-shard compute is stood in for by the oracle (this is a test of the sharding/gather plumbing)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle as orc
    from gonomics_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)  # same batch on every rank
    al = [rng.integers(0, 4, int(rng.integers(0, 120)), dtype=np.uint8) for _ in range(301)]
    be = [rng.integers(0, 4, int(rng.integers(0, 60)), dtype=np.uint8) for _ in range(301)]
    ao = np.concatenate([[0], np.cumsum([len(x) for x in al])]).astype(np.int64)
    bo = np.concatenate([[0], np.cumsum([len(x) for x in be])]).astype(np.int64)
    ac, bc = np.concatenate(al), np.concatenate(be)
    S = orc.HUMAN_CHIMP_TWO_SCORE_MATRIX
    bounds = shard.shard_bounds(ao, bo, world)
    lo, hi = bounds[rank]
    sa, sao, sb, sbo = shard.slice_batch(ac, ao, bc, bo, lo, hi)
    sc, off, cig = orc.batch(sa, sao, sb, sbo, S, -600, -150, 1, True, 1)
    g_sc, g_off, g_cig = shard.gather_results(sc, off, cig)
    w_sc, w_off, w_cig = orc.batch(ac, ao, bc, bo, S, -600, -150, 1, True, 1)
    ok = (np.array_equal(g_sc, w_sc) and np.array_equal(g_off, w_off)
          and np.array_equal(g_cig["run_length"], w_cig["run_length"]) and np.array_equal(g_cig["op"], w_cig["op"]))
    g2, _, _ = shard.gather_results(sc, None, None)
    ok = ok and np.array_equal(g2, w_sc)
    # the device-tensor form bench.py times on NCCL (here: CPU tensors over gloo)
    t_sc, t_off = torch.from_numpy(sc), torch.from_numpy(off)
    t_cig = torch.from_numpy(np.ascontiguousarray(cig).view(np.uint8).copy())
    a_sc, a_cnt, a_cig, metas, moved = shard.gather_device(t_sc, t_off, t_cig, cig.dtype.itemsize)
    d_sc, d_off, d_cig = shard.compact_gathered(a_sc, a_cnt, a_cig, metas, cig.dtype)
    ok = ok and np.array_equal(d_sc, w_sc) and np.array_equal(d_off, w_off) and moved > 0
    ok = ok and np.array_equal(d_cig["run_length"], w_cig["run_length"]) and np.array_equal(d_cig["op"], w_cig["op"])
    a_sc, a_cnt, _, metas, _ = shard.gather_device(t_sc, None, None)
    ok = ok and a_cnt is None and np.array_equal(shard.compact_gathered(a_sc, None, None, metas, cig.dtype)[0], w_sc)
    ret[rank] = (ok, bounds)
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret[r][0] for r in range(world))
    bounds = ret[0][1]
    assert bounds[0][0] == 0 and bounds[-1][1] == 301 and bounds[0][1] == bounds[1][0]


def test_shard_bounds_balance_and_edges():
    from gonomics_b200 import shard
    ao = np.arange(0, 1001 * 500, 500, dtype=np.int64)
    bo = np.arange(0, 1001 * 150, 150, dtype=np.int64)
    b = shard.shard_bounds(ao, bo, 8)
    assert [hi - lo for lo, hi in b] == [125] * 8
    b = shard.shard_bounds(np.zeros(1, dtype=np.int64), np.zeros(1, dtype=np.int64), 4)
    assert b == [(0, 0)] * 4
    ao = np.array([0, 10, 10, 10, 1000], dtype=np.int64)  # one heavy pair at the end
    bo = np.array([0, 10, 20, 30, 1000], dtype=np.int64)
    b = shard.shard_bounds(ao, bo, 2)
    assert b[0][0] == 0 and b[1][1] == 4 and b[0][1] == b[1][0]
