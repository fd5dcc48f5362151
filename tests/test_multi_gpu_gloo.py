"""N > 1 path on CPU: world_size-2 gloo processes shard the chains, all-gather the MC3 swap statistics
and reach identical swap decisions (the GPU evaluation itself needs no collective)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from mcmc_date_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B = 64
    a, b = sharding.shard_range(B, world, rank)
    rng = np.random.default_rng(0)
    stats_all = rng.normal(size=(B, 2)) * 10.0          # stands in for (ln prior, ln lik) of every chain
    local = torch.from_numpy(stats_all[a:b].copy())
    gathered = sharding.allgather_swap_stats(local, world, dist).numpy()
    # the device path's decision rule (mcd_mc3_swap, restated in tests/mh_ref.py): every rank updates its replica of the
    # slot table from the gathered statistics and the shared Philox counters
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import mh_ref
    C = 8
    slot, cos = np.arange(B) % C, np.arange(B)
    ladder = 1.0 / (1.0 + 0.4 * np.arange(C))
    for it in range(5):
        mh_ref.mc3_swap(gathered, slot, cos, ladder, ladder, C, -1, 99, it)
    swaps = slot.tolist()
    q.put((rank, a, b, gathered.tolist(), swaps))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_allgather_and_swaps():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, a0, b0, g0, s0), (r1, a1, b1, g1, s1) = res
    assert (a0, b0, a1, b1) == (0, 32, 32, 64)
    rng = np.random.default_rng(0)
    full = rng.normal(size=(64, 2)) * 10.0
    assert np.array_equal(np.array(g0), full) and np.array_equal(np.array(g1), full)
    assert s0 == s1
    slots = np.array(s0)
    assert (np.sort(slots.reshape(-1, 8), axis=1) == np.arange(8)).all() and (slots != np.arange(64) % 8).any()


def test_shard_range_needs_an_even_split():
    sys.path.insert(0, ROOT)
    import pytest
    from mcmc_date_b200 import sharding
    assert [sharding.shard_range(64, 4, r) for r in range(4)] == [(0, 16), (16, 32), (32, 48), (48, 64)]
    with pytest.raises(ValueError):
        sharding.shard_range(65, 4, 0)
    with pytest.raises(ValueError):
        sharding.shard_range(64, 4, 4)
