"""CPU: the N > 1 host logic (scenario sharding, max-over-ranks timing) on gloo, world size 2."""
import os
import socket

import pytest


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from hybrid_vehicle_platoon_b200.dist import max_over_ranks, shard_range, sum_over_ranks
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(4097, rank, world)
    t = max_over_ranks([1.0 + rank, 5.0 - rank])
    n = sum_over_ranks([hi - lo])
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, lo, hi, t, n))


def test_sharding_and_reductions_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    out = sorted(q.get(timeout=120) for _ in ps)
    [p.join(timeout=60) for p in ps]
    (r0, lo0, hi0, t0, n0), (r1, lo1, hi1, t1, n1) = out
    assert (lo0, hi0, lo1, hi1) == (0, 2049, 2049, 4097)
    assert t0 == t1 == [2.0, 5.0]
    assert n0 == n1 == [4097.0]


def test_shard_range_properties():
    from hybrid_vehicle_platoon_b200.dist import shard_range
    for total in (0, 1, 7, 4096, 4097):
        for world in (1, 2, 4, 8):
            blocks = [shard_range(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_balanced_shards_partition_and_balance():
    import numpy as np
    from hybrid_vehicle_platoon_b200.dist import balanced_shards
    rng = np.random.default_rng(3)
    costs = [7 * int(rng.integers(5, 16)) * int(rng.integers(4, 11)) for _ in range(4096)]
    for world in (1, 2, 4, 8):
        sh = balanced_shards(costs, world)
        assert sorted(i for s_ in sh for i in s_) == list(range(4096))
        loads = [sum(costs[i] for i in s_) for s_ in sh]
        assert max(loads) - min(loads) <= max(costs)


def _winner_worker(rank, world, port, q):
    """The exchange step of the multi-GPU tree split (dist.reduce_winner) on gloo with made-up per-rank results."""
    import torch
    import torch.distributed as dist
    from hybrid_vehicle_platoon_b200 import dist as D
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    inf = float("inf")
    # problem 0: rank 1 wins; 1: tie -> lowest rank (0); 2: nobody has a leaf; 3: rank 0 wins
    objs = [[5.0, 2.0, inf, 1.0], [3.0, 2.0, inf, inf]][rank]
    B = 4
    mine = dict(obj=torch.tensor(objs, dtype=torch.float64),
                u=torch.full((B, 1, 2), float(rank + 1), dtype=torch.float64),
                x=torch.full((B, 1, 2, 3), float(10 * (rank + 1)), dtype=torch.float64),
                extra=torch.zeros((B, 1), dtype=torch.float64),
                modes=torch.full((B, 1, 2), rank + 3, dtype=torch.int32),
                nodes=torch.full((B,), 7 + rank, dtype=torch.int32), qp_iters=torch.full((B,), 20, dtype=torch.int32),
                numeric=torch.zeros(B, dtype=torch.int32))
    out = D.reduce_winner(mine, rank, world, D.dist_allreduce)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, {k: v.tolist() for k, v in out.items()}))


def test_tree_split_exchange_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_winner_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    out = dict(q.get(timeout=120) for _ in ps)
    [p.join(timeout=60) for p in ps]
    assert out[0] == out[1]                                   # every rank ends with the same answer
    o = out[0]
    inf = float("inf")
    assert o["obj"] == [3.0, 2.0, inf, 1.0]
    assert o["winner"] == [1, 0, 2, 0]
    assert o["status"] == [2, 2, 3, 2]
    assert [u[0][0] for u in o["u"]] == [2.0, 1.0, 0.0, 1.0]
    assert [m[0][0] for m in o["modes"]] == [4, 3, -1, 3]
    assert o["nodes"] == [15] * 4
