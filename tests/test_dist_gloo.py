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


class _StubMpc:
    """Stands in for api.CompiledMpc in the protocol test: records every shard call and answers with made-up results
    (rank r finds objective 10 - r in the probing wave and, given a bound, bound - 1 - r in the full wave)."""
    n_local, N, n_extra = 3, 4, 0

    def __init__(self):
        self.calls = []

    def solve_shard_device(self, B, x0, mass, params, rank, world, groups, prefix_depth, node_budget, incumbent, u, x, extra,
                           modes, obj, status, nodes, qp_iters, stream=None):
        self.calls.append(dict(rank=rank, world=world, groups=groups, prefix_depth=prefix_depth, node_budget=node_budget,
                               incumbent=None if incumbent is None else incumbent.clone()))
        val = (10.0 - rank) if incumbent is None else float(incumbent[0]) - 1.0 - rank
        obj.fill_(val); u.fill_(val); x.fill_(val); extra.zero_(); modes.fill_(rank); status.fill_(2)
        nodes.fill_(5); qp_iters.fill_(50)


def test_tree_split_protocol_waves_and_defaults(monkeypatch):
    """dist.solve_tree_split with a stub formulation (no GPU): one device = ONE full wave (its workers share the incumbent
    and adopt sub-trees, so there is nothing to probe); several devices = a probing wave with few workers and a node budget,
    allreduce(min) of the bound, the full wave from that bound -- and, when the caller leaves the prefix depth to the
    library, two more levels per doubling of the world on top of n + 2."""
    import types
    import torch
    from hybrid_vehicle_platoon_b200 import dist as D
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: types.SimpleNamespace(cuda_stream=0))
    B = 2
    x0 = torch.zeros((B, 3, 2), dtype=torch.float64); mass = torch.ones((B, 3), dtype=torch.float64)
    params = torch.zeros((B, 10), dtype=torch.float64)
    # one device
    cm = _StubMpc()
    out = D.solve_tree_split(cm, x0, mass, params, groups=256, rank=0, world=1, allreduce=lambda t, op: None)
    assert len(cm.calls) == 1
    assert cm.calls[0]["node_budget"] == 0 and cm.calls[0]["incumbent"] is None and cm.calls[0]["groups"] == 256
    assert cm.calls[0]["prefix_depth"] == 0                                   # library default (n + 2), chosen in pm_build.cu
    assert out["obj"].tolist() == [10.0, 10.0] and out["status"].tolist() == [2, 2] and out["nodes"].tolist() == [5, 5]
    # four devices, played by threads
    cms = [_StubMpc() for _ in range(4)]
    ranks = D.ThreadRanks(4)
    outs = ranks.run(lambda r, w, ar: D.solve_tree_split(cms[r], x0, mass, params, groups=256, wave_budget=8, rank=r, world=w,
                                                         allreduce=ar))
    for r, c in enumerate(cms):
        a, b = c.calls
        assert a["node_budget"] == 8 and a["groups"] == 16 and a["incumbent"] is None          # the probe
        assert b["node_budget"] == 0 and b["groups"] == 256
        assert b["incumbent"].tolist() == [7.0, 7.0]                                          # min over ranks of 10 - r
        assert a["prefix_depth"] == b["prefix_depth"] == 3 + 2 + 2 * 2
    for o in outs:                                                          # rank 3 wins the full wave: 7 - 1 - 3
        assert o["obj"].tolist() == [3.0, 3.0] and o["winner"].tolist() == [3, 3]
        assert o["nodes"].tolist() == [40, 40] and o["bound_after_wave_a"].tolist() == [7.0, 7.0]
