"""Scenario sharding across GPUs: scenarios are independent, each rank owns a
contiguous block, there is no data-path collective; torch.distributed is used only for the
barrier, the max-over-ranks of device times and gathering small per-rank metrics."""
from __future__ import annotations


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of `total` scenarios owned by `rank` (sizes differ by <= 1)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(values, device=None):
    """Element-wise max of a small list of floats over all ranks (no-op without a process group)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def sum_over_ranks(values, device=None):
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(v) for v in t]


def balanced_shards(costs, world: int):
    """Load-balanced assignment for MIXED scenario sizes (n = 5..15, N = 4..10):
    sort by the cost proxy (7 n N binaries per scenario) and deal round-robin.  Returns one index list per rank."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    return [sorted(order[r::world]) for r in range(world)]


# ---- one LARGE MIQP tree split across GPUs (SURVEY.md 8e "Collective"; include/hvp.h hvp_mpc_solve_shard_dev) -------
# Protocol (every rank holds the same batch of problems):
#   wave A  each rank searches its share of every tree under a small node budget      -> a first incumbent per rank
#           allreduce(min) of the objectives (NCCL over NVLink; 8 bytes per problem)  -> the bound every rank starts from
#   wave B  each rank searches its share to completion, pruning against that bound
#           allreduce(min) of the objectives, allreduce(min) of the rank that attains it, and a masked allreduce(sum)
#           that moves the winner's solution to every rank.
# A rank's share = the sub-trees whose mode-prefix ordinal falls into its residue classes, so the shares partition the
# tree and the minimum over ranks is the global optimum -- the same argument as the in-device split of pm_kernel.cu.

ST_OPTIMAL, ST_INFEASIBLE, ST_NUMERIC = 2, 3, 12      # include/hvp.h status codes


def shard_wave(cm, x0, mass, params, rank, world, *, groups, prefix_depth, node_budget, incumbent):
    """One rank's share of every problem (torch CUDA tensors in, dict of torch tensors out)."""
    import torch
    B = x0.shape[0]
    nl, N, dev, f64, i32 = cm.n_local, cm.N, x0.device, torch.float64, torch.int32
    out = dict(u=torch.empty((B, nl, N), dtype=f64, device=dev), x=torch.empty((B, nl, 2, N + 1), dtype=f64, device=dev),
               extra=torch.empty((B, max(cm.n_extra, 1)), dtype=f64, device=dev),
               modes=torch.empty((B, nl, N), dtype=i32, device=dev), obj=torch.empty(B, dtype=f64, device=dev),
               status=torch.empty(B, dtype=i32, device=dev), nodes=torch.empty(B, dtype=i32, device=dev),
               qp_iters=torch.empty(B, dtype=i32, device=dev))
    cm.solve_shard_device(B, x0, mass, params, rank, world, groups, prefix_depth, node_budget, incumbent, out["u"],
                          out["x"], out["extra"], out["modes"], out["obj"], out["status"], out["nodes"], out["qp_iters"],
                          stream=torch.cuda.current_stream().cuda_stream)
    return out


def best_of_waves(a, b):
    """Per problem, the better of a rank's two wave results (b started from a bound <= a's objective)."""
    import torch
    pick = b["obj"] < a["obj"]
    sel = lambda k: torch.where(pick.view(-1, *([1] * (a[k].dim() - 1))), b[k], a[k])
    out = {k: sel(k) for k in ("u", "x", "extra", "modes", "obj")}
    out["nodes"] = a["nodes"] + b["nodes"]
    out["qp_iters"] = a["qp_iters"] + b["qp_iters"]
    out["numeric"] = ((a["status"] == ST_NUMERIC) | (b["status"] == ST_NUMERIC)).to(torch.int32)
    return out


def reduce_winner(mine, rank, world, allreduce):
    """Combine the ranks.  `allreduce(tensor, op)`, op in {"min", "sum", "max"}, reduces in place over the ranks.
    Returns the global optimum (objective; solution of the lowest rank attaining it) on every rank."""
    import torch
    best = mine["obj"].clone()
    allreduce(best, "min")                                      # THE incumbent exchange: 8 bytes per problem
    i32 = mine["nodes"]
    who = torch.where((mine["obj"] == best) & torch.isfinite(best), torch.full_like(i32, rank), torch.full_like(i32, world))
    allreduce(who, "min")
    won = who == rank
    out = {}
    for k in ("u", "x", "extra", "modes"):
        t = torch.where(won.view(-1, *([1] * (mine[k].dim() - 1))), mine[k], torch.zeros_like(mine[k]))
        allreduce(t, "sum")
        out[k] = t
    nodes = mine["nodes"].clone(); iters = mine["qp_iters"].clone(); numeric = mine["numeric"].clone()
    allreduce(nodes, "sum"); allreduce(iters, "sum"); allreduce(numeric, "max")
    feas = torch.isfinite(best)
    out["obj"] = best
    out["status"] = torch.where(numeric > 0, torch.full_like(i32, ST_NUMERIC),
                                torch.where(feas, torch.full_like(i32, ST_OPTIMAL), torch.full_like(i32, ST_INFEASIBLE)))
    out["modes"] = torch.where(feas.view(-1, 1, 1), out["modes"], torch.full_like(out["modes"], -1))
    out["nodes"], out["qp_iters"], out["winner"] = nodes, iters, who
    return out


def dist_allreduce(t, op):
    """In-place all-reduce over the default process group (NCCL on GPUs; a no-op for a single process)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op={"min": dist.ReduceOp.MIN, "sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX}[op])


def solve_tree_split(cm, x0, mass, params, *, groups=64, prefix_depth=0, wave_budget=8, probe_groups=16, rank=None,
                     world=None, allreduce=dist_allreduce):
    """Every rank calls this with the same problems (torch CUDA tensors on its own GPU): the tree of every MIQP is
    searched by all ranks together; the proven optimum comes back on every rank."""
    import torch.distributed as dist
    if rank is None:
        on = dist.is_available() and dist.is_initialized()
        rank, world = (dist.get_rank(), dist.get_world_size()) if on else (0, 1)
    if prefix_depth == 0 and world > 1:
        # The ranks own residue classes of the prefix ORDINAL, so there must be several prefixes per rank -- and the first
        # nl decisions (stage 0 of every vehicle) are all but fixed by the initial velocities.  Within a device waiting
        # workers adopt sub-trees, so one device takes the library default (nl + 2: a handful of prefixes); every doubling
        # of the world adds two levels.  Deeper is not free: prefix levels are enumerated, never solved, nothing prunes there.
        import math
        prefix_depth = int(cm.n_local + 2 + 2 * math.ceil(math.log2(world)))
    kw = dict(groups=groups, prefix_depth=prefix_depth)
    if world == 1:
        # one device: its workers share the incumbent through global memory and waiting workers adopt sub-trees of busy
        # ones (pm_kernel.cu), so the probing wave buys nothing -- r02r launch list: wave A 31 ms, wave B 24 ms
        b = shard_wave(cm, x0, mass, params, rank, world, node_budget=0, incumbent=None, **kw)
        b["numeric"] = (b["status"] == ST_NUMERIC).to(b["status"].dtype)
        out = reduce_winner(b, rank, world, allreduce)
        out["bound_after_wave_a"] = out["obj"].clone()
        return out
    # wave A is a probe (first dives under a node budget, no adoption): a few workers per problem are enough for a first
    # incumbent, and its length is pure latency in front of wave B (r02t: 256 workers x 16 nodes cost 20-50 ms at n = 8 / 10)
    a = shard_wave(cm, x0, mass, params, rank, world, node_budget=wave_budget, incumbent=None,
                   groups=min(groups, probe_groups), prefix_depth=prefix_depth)
    bound = a["obj"].clone()
    allreduce(bound, "min")
    b = shard_wave(cm, x0, mass, params, rank, world, node_budget=0, incumbent=bound, **kw)
    out = reduce_winner(best_of_waves(a, b), rank, world, allreduce)
    out["bound_after_wave_a"] = bound
    return out


class ThreadRanks:
    """`world` ranks played by threads of ONE process (tests, single-GPU checks of the split protocol): the
    all-reduce meets at a barrier and reduces the deposited tensors, like the collective does across processes."""

    def __init__(self, world):
        import threading
        self.world, self.slots, self.bar = world, [None] * world, threading.Barrier(world)

    def allreduce(self, rank):
        import torch

        def f(t, op):
            self.slots[rank] = t
            self.bar.wait()
            st = torch.stack(list(self.slots))
            red = st.min(dim=0).values if op == "min" else (st.max(dim=0).values if op == "max" else st.sum(dim=0))
            self.bar.wait()
            t.copy_(red)
            self.bar.wait()
        return f

    def run(self, fn):
        """fn(rank, world, allreduce) on every rank; returns the list of results."""
        import threading
        res, err = [None] * self.world, []

        def body(r):
            try:
                res[r] = fn(r, self.world, self.allreduce(r))
            except BaseException as e:          # noqa: BLE001 -- re-raised in the caller
                err.append(e)
                self.bar.abort()
        ts = [threading.Thread(target=body, args=(r,)) for r in range(self.world)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        if err:
            raise err[0]
        return res
