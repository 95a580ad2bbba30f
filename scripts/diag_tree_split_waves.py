"""Dev helper: the two waves of dist.solve_tree_split for each rank of an emulated world, timed one after the other on ONE
GPU (no NCCL): where do the 48 / 100 ms of the two-device solve go?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hybrid_vehicle_platoon_b200 as hvp
from hybrid_vehicle_platoon_b200 import dist as D
from hybrid_vehicle_platoon_b200 import synth_mpc as G
dev = torch.device("cuda", 0)
ctx = hvp.Context(0)
world = int(os.environ.get("EMU_WORLD", 2))
for n, N in ((8, 6), (10, 6)):
    x0, params = G.cent_cases(np.random.default_rng(5), 4, n, N, stress=False)
    mpc = hvp.api.CompiledMpc(G.CENT, N, n_local=n, ctx=ctx)
    tx0 = torch.as_tensor(x0, device=dev); tp = torch.as_tensor(params, device=dev)
    tm = torch.full((4, n), 800.0, dtype=torch.float64, device=dev)
    depth = n + 2 + 2 * int(np.ceil(np.log2(world)))

    def timed(fn):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = fn(); b.record(); torch.cuda.synchronize()
        return out, a.elapsed_time(b)

    for rep in range(2):
        line = []
        bounds = []
        for r in range(world):
            oa, ta = timed(lambda: D.shard_wave(mpc, tx0, tm, tp, r, world, groups=16, prefix_depth=depth, node_budget=8, incumbent=None))
            bounds.append(oa["obj"]); line.append(f"A[{r}] {ta:.1f} ms")
        bound = torch.stack(bounds).min(0).values
        for r in range(world):
            ob, tb = timed(lambda: D.shard_wave(mpc, tx0, tm, tp, r, world, groups=256, prefix_depth=depth, node_budget=0, incumbent=bound))
            line.append(f"B[{r}] {tb:.1f} ms nodes {ob['nodes'].tolist()}")
        one, t1 = timed(lambda: D.shard_wave(mpc, tx0, tm, tp, 0, 1, groups=256, prefix_depth=0, node_budget=0, incumbent=None))
        if rep:
            print(f"n={n} world={world} depth={depth}: " + " | ".join(line) + f" || one device {t1:.1f} ms nodes {one['nodes'].tolist()}", flush=True)
