"""Dev helper: what a shard launch costs before any search -- one worker per problem, node budgets 1, 2, 4, 16."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hybrid_vehicle_platoon_b200 as hvp
from hybrid_vehicle_platoon_b200 import dist as D
from hybrid_vehicle_platoon_b200 import synth_mpc as G
dev = torch.device("cuda", 0)
ctx = hvp.Context(0)
for n, N in ((8, 6), (10, 6)):
    x0, params = G.cent_cases(np.random.default_rng(5), 4, n, N, stress=False)
    mpc = hvp.api.CompiledMpc(G.CENT, N, n_local=n, ctx=ctx)
    tx0 = torch.as_tensor(x0, device=dev); tp = torch.as_tensor(params, device=dev)
    tm = torch.full((4, n), 800.0, dtype=torch.float64, device=dev)
    for groups, budget in ((1, 1), (1, 2), (1, 4), (1, 16), (16, 1), (256, 1)):
        for rep in range(2):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            o = D.shard_wave(mpc, tx0, tm, tp, 0, 1, groups=groups, prefix_depth=0, node_budget=budget, incumbent=None)
            b.record(); torch.cuda.synchronize()
        print(f"n={n} groups={groups} budget={budget}: {a.elapsed_time(b):.2f} ms nodes {o['nodes'].tolist()} qp_iters {o['qp_iters'].tolist()}", flush=True)
