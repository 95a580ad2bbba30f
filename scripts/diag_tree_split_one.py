"""Dev helper: ONE tree-split solve of the bench's n = 8, N = 6 problems (for an ncu launch list)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hybrid_vehicle_platoon_b200 as hvp
from hybrid_vehicle_platoon_b200 import dist as D
from hybrid_vehicle_platoon_b200 import synth_mpc as G
dev = torch.device("cuda", 0)
ctx = hvp.Context(0)
n, N, problems = 8, 6, 4
x0, params = G.cent_cases(np.random.default_rng(5), problems, n, N, stress=False)
mpc = hvp.api.CompiledMpc(G.CENT, N, n_local=n, ctx=ctx)
tx0 = torch.as_tensor(x0, device=dev); tp = torch.as_tensor(params, device=dev)
tm = torch.full((problems, n), 800.0, dtype=torch.float64, device=dev)
for rep in range(2):
    out = D.solve_tree_split(mpc, tx0, tm, tp, groups=256, prefix_depth=20, wave_budget=int(os.environ.get("WAVE", 16)))
    torch.cuda.synchronize()
print(out["nodes"].tolist(), out["obj"].tolist())
