"""The heaviest group of the mixed sweep (N = 10, time headway) alone, repeated: time and nodes per run."""
import os, sys, time, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hybrid_vehicle_platoon_b200 as hvp
from hybrid_vehicle_platoon_b200.misc import ConstantSpacingPolicy, ConstantTimePolicy, StopAndGoLeaderTrajectory, spacing_params
from hybrid_vehicle_platoon_b200.sweep import run_mixed_sweep
ctx = hvp.Context(0)
S, T = 4096, 10
rng = np.random.default_rng(1234 + 3)
scen = []
for _ in range(S):
    n, N = int(rng.integers(5, 16)), int(rng.integers(4, 11))
    v = np.floor(rng.uniform(8, 30, n)); gaps = rng.uniform(60, 160, n)
    p = np.floor(3000.0 - np.cumsum(gaps) + gaps[0])
    x0 = np.empty(2 * n); x0[0::2] = p; x0[1::2] = v
    lx = StopAndGoLeaderTrajectory(p=3000, vh=20, vl=float(rng.uniform(8, 14)), vf=float(rng.uniform(22, 32)),
                                   v_change_steps=[int(rng.integers(2, 5)), int(rng.integers(5, 9))], trajectory_len=T + 10 + 12, ts=1).get_leader_trajectory()
    pol = ConstantSpacingPolicy(50) if rng.random() < 0.5 else ConstantTimePolicy(10, 3)
    scen.append(dict(n=n, N=N, x0=x0, leader_x=lx, masses=None, spacing_policy=pol))
groups = {}
for sc in scen:
    groups.setdefault((sc["N"], spacing_params(sc["spacing_policy"])), []).append(sc)
import os
KEYS = {"a": (10, (10.0, 3.0)), "b": (9, (10.0, 3.0)), "c": (10, (50.0, 0.0))}
for key in [KEYS[k] for k in os.environ.get("HEAVY_GROUPS", "abc")]:
    g = groups[key]
    run_mixed_sweep(g[:16], 1, device=0, ctx=ctx)
    for rep in range(int(os.environ.get("REPS", 5))):
        gc.collect(); torch.cuda.synchronize(); t0 = time.perf_counter(); out = run_mixed_sweep(g, T, device=0, ctx=ctx); torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        nodes = np.concatenate([r["nodes"].ravel() for r in out.values()])
        per_t = np.stack([np.concatenate([r["nodes"][t].ravel() for r in out.values()]) for t in range(T)])
        print(f"N={key[0]} policy={key[1]} rep {rep}: {dt*1e3:7.1f} ms nodes sum {nodes.sum()} max {nodes.max()}  per-step max {per_t.max(1).tolist()}", flush=True)
