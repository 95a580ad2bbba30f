"""Run-to-run variation of the mixed-size sweep (bench.py mixed_sweep_leg): time and total node count of repeated runs."""
import os, sys, time, gc
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import hybrid_vehicle_platoon_b200 as hvp
from hybrid_vehicle_platoon_b200.misc import ConstantSpacingPolicy, ConstantTimePolicy, StopAndGoLeaderTrajectory
from hybrid_vehicle_platoon_b200.sweep import run_mixed_sweep
S, T = 4096, 10
ctx = hvp.Context(0)
rng = np.random.default_rng(1234 + 3)
scen = []
for _ in range(S):
    n, N = int(rng.integers(5, 16)), int(rng.integers(4, 11))
    v = np.floor(rng.uniform(8, 30, n)); gaps = rng.uniform(60, 160, n)
    p = np.floor(3000.0 - np.cumsum(gaps) + gaps[0])
    x0 = np.empty(2 * n); x0[0::2] = p; x0[1::2] = v
    lx = StopAndGoLeaderTrajectory(p=3000, vh=20, vl=float(rng.uniform(8, 14)), vf=float(rng.uniform(22, 32)),
                                   v_change_steps=[int(rng.integers(2, 5)), int(rng.integers(5, 9))],
                                   trajectory_len=T + 10 + 12, ts=1).get_leader_trajectory()
    pol = ConstantSpacingPolicy(50) if rng.random() < 0.5 else ConstantTimePolicy(10, 3)
    scen.append(dict(n=n, N=N, x0=x0, leader_x=lx, masses=None, spacing_policy=pol))
run_mixed_sweep(scen[:64], 2, rank=0, world=1, device=0, ctx=ctx)
for rep in range(int(os.environ.get("REPS", 6))):
    gc.collect(); torch.cuda.synchronize()
    t0 = time.perf_counter(); out = run_mixed_sweep(scen, T, rank=0, world=1, device=0, ctx=ctx); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    nodes = sum(int(r["nodes"].sum()) for r in out.values())
    big = max(int(r["nodes"].max()) for r in out.values())
    print("run %d: %.3f s  nodes %d  largest tree %d" % (rep, dt, nodes, big), flush=True)
