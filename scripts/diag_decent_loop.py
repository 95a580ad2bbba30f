"""Where the time of the decentralized closed-loop leg (bench.py closed_loop_leg: 4096 platoons x 20 steps) goes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import hybrid_vehicle_platoon_b200 as hvp
from hybrid_vehicle_platoon_b200.sweep import BatchedDecentSweep
from hybrid_vehicle_platoon_b200.misc import StopAndGoLeaderTrajectory

S, n, N = int(os.environ.get("S", 4096)), 10, 6
ctx = hvp.Context(0)
rng = np.random.default_rng(1234 + 3)
v = np.floor(rng.uniform(5, 35, (S, n))); gaps = rng.uniform(60, 160, (S, n))
p = np.floor(3000.0 - np.cumsum(gaps, 1) + gaps[:, :1])
x0 = np.empty((S, 2 * n)); x0[:, 0::2] = p; x0[:, 1::2] = v
lx = StopAndGoLeaderTrajectory(p=3000, vh=20, vl=10, vf=30, v_change_steps=[5, 12], trajectory_len=100, ts=1).get_leader_trajectory()
for kw in (dict(fused=False), dict(), dict(graph=True), dict(fused=False), dict()):
    sw = BatchedDecentSweep(n, N, ctx=ctx, **kw)
    sw.run(x0[:256], lx, 3)
    for T in (20, 20, 40, 40):
        torch.cuda.synchronize()
        t0 = time.perf_counter(); out = sw.run(x0, lx, T); dt = time.perf_counter() - t0
        print(kw, "T", T, "%.1f ms" % (dt * 1e3), "nodes %.2f" % out["nodes"].mean(), flush=True)
