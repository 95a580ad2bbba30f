"""Run-to-run variation of the g-ADMM closed-loop leg (bench.py gadmm_loop_leg) over repeated runs of one sweep object."""
import os, sys, time, gc
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import hybrid_vehicle_platoon_b200 as hvp
from hybrid_vehicle_platoon_b200.sweep import BatchedGAdmmSweep
from hybrid_vehicle_platoon_b200.misc import ConstantVelocityLeaderTrajectory
S, T, n, N, iters = 1024, 2, 15, 8, 100
ctx = hvp.Context(0)
rng = np.random.default_rng(1234 + 2)
v = np.floor(rng.uniform(12, 28, (S, n))); gaps = rng.uniform(60, 120, (S, n))
p = np.floor(3000.0 - np.cumsum(gaps, 1) + gaps[:, :1])
x0 = np.empty((S, 2 * n)); x0[:, 0::2] = p; x0[:, 1::2] = v
lx = ConstantVelocityLeaderTrajectory(p=3000, v=20, trajectory_len=T + N + 10, ts=1).get_leader_trajectory()
for kw in (dict(), dict(graph=False)):
    sw = BatchedGAdmmSweep(n, N, admm_iters=iters, rho=0.5, ctx=ctx, **kw)
    sw.run(x0[:16], lx, 1)
    for rep in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter(); out = sw.run(x0, lx, T); dt = time.perf_counter() - t0
        t1 = time.perf_counter(); gc.collect(); torch.cuda.synchronize(); dg = time.perf_counter() - t1
        print(kw, "run %d: %.3f s  (gc+sync after: %.3f s)  solved %.4f" % (rep, dt, dg, out["solved"].mean()), flush=True)
