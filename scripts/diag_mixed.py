import sys, time, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import hybrid_vehicle_platoon_b200 as hvp
from hybrid_vehicle_platoon_b200.sweep import BatchedDecentSweep
from hybrid_vehicle_platoon_b200.misc import ConstantSpacingPolicy, ConstantTimePolicy, StopAndGoLeaderTrajectory
rng = np.random.default_rng(3)
T = 5
for pol in (ConstantSpacingPolicy(50), ConstantTimePolicy(10, 3)):
    for N in (4, 6, 8, 9, 10):
        for S in (53, 4096):
            n = 10
            v = np.floor(rng.uniform(8, 30, (S, n))); gaps = rng.uniform(60, 160, (S, n))
            p = np.floor(3000.0 - np.cumsum(gaps, 1) + gaps[:, :1])
            x0 = np.empty((S, 2 * n)); x0[:, 0::2] = p; x0[:, 1::2] = v
            lx = StopAndGoLeaderTrajectory(p=3000, vh=20, vl=10, vf=30, v_change_steps=[2, 4], trajectory_len=T + 22, ts=1).get_leader_trajectory()
            sw = BatchedDecentSweep(n, N, spacing_policy=pol)
            sw.run(x0[:8], lx, 1)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            out = sw.run(x0, lx, T)
            dt = time.perf_counter() - t0
            print(type(pol).__name__, "N", N, "S", S, f"{dt*1e3:8.1f} ms  nodes mean {out['nodes'].mean():8.1f} max {out['nodes'].max()}  opt {(out['status']==2).mean():.3f}", flush=True)
