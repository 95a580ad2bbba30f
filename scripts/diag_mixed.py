"""Dev helper: closed-loop decentralized sweep, specialised local kernel vs compiled-MPC LOCAL formulation, by horizon and
spacing policy (n = 10, 5 timesteps, 300 and 4096 scenarios) -- the measurements behind BatchedDecentSweep's "auto" rule."""
import sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch
import hybrid_vehicle_platoon_b200 as hvp
from hybrid_vehicle_platoon_b200.sweep import BatchedDecentSweep
from hybrid_vehicle_platoon_b200.misc import ConstantSpacingPolicy, ConstantTimePolicy, StopAndGoLeaderTrajectory
T = 5
for pol in (ConstantSpacingPolicy(50), ConstantTimePolicy(10, 3)):
    for N in (5, 6, 7, 8, 9, 10):
        for S in (300, 4096):
            rng = np.random.default_rng(3)
            n = 10
            v = np.floor(rng.uniform(8, 30, (S, n))); gaps = rng.uniform(60, 160, (S, n))
            p = np.floor(3000.0 - np.cumsum(gaps, 1) + gaps[:, :1])
            x0 = np.empty((S, 2 * n)); x0[:, 0::2] = p; x0[:, 1::2] = v
            lx = StopAndGoLeaderTrajectory(p=3000, vh=20, vl=10, vf=30, v_change_steps=[2, 4], trajectory_len=T + 22, ts=1).get_leader_trajectory()
            res = []
            for solver in ("local", "compiled"):
                sw = BatchedDecentSweep(n, N, spacing_policy=pol, solver=solver)
                sw.run(x0[:8], lx, 1)
                sw.run(x0, lx, T)                       # full-size warm-up (allocator, clocks)
                torch.cuda.synchronize(); t0 = time.perf_counter()
                out = sw.run(x0, lx, T)
                res.append(((time.perf_counter() - t0) * 1e3, out["nodes"].mean(), out["nodes"].max(), (out["status"] == 2).mean()))
            auto = BatchedDecentSweep(n, N, spacing_policy=pol).use_compiled
            print(f"{type(pol).__name__:22s} N={N:2d} S={S:4d}: local {res[0][0]:8.1f} ms (nodes {res[0][1]:6.1f}/{res[0][2]:6d}) | "
                  f"compiled {res[1][0]:8.1f} ms (nodes {res[1][1]:6.1f}/{res[1][2]:6d}) | auto -> {'compiled' if auto else 'local'}", flush=True)
