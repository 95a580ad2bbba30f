"""Dev helper for launch lists: a short g-ADMM closed loop (1024 scenarios, n = 15, N = 8, 6 rounds, eager)."""
import os, sys
os.environ.setdefault("HVP_SWEEP_GRAPH", "0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hybrid_vehicle_platoon_b200 as hvp
from hybrid_vehicle_platoon_b200 import sweep as SW
from hybrid_vehicle_platoon_b200.misc import ConstantVelocityLeaderTrajectory
kind = sys.argv[1] if len(sys.argv) > 1 else "gadmm"
ctx = hvp.Context(0)
rng = np.random.default_rng(1237)
S, T, n, N = 1024, 1, 15, 8
v = np.floor(rng.uniform(12, 28, (S, n))); gaps = rng.uniform(60, 120, (S, n))
p = np.floor(3000.0 - np.cumsum(gaps, 1) + gaps[:, :1])
x0 = np.empty((S, 2 * n)); x0[:, 0::2] = p; x0[:, 1::2] = v
lx = ConstantVelocityLeaderTrajectory(p=3000, v=20, trajectory_len=T + N + 10, ts=1).get_leader_trajectory()
sw = SW.BatchedGAdmmSweep(n, N, admm_iters=6, rho=0.5, ctx=ctx) if kind == "gadmm" else SW.BatchedAdmmSweep(n, N, admm_iters=4, rho=0.5, ctx=ctx)
sw.run(x0, lx, T)
