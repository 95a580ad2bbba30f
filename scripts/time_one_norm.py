"""Dev helper: throughput / work of the 1-norm (MILP) variant next to the 2-norm one, C1 and C2 shapes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hybrid_vehicle_platoon_b200 as hvp
from hybrid_vehicle_platoon_b200 import synth_mpc as G
from hybrid_vehicle_platoon_b200.synth_local import platoon_local_problems
ctx = hvp.Context(0)
rng = np.random.default_rng(5)
for n, N, B in ((3, 5, 2048),):
    x0, params = G.cent_cases(rng, B, n, N, False)
    for on in (False, True):
        mpc = hvp.api.CompiledMpc(G.CENT, N, n_local=n, one_norm=on, ctx=ctx)
        mpc.solve(x0[:64], 800.0, params[:64])
        t0 = time.perf_counter(); r = mpc.solve(x0, 800.0, params); dt = time.perf_counter() - t0
        print(f"cent n={n} N={N} one_norm={on}: {B/dt:.0f} solves/s, nodes {r['nodes'].mean():.1f}, iters {r['qp_iters'].mean():.0f}, "
              f"status2 {(r['status']==2).mean():.4f} other {np.unique(r['status'])}")
c = platoon_local_problems(rng, 1024, 10, 6)
for fl in (0, 3, 4):
    sel = np.nonzero(c["flags"] == fl)[0]
    params = np.concatenate([c[k][sel].reshape(len(sel), -1) for k in ("xf", "xb", "xl")], axis=1)
    for on in (False, True):
        mpc = hvp.api.CompiledMpc(G.LOCAL, 6, flags=fl, one_norm=on, ctx=ctx)
        mpc.solve(c["x0"][sel][:32].reshape(-1, 1, 2), c["mass"][sel][:32].reshape(-1, 1), params[:32])
        t0 = time.perf_counter(); r = mpc.solve(c["x0"][sel].reshape(-1, 1, 2), c["mass"][sel].reshape(-1, 1), params); dt = time.perf_counter() - t0
        print(f"local N=6 flags={fl} one_norm={on}: {len(sel)/dt:.0f} solves/s, nodes {r['nodes'].mean():.1f}, iters {r['qp_iters'].mean():.0f}, "
              f"status2 {(r['status']==2).mean():.4f} other {np.unique(r['status'])}")
