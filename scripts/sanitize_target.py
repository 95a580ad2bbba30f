"""Small workloads for compute-sanitizer (racecheck / memcheck / synccheck), SURVEY.md section 5.

    compute-sanitizer --tool racecheck python scripts/sanitize_target.py flat
    compute-sanitizer --tool memcheck  python scripts/sanitize_target.py pm

`flat`: the per-lane flat local-MIQP kernel forced onto a small batch, where nearly every tree is
"tail" and the intra-warp sub-tree adoption (shuffles + per-lane scratch rows) runs constantly.
`coop`: the 8-lane cooperative kernel.  `pm`: the compiled-MPC kernel with a tiny first-pass node
budget so that most trees go through the multi-warp split.  `rollout`: the rollout kernel.
Each mode checks its results against the CPU oracle, so a sanitizer run is also a parity run."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
mode = sys.argv[1] if len(sys.argv) > 1 else "flat"
if mode in ("flat", "coop"):
    os.environ["HVP_LOCAL_KERNEL"] = mode
if mode == "pm":
    os.environ["HVP_MPC_BUDGET"] = "4"
    os.environ["HVP_MPC_SPLIT_M"] = "8"

import numpy as np  # noqa: E402
import hybrid_vehicle_platoon_b200 as hvp  # noqa: E402
from oracle import oracle as O  # noqa: E402
from gen_cases import platoon_local_problems  # noqa: E402
import gen_mpc_cases as G  # noqa: E402

rng = np.random.default_rng(7)
if mode in ("flat", "coop"):
    N, n, S = 6, 10, 48
    c = platoon_local_problems(rng, S, n, N, 3, True, True)
    args = (N, c["flags"], c["mass"], c["x0"], c["xf"], c["xb"], c["xl"])
    r = hvp.local_miqp(*args, d0=10.0, t0=3.0)
    ro = O.local_miqp(*args, d0=10.0, t0=3.0)
    ok = ro["status"] == 2
    assert (r["status"] == ro["status"]).all()
    rel = np.abs(r["obj"][ok] - ro["obj"][ok]) / np.maximum(1.0, np.abs(ro["obj"][ok]))
    print(f"{mode}: {len(ok)} problems, {ok.sum()} optimal, max rel obj err {rel.max():.2e}, mean nodes {r['nodes'].mean():.1f}")
    assert rel.max() < 1e-6
elif mode == "pm":
    x0, params = G.cent_cases(rng, 12, 3, 4, stress=True)
    mpc = hvp.CompiledMpc(G.CENT, 4, n_local=3)
    r = mpc.solve(x0, 800.0, params)
    ro = O.mpc_solve(O.CENT, 3, 4, x0, 800.0, params, method=1)
    ok = ro["status"] == 2
    assert (r["status"] == ro["status"]).all()
    rel = np.abs(r["obj"][ok] - ro["obj"][ok]) / np.maximum(1.0, np.abs(ro["obj"][ok]))
    print(f"pm: {len(ok)} problems, {ok.sum()} optimal, max rel obj err {rel.max():.2e}, mean nodes {r['nodes'].mean():.1f}")
    assert rel.max() < 1e-6
else:
    n, B = 10, 512
    v = rng.uniform(6, 33, (B, n)); gaps = rng.uniform(20, 150, (B, n))
    p = 3000.0 - np.cumsum(gaps, 1)
    x = np.empty((B, 2 * n)); x[:, 0::2] = p; x[:, 1::2] = v
    u = rng.uniform(-1, 1, (B, n))
    leader = np.stack([p[:, 0] + 5.0, np.full(B, 20.0)], 1)
    xo, cst, vi, e = hvp.rollout_step(x, u, None, None, leader)
    xr, cr, vr, er = O.env_step(x, u, None, None, leader)
    assert (e == er).all() and np.array_equal(xo[er == 0], xr[er == 0])
    print(f"rollout: {B} scenarios bit-exact")
print("OK")
