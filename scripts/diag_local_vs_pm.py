"""Dev helper: per-vehicle local MIQPs through the specialised kernel vs the compiled-MPC kernel (LOCAL kind: hull
tightening + tree splitting of heavy problems), on hard instances (long horizon, time-headway spacing)."""
import sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch
import hybrid_vehicle_platoon_b200 as hvp
from gen_cases import platoon_local_problems
G_LOCAL = 2
for (d0, t0) in ((50.0, 0.0), (10.0, 3.0)):
    for N in (6, 8, 10):
        for S in (53, 2048):
            rng = np.random.default_rng(7)
            cs = platoon_local_problems(rng, S, 10, N)
            t = time.perf_counter()
            r = hvp.local_miqp(N, cs["flags"], cs["mass"], cs["x0"], cs["xf"], cs["xb"], cs["xl"], d0=d0, t0=t0)
            t_loc = time.perf_counter() - t
            t_pm, ok, nodes_pm = 0.0, True, []
            for fl in np.unique(cs["flags"]):
                sel = cs["flags"] == fl
                mpc = hvp.api.CompiledMpc(G_LOCAL, N, flags=int(fl), d0=d0, t0=t0)
                params = np.concatenate([cs[k][sel].reshape(sel.sum(), -1) for k in ("xf", "xb", "xl")], axis=1)
                mpc.solve(cs["x0"][sel][:2, None, :], cs["mass"][sel][:2, None], params[:2])
                t = time.perf_counter()
                g = mpc.solve(cs["x0"][sel][:, None, :], cs["mass"][sel][:, None], params)
                t_pm += time.perf_counter() - t
                good = (g["status"] == 2) & (r["status"][sel] == 2)
                ok &= bool((g["status"] == r["status"][sel]).all()) and bool(np.allclose(g["obj"][good], r["obj"][sel][good], rtol=1e-8))
                nodes_pm.append(g["nodes"])
            nodes_pm = np.concatenate(nodes_pm)
            print(f"d0={d0} t0={t0} N={N} S={S}: local {t_loc*1e3:8.1f} ms nodes {r['nodes'].mean():7.1f}/{r['nodes'].max():6d} | "
                  f"compiled {t_pm*1e3:8.1f} ms nodes {nodes_pm.mean():7.1f}/{nodes_pm.max():6d} agree={ok}", flush=True)
