"""Dev diagnostic (GPU box): replay one golden closed loop on libhvp.so and shadow EVERY compiled-MPC solve with the
CPU oracle on the same inputs; report the solves where the two disagree (objective, modes, inputs).
usage: python scripts/diag_golden_case.py admm_default_s3"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import hybrid_vehicle_platoon_b200 as hvp
from hybrid_vehicle_platoon_b200 import api
from oracle import oracle as O
import test_fleet_golden as T

name = sys.argv[1]
calls = [0]
orig = api.CompiledMpc.solve


def shadow(self, x0, mass, params, fixed_modes=None):
    r = orig(self, x0, mass, params, fixed_modes=fixed_modes)
    d = self.desc
    ro = O.mpc_solve(d.kind, d.n_local, d.N, np.asarray(x0, dtype=np.float64).reshape(-1, d.n_local, 2), mass, params,
                     fixed_modes=fixed_modes, method=1, model=d.model, flags=d.flags, leader_index=d.leader_index,
                     n_front=d.n_front, n_behind=d.n_behind, d0=d.d0, t0=d.t0, tight=d.tight, rho=d.rho)
    calls[0] += 1
    for j in range(len(r["obj"])):
        rel = abs(r["obj"][j] - ro["obj"][j]) / max(1.0, abs(ro["obj"][j]))
        du = np.abs(r["u"][j] - ro["u"][j]).max()
        mm = (r["modes"][j] != ro["modes"][j]).any()
        dx = np.abs(r["x"][j] - ro["x"][j]).max()
        de = np.abs(r["extra"][j] - ro["extra"][j]).max() if r.get("extra") is not None and r["extra"].size else 0.0
        if dx > 1e-6 or de > 1e-6:
            print(f"call {calls[0]} problem {j}: dx {dx:.2e} dextra {de:.2e} (du {du:.2e}, rel obj {rel:.2e}) kind {d.kind} flags {d.flags}")
            if de > 1e-6:
                print("   extra gpu   ", np.round(r["extra"][j], 6)); print("   extra oracle", np.round(ro["extra"][j], 6))
        if rel > 1e-6 or du > 1e-6 or mm or r["status"][j] != ro["status"][j]:
            print(f"call {calls[0]} problem {j}: status {r['status'][j]}/{ro['status'][j]} obj {r['obj'][j]:.10f} / {ro['obj'][j]:.10f} "
                  f"(rel {rel:.2e}) du {du:.2e} modes gpu {r['modes'][j].ravel()} oracle {ro['modes'][j].ravel()} nodes {r['nodes'][j]}")
    return r


api.CompiledMpc.solve = shadow
out = T.replay(name)
g = T.G
dU = np.abs(out["U"] - g[f"{name}/U"]).max(axis=tuple(range(1, out["U"].ndim)))
first = np.argmax(dU > 1e-7) if (dU > 1e-7).any() else -1
print(f"{name}: {calls[0]} solves shadowed; max |dU| {dU.max():.3e}, first timestep with |dU| > 1e-7: {first}; max |dX| {np.abs(out['X'] - g[f'{name}/X']).max():.3e}")
