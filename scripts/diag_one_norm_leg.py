"""The problems of bench.py's 1-norm per-vehicle leg that do not end with status 2: which, and what HiGHS says."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import hybrid_vehicle_platoon_b200 as hvp
from hybrid_vehicle_platoon_b200 import synth_mpc as G
from hybrid_vehicle_platoon_b200.synth_local import platoon_local_problems
from hybrid_vehicle_platoon_b200.models import Platoon
import mld_bigm as MB
rng = np.random.default_rng(1234 + 2)
# the generator state at this leg depends on the legs before it in bench.py: take a fresh, larger draw instead
N = 6
bad_total = 0
for seed in range(3):
    rng = np.random.default_rng(100 + seed)
    c = platoon_local_problems(rng, 2048, 10, N)
    sel = np.nonzero(c["flags"] == 0)[0]
    pl = np.concatenate([c[k][sel].reshape(len(sel), -1) for k in ("xf", "xb", "xl")], axis=1)
    r = hvp.api.CompiledMpc(G.LOCAL, N, flags=0, one_norm=True).solve(c["x0"][sel].reshape(-1, 1, 2), c["mass"][sel].reshape(-1, 1), pl)
    bad = np.nonzero(r["status"] != 2)[0]
    print(f"seed {seed}: {len(sel)} problems, statuses", dict(zip(*np.unique(r["status"], return_counts=True))), flush=True)
    for j in bad[:6]:
        b = sel[j]
        sysd = Platoon(1, "pwa_gear", masses=[float(c["mass"][b])]).get_vehicle_system_dicts(1.0)[0]
        M, x, u, dl = MB.build_local(sysd, N, c["x0"][b], c["xf"][b], c["xb"][b], c["xl"][b], is_front=False, is_leader=False,
                                     is_trailer=False, d0=50.0, t0=0.0, quadratic=False)
        ok, xs, obj = MB.solve_milp(M)
        print(f"   problem {b}: gpu status {r['status'][j]} obj {r['obj'][j]:.6f} nodes {r['nodes'][j]} | milp ok {ok} obj {obj:.6f}  x0 {c['x0'][b]}")
    bad_total += len(bad)
print("not status 2:", bad_total)
