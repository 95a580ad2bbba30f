"""One 1-norm case of scripts/stress_one_norm.py in detail: usage diag_one_norm_case.py seed per N flags case."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import hybrid_vehicle_platoon_b200 as hvp
import mld_bigm as MB
import gen_mpc_cases as G
from gen_cases import platoon_local_problems
from hybrid_vehicle_platoon_b200.models import Platoon

seed, per, wantN, wantfl, case = (int(a) for a in sys.argv[1:6])
want_stress = len(sys.argv) < 7 or bool(int(sys.argv[6]))
rng = np.random.default_rng(seed)
ctx = hvp.Context(0)
for N, stress, t0, d0 in ((3, True, 0.0, 50.0), (4, True, 3.0, 10.0), (5, False, 0.0, 50.0), (5, True, 3.0, 10.0)):
    c = platoon_local_problems(rng, per // 4, 4, N, 0, stress, True)
    if N != wantN or stress != want_stress or int(c["flags"][case]) != wantfl:
        continue
    b, fl = case, wantfl
    mpc = hvp.api.CompiledMpc(G.LOCAL, N, flags=fl, d0=d0, t0=t0, one_norm=True, ctx=ctx)
    params = np.concatenate([c[k][[b]].reshape(1, -1) for k in ("xf", "xb", "xl")], axis=1)
    r = mpc.solve(c["x0"][[b]].reshape(-1, 1, 2), c["mass"][[b]].reshape(-1, 1), params)
    sysd = Platoon(1, "pwa_gear", masses=[float(c["mass"][b])]).get_vehicle_system_dicts(1.0)[0]
    M, x, u, dl = MB.build_local(sysd, N, c["x0"][b], c["xf"][b], c["xb"][b], c["xl"][b], is_front=bool(fl & 1),
                                 is_leader=bool(fl & 2), is_trailer=bool(fl & 4), d0=d0, t0=t0, quadratic=False)
    ok, xs, obj = MB.solve_milp(M)
    seq = np.array([int(np.argmax([xs[int(dl[rg, k])] for rg in range(dl.shape[0])])) for k in range(N)], dtype=np.int32)
    print("x0", c["x0"][b], "mass", c["mass"][b], "stress", stress)
    print("gpu  obj %.9f status %d nodes %d modes %s" % (r["obj"][0], r["status"][0], r["nodes"][0], r["modes"][0].reshape(-1)))
    print("milp obj %.9f modes %s" % (obj, seq))
    for name, md in (("gpu", r["modes"][0].reshape(-1)), ("milp", seq)):
        fixed = {int(dl[rg, k]): (1.0 if md[k] == rg else 0.0) for rg in range(dl.shape[0]) for k in range(N)}
        okf, xf_, of = MB.solve_qp_fixed(M, fixed)
        rf = mpc.solve(c["x0"][[b]].reshape(-1, 1, 2), c["mass"][[b]].reshape(-1, 1), params, fixed_modes=np.asarray(md, np.int32).reshape(1, 1, N))
        print(f"  sequence of {name}: HiGHS LP {of:.9f} (ok {okf})   GPU LP {rf['obj'][0]:.9f} status {rf['status'][0]}")
        if name == "milp":
            print("   HiGHS u", xf_[np.asarray(u).reshape(-1)], " x", xf_[np.asarray(x).reshape(-1)].reshape(np.asarray(x).shape))
            print("   GPU   u", rf["u"][0].reshape(-1), " x", rf["x"][0].reshape(-1, 2).T if rf["x"][0].size == 2 * (N + 1) else rf["x"][0])
    for env in ("HVP_MPC_SIBLING",):
        pass
