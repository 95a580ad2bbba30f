"""Randomised parity campaign of the 1-norm (MILP) variant: CUDA kernels vs scipy.optimize.milp (HiGHS) on the explicit
big-M MLD model (tests/mld_bigm.py).  usage: python scripts/stress_one_norm.py [seed] [cases per shape]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import hybrid_vehicle_platoon_b200 as hvp
import mld_bigm as MB
import gen_mpc_cases as G
from gen_cases import platoon_local_problems
from hybrid_vehicle_platoon_b200.models import Platoon

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
per = int(sys.argv[2]) if len(sys.argv) > 2 else 48
rng = np.random.default_rng(seed)
ctx = hvp.Context(0)
bad = tot = polished = 0
t00 = time.time()


def polish(M, xs, obj):
    """HiGHS accepts binaries within 1e-6 of an integer, which through a big-M row buys up to M*1e-6 of slack: re-solve the
    LP with the binaries fixed at the rounded values (tolerance 1e-9) so that the comparison is with a true mixed-integer point."""
    fixed = {k: round(float(xs[k])) for k in range(M.n) if M.integ[k]}
    ok, _, o = MB.solve_qp_fixed(M, fixed)
    return o if ok else obj

for N, stress, t0, d0 in ((3, True, 0.0, 50.0), (4, True, 3.0, 10.0), (5, False, 0.0, 50.0), (5, True, 3.0, 10.0)):
    c = platoon_local_problems(rng, per // 4, 4, N, 0, stress, True)
    worst = 0.0
    for fl in sorted(set(int(f) for f in c["flags"])):
        sel = np.nonzero(c["flags"] == fl)[0]
        mpc = hvp.api.CompiledMpc(G.LOCAL, N, flags=fl, d0=d0, t0=t0, one_norm=True, ctx=ctx)
        params = np.concatenate([c[k][sel].reshape(len(sel), -1) for k in ("xf", "xb", "xl")], axis=1)
        r = mpc.solve(c["x0"][sel].reshape(-1, 1, 2), c["mass"][sel].reshape(-1, 1), params)
        for j, b in enumerate(sel):
            sysd = Platoon(1, "pwa_gear", masses=[float(c["mass"][b])]).get_vehicle_system_dicts(1.0)[0]
            M, x, u, dl = MB.build_local(sysd, N, c["x0"][b], c["xf"][b], c["xb"][b], c["xl"][b], is_front=bool(fl & 1),
                                         is_leader=bool(fl & 2), is_trailer=bool(fl & 4), d0=d0, t0=t0, quadratic=False)
            ok, xs, obj = MB.solve_milp(M)
            tot += 1
            if not ok:
                if r["status"][j] != 3:
                    bad += 1; print("  status mismatch (milp infeasible)", N, fl, b, r["status"][j])
                continue
            rel = abs(r["obj"][j] - obj) / max(1.0, abs(obj))
            if rel > 1e-6 and r["obj"][j] > obj:
                obj = polish(M, xs, obj); polished += 1
                rel = abs(r["obj"][j] - obj) / max(1.0, abs(obj))
            worst = max(worst, rel)
            if r["status"][j] != 2 or rel > 1e-6:
                bad += 1; print(f"  MISMATCH local N={N} flags={fl} case {b}: status {r['status'][j]} gpu {r['obj'][j]:.9f} milp {obj:.9f} rel {rel:.2e}")
    print(f"local N={N} stress={stress} t0={t0}: {len(c['flags'])} cases, worst rel obj {worst:.2e}", flush=True)
for n, N in ((2, 3), (2, 4), (3, 3)):
    Bn = max(4, per // 6)
    x0, params = G.cent_cases(rng, Bn, n, N, True)
    r = hvp.api.CompiledMpc(G.CENT, N, n_local=n, one_norm=True, ctx=ctx).solve(x0, 800.0, params)
    systems = Platoon(n, "pwa_gear", masses=[800.0] * n).get_vehicle_system_dicts(1.0)
    worst = 0.0
    for b in range(Bn):
        M, xs, us, ds = MB.build_cent(systems, N, x0[b], params[b].reshape(2, N + 1), quadratic=False)
        ok, sol, obj = MB.solve_milp(M, time_limit=120.0)
        tot += 1
        if not ok:
            if r["status"][b] != 3:
                bad += 1; print("  status mismatch (milp infeasible)", n, N, b, r["status"][b])
            continue
        rel = abs(r["obj"][b] - obj) / max(1.0, abs(obj))
        if rel > 1e-6 and r["obj"][b] > obj:
            obj = polish(M, sol, obj); polished += 1
            rel = abs(r["obj"][b] - obj) / max(1.0, abs(obj))
        worst = max(worst, rel)
        if r["status"][b] != 2 or rel > 1e-6:
            bad += 1; print(f"  MISMATCH cent n={n} N={N} case {b}: status {r['status'][b]} gpu {r['obj'][b]:.9f} milp {obj:.9f} rel {rel:.2e}")
    print(f"cent n={n} N={N}: {Bn} cases, worst rel obj {worst:.2e}", flush=True)
print(f"seed {seed}: {tot} cases, {bad} mismatches, {polished} HiGHS optima re-solved with fixed binaries, {time.time() - t00:.0f} s")
sys.exit(1 if bad else 0)
