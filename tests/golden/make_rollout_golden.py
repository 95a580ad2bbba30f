"""Generate golden vectors for the rollout/cost half of the hot path by running the
UNMODIFIED reference (`/root/reference/env.py`, `models.py`, `misc/*`) under the import
stubs in `oracle/refshim`.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_rollout_golden.py

Writes `tests/golden/rollout_golden.npz` (committed).  Test infrastructure, not product.
"""
import contextlib
import io
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))

from env import PlatoonEnv  # noqa: E402  (reference)
from misc.common_controller_params import Sim, Sim_n_task_2  # noqa: E402
from misc.leader_trajectory import (  # noqa: E402
    ConstantVelocityLeaderTrajectory,
    StopAndGoLeaderTrajectory,
)
from misc.spacing_policy import ConstantSpacingPolicy, ConstantTimePolicy  # noqa: E402
from models import GearTransimission, Platoon, PwaGearVehicle, Vehicle  # noqa: E402

quiet = contextlib.redirect_stdout(io.StringIO())


def ref_step(n, masses, spacing, leader_index, x, u, gears, leader_col, quadratic=True,
             real_ref=False):
    """One reference env.step from an arbitrary state. Returns (x_new, r, viol, gears_used)
    or the raised exception class name."""
    platoon = Platoon(n, "pwa_gear", masses=None if masses is None else list(masses))
    if spacing[1] == 0.0:
        pol = ConstantSpacingPolicy(spacing[0])
    else:
        pol = ConstantTimePolicy(spacing[0], spacing[1])
    env = PlatoonEnv(n=n, platoon=platoon, ep_len=4, leader_index=leader_index,
                     spacing_policy=pol, quadratic_cost=quadratic,
                     real_vehicle_as_reference=real_ref)
    env.reset(seed=0)
    env.x = x.reshape(2 * n, 1).astype(np.float64).copy()
    env.leader_x = np.tile(leader_col.reshape(2, 1), (1, 8))
    env.step_counter = 1
    if gears is None:
        action = u.reshape(n, 1).copy()
        g_used = np.array([platoon.get_gear_from_vehicle_velocity(i, x[2 * i + 1])
                           for i in range(n)], dtype=np.int32)
    else:
        action = np.vstack([u.reshape(n, 1), gears.reshape(n, 1).astype(np.float64)])
        g_used = gears.astype(np.int32)
    with quiet, warnings.catch_warnings():
        warnings.simplefilter("ignore")
        x_new, r, _, _, _ = env.step(action)
    viol = int(env.viol_counter[-1][1] == 100)
    return x_new.ravel(), float(np.asarray(r).ravel()[0]), viol, g_used


def valid_gears_for(v):
    gm = GearTransimission()
    ok = [j + 1 for j in range(6) if gm.v[j][0] + 2.5 < v < gm.v[j][3] - 5.5]
    return ok if ok else [PwaGearVehicle(800).get_gear_from_velocity(v)]


def make_cases(rng, n, count, hetero, spacing, given_gears, leader_index=0, quadratic=True,
               real_ref=False):
    X, U, G, M, L, XN, R, V = [], [], [], [], [], [], [], []
    while len(X) < count:
        v = rng.uniform(5.0, 34.0, n)
        gaps = rng.uniform(15.0, 160.0, n)  # some below d_safe=25 -> violations
        p = 3000.0 - np.cumsum(gaps) + gaps[0]
        x = np.empty(2 * n)
        x[0::2] = p
        x[1::2] = v
        u = rng.uniform(-1.0, 1.0, n)
        masses = rng.uniform(700.0, 1000.0, n) if hetero else None
        if given_gears:
            g = np.array([rng.choice(valid_gears_for(vi)) for vi in v], dtype=np.int32)
        else:
            g = None
        leader = np.array([p[leader_index] + rng.uniform(-30, 30), rng.uniform(10, 30)])
        try:
            x_new, r, viol, g_used = ref_step(n, masses, spacing, leader_index, x, u, g, leader,
                                              quadratic, real_ref)
        except RuntimeError:
            continue  # reference raised (Q6); error cases are covered separately below
        X.append(x); U.append(u); G.append(g_used)
        M.append(masses if hetero else np.full(n, 800.0))
        L.append(leader); XN.append(x_new); R.append(r); V.append(viol)
    return dict(x=np.array(X), u=np.array(U), gear=np.array(G), mass=np.array(M),
                leader=np.array(L), x_new=np.array(XN), r=np.array(R),
                viol=np.array(V, dtype=np.uint8))


def main():
    out = {}
    rng = np.random.default_rng(20261018)

    # ---- (1) random single steps over the configurations named in SURVEY.md 7.2 -------------
    cfgs = []
    for n in (3, 10, 15):
        for hetero in (False, True):
            for spacing in ((50.0, 0.0), (10.0, 3.0)):
                for given in (False, True):
                    cfgs.append((n, hetero, spacing, given, 0, True, False))
    cfgs.append((4, True, (10.0, 3.0), True, 3, True, False))   # leader at the back
    cfgs.append((5, False, (50.0, 0.0), False, 2, True, False))  # leader in the middle
    cfgs.append((3, False, (50.0, 0.0), False, 0, False, False))  # 1-norm stage cost
    cfgs.append((4, True, (10.0, 3.0), True, 0, False, False))   # 1-norm + headway
    cfgs.append((3, False, (50.0, 0.0), False, 0, True, True))   # real vehicle as reference
    meta = []
    for ci, (n, hetero, spacing, given, li, quad, rr) in enumerate(cfgs):
        c = make_cases(rng, n, 24, hetero, spacing, given, li, quad, rr)
        for k, v in c.items():
            out[f"c{ci}_{k}"] = v
        meta.append([n, int(hetero), spacing[0], spacing[1], int(given), li, int(quad), int(rr)])
    out["cfg_meta"] = np.array(meta, dtype=np.float64)

    # ---- (2) KATs of SURVEY.md 4.3: seeds, resets, two steps --------------------------------
    out["seedseq"] = np.array([np.random.SeedSequence(s).generate_state(1)[0] for s in range(4)],
                              dtype=np.uint64)
    for tag, seed in (("kat0", 2968811710), ("kat1", 1835504127)):
        platoon = Platoon(3, "pwa_gear")
        env = PlatoonEnv(3, platoon, 150)
        x0, _ = env.reset(seed=seed)
        xs, rs = [np.asarray(x0, dtype=np.float64).ravel()], []
        with quiet:
            for _ in range(2):
                xn, r, *_ = env.step(np.array([[0.3], [-0.2], [0.1]]))
                xs.append(xn.ravel()); rs.append(float(np.asarray(r).ravel()[0]))
        out[f"{tag}_x"] = np.array(xs)
        out[f"{tag}_r"] = np.array(rs)
    # resets for several n / seeds (Q1 int truncation, Q2 legacy RNG)
    for n in (3, 5, 10, 15):
        xs = []
        for s in range(4):
            seed = int(np.random.SeedSequence(s).generate_state(1)[0])
            env = PlatoonEnv(n, Platoon(n, "pwa_gear"), 150)
            x0, _ = env.reset(seed=seed)
            xs.append(np.asarray(x0).ravel().astype(np.float64))
        out[f"reset_n{n}"] = np.array(xs)
    # start_from_platoon resets (both policies)
    for tag, pol in (("const", ConstantSpacingPolicy(50)), ("headway", ConstantTimePolicy(10, 3))):
        env = PlatoonEnv(4, Platoon(4, "pwa_gear"), 150, spacing_policy=pol,
                         start_from_platoon=True)
        x0, _ = env.reset(seed=1)
        out[f"reset_platoon_{tag}"] = np.asarray(x0).ravel().astype(np.float64)

    # task-2 KAT (heterogeneous masses, explicit gears)
    sim = Sim_n_task_2(4, seed=0)
    out["task2_masses"] = np.array(sim.masses)
    platoon = Platoon(4, "pwa_gear", masses=sim.masses)
    env = PlatoonEnv(4, platoon, 150, spacing_policy=sim.spacing_policy,
                     leader_trajectory=sim.leader_trajectory)
    x0, _ = env.reset(seed=2968811710)
    with quiet:
        x1, r, *_ = env.step(np.array([[0.5], [-0.5], [0.25], [1.0], [4], [4], [2], [4]]))
    out["task2_x0"] = np.asarray(x0).ravel().astype(np.float64)
    out["task2_x1"] = x1.ravel()
    out["task2_r"] = np.array([float(np.asarray(r).ravel()[0])])
    out["task2_leader"] = sim.leader_trajectory.get_leader_trajectory()
    out["const_leader"] = Sim().leader_trajectory.get_leader_trajectory()

    # ---- (3) scalar tables -------------------------------------------------------------------
    gm = GearTransimission()
    tv = np.array([(3.0, 1), (6.0, 1), (12.82, 2), (12.9, 3), (20, 4), (33, 5), (33, 6),
                   (7.0, 4), (9.9, 4), (4.5, 2), (10.5, 1), (30.0, 4), (45.0, 6), (8.0, 3)])
    out["traction_in"] = tv
    out["traction_out"] = np.array([gm.get_traction(float(v), int(j)) for v, j in tv])
    veh = PwaGearVehicle(800)
    vs = np.array([5, 9.235, 12.855, 16.93, 22.92, 23.315, 32.47, 40, 9.2349999, 3.0, 50.0])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out["gearmap_in"] = vs
        out["gearmap_out"] = np.array([veh.get_gear_from_velocity(float(v)) for v in vs])
    out["v_gear_lim"] = np.array(veh.v_gear_lim)
    out["constvel_in"] = np.array([10.0, 20.0, 30.0])
    out["constvel_out"] = np.array([veh.get_u_for_constant_vel(v) for v in (10.0, 20.0, 30.0)])
    # PWA discrete systems for two masses: a_r = A[r][1,1], b_r = B[r][1,0], c_r = c[r][1,0],
    # region bounds from S,T
    for m in (800.0, 914.5568099117258):
        sysd = PwaGearVehicle(m).get_discrete_system(1.0)
        a = np.array([A[1, 1] for A in sysd["A"]])
        b = np.array([B[1, 0] for B in sysd["B"]])
        c = np.array([cc[1, 0] for cc in sysd["c"]])
        hi = np.array([T[0, 0] if S[0, 1] == 1 else np.inf for S, T in zip(sysd["S"], sysd["T"])])
        lo = np.array([-T[1, 0] if S[1, 1] == -1 else -np.inf for S, T in zip(sysd["S"], sysd["T"])])
        assert all((A[0] == [1, 1]).all() for A in sysd["A"])
        key = "m800" if m == 800.0 else "m914"
        out[f"pwa_{key}"] = np.vstack([a, b, c, lo, hi])
        out[f"pwa_{key}_DEFG"] = np.concatenate([sysd["E"].ravel(), sysd["G"].ravel()])
    out["vehicle_consts"] = np.array([Vehicle.c_fric, Vehicle.mu, Vehicle.grav, Vehicle.v_min,
                                      Vehicle.v_max, PwaGearVehicle.alpha, PwaGearVehicle.c1,
                                      PwaGearVehicle.c2, PwaGearVehicle.d, PwaGearVehicle.beta])

    # ---- (4) error cases (Q6): (v, gear, u) for a single vehicle that must raise -------------
    errs_in = np.array([
        [2.0, 1, 0.0],      # velocity below true-model bound at entry
        [60.5, 6, 0.0],     # above
        [20.0, 1, 0.0],     # velocity outside gear-1 curve (>= 12.38)
        [5.0, 6, 0.0],      # below gear-6 curve (<= 10.027)
        [20.0, 0, 0.0],     # gear index out of range
        [20.0, 7, 0.0],
        [12.2, 1, 1.0],     # leaves gear-1 range during the 10 sub-steps
        [2.2, 1, -1.0],     # decelerates below 2.0706 during sub-steps
    ])
    kinds = []
    for v, j, u in errs_in:
        platoon = Platoon(1, "pwa_gear")
        try:
            platoon.step_platoon(np.array([[100.0], [v]]), np.array([[u]]), np.array([[j]]), 1.0)
            kinds.append(0)
        except RuntimeError as e:
            msg = str(e)
            if "Gear value" in msg:
                kinds.append(2)
            elif "exeeds true model bounds" in msg or "Velocity value out of range" in msg:
                kinds.append(1)
            else:
                kinds.append(3)
    out["err_in"] = errs_in
    out["err_kind"] = np.array(kinds, dtype=np.int32)

    # ---- (5) a 40-step closed sequence (n=10, headway, hetero masses, derived gears) ---------
    n = 10
    rng2 = np.random.default_rng(7)
    masses = rng2.uniform(700, 1000, n)
    platoon = Platoon(n, "pwa_gear", masses=list(masses))
    traj = StopAndGoLeaderTrajectory(p=3000, vh=20, vl=10, vf=30, v_change_steps=[10, 25],
                                     trajectory_len=100, ts=1)
    env = PlatoonEnv(n, platoon, 60, spacing_policy=ConstantTimePolicy(10, 3),
                     leader_trajectory=traj)
    x0, _ = env.reset(seed=5)
    xs, us, rs = [np.asarray(x0, dtype=np.float64).ravel()], [], []
    with quiet, warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for t in range(40):
            vel = np.asarray(env.x, dtype=np.float64)[1::2]
            u = np.clip(0.1 * (20.0 - vel) + rng2.uniform(-0.4, 0.4, (n, 1)), -1.0, 1.0)
            xn, r, *_ = env.step(u)
            xs.append(xn.ravel()); us.append(u.ravel()); rs.append(float(np.asarray(r).ravel()[0]))
    out["seq_masses"] = masses
    out["seq_leader"] = traj.get_leader_trajectory()
    out["seq_x"] = np.array(xs); out["seq_u"] = np.array(us); out["seq_r"] = np.array(rs)
    out["seq_viol"] = np.array(env.viol_counter[-1][:40])

    path = os.path.join(HERE, "rollout_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
