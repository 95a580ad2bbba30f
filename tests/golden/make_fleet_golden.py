"""Golden closed-loop runs of the REFERENCE's own fleet scripts (coordinators, simulate(), env, misc).

What runs, unmodified, from /root/reference:  fleet_cent_mld.py, fleet_decent_mld.py, fleet_seq_mld.py,
fleet_event_based.py, fleet_naive_admm.py (their coordinator classes and simulate()), env.py,
models.py, misc/*.  What is substituted: the module-level MPC class names the scripts instantiate
(`LocalMpcMld`, `LocalMpc`, `LocalMpcADMM`, `MpcMldCent`, ...) are re-bound to the repo's controller
classes, whose solves go to the CPU oracle here (tests/oracle_backend.py; the build container has no GPU
and no Gurobi), and the un-vendored third-party names (dmpcpwa MldAgent, gymnasium TimeLimit, mpcrl
MonitorEpisodes) resolve to the repo's restatements through oracle/refshim.  This is exactly the drop-in
INTEGRATION.md section 1 describes ("only the imports change").

The committed fixture therefore pins, against reference CODE: the hand-offs of every coordinator
(fleet_decent_mld.py:314-455, fleet_seq_mld.py:332-431, fleet_event_based.py:460-646,
fleet_naive_admm.py:379-587, fleet_cent_mld.py:80-101), the reference env in the loop, and the
reference's Sim / leader-trajectory / spacing-policy objects.  tests/test_fleet_golden.py replays the
same runs with the REPO's coordinators + env (on the oracle backend on CPU, on libhvp.so on the GPU box,
and through the on-device Batched*Sweep classes) and compares X / U / R.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_fleet_golden.py            # writes tests/golden/fleet_golden.npz

Test infrastructure, not product."""
import contextlib
import io
import os
import pickle
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests"), "/root/reference", os.path.join(ROOT, "oracle", "refshim")):
    sys.path.insert(0, p)

import hybrid_vehicle_platoon_b200 as hvp  # noqa: E402
from hybrid_vehicle_platoon_b200 import mpc as gmpc  # noqa: E402
from oracle_backend import oracle_backend  # noqa: E402

# ---- the reference, unmodified ----------------------------------------------------------------
import fleet_cent_mld as ref_cent  # noqa: E402
import fleet_decent_mld as ref_decent  # noqa: E402
import fleet_event_based as ref_event  # noqa: E402
import fleet_naive_admm as ref_admm  # noqa: E402
import fleet_seq_mld as ref_seq  # noqa: E402
from misc.common_controller_params import Sim, Sim_n_task_2  # noqa: E402  (reference)

assert ref_decent.__file__.startswith("/root/reference/")
assert ref_decent.PlatoonEnv.__module__ == "env" and sys.modules["env"].__file__.startswith("/root/reference/")

# ---- the drop-in: re-bind the MPC class names the reference scripts instantiate ----------------
ref_cent.MpcMldCent = gmpc.MpcMldCent
ref_cent.MpcGearCent = gmpc.MpcGearCent
ref_decent.LocalMpcMld = gmpc.LocalMpcMld
ref_decent.LocalMpcGear = gmpc.LocalMpcGear
ref_seq.LocalMpcMld = gmpc.LocalMpcMld
ref_seq.LocalMpcGear = gmpc.LocalMpcGear
ref_event.LocalMpc = gmpc.EventLocalMpc
ref_event.LocalMpcGear = gmpc.EventLocalMpc
ref_admm.LocalMpcADMM = gmpc.LocalMpcADMM
ref_admm.LocalMpcGear = gmpc.LocalMpcADMM

REF = dict(cent=ref_cent, decent=ref_decent, seq=ref_seq, event=ref_event, admm=ref_admm)


def make_sim(spec):
    """spec: dict(task=0|2, n, N, ep_len, mass_seed, model)."""
    if spec["task"] == 2:
        sim = Sim_n_task_2(spec["n"], seed=spec.get("mass_seed", 0), leader_index=spec.get("leader_index"), N=spec["N"])
    else:
        sim = Sim()
        sim.n, sim.N = spec["n"], spec["N"]
        sim.id = f"default_n_{sim.n}_N_{sim.N}"
    sim.ep_len = spec["ep_len"]            # the leader trajectory keeps its 200 columns
    if "model" in spec:
        sim.vehicle_model_type = spec["model"]
    return sim


def run_reference(ctrl, spec, seed, **kw):
    """Calls the reference's simulate(sim, save=True, plot=False, ...) in a scratch directory and reads
    the 7-object pickle it writes (e.g. fleet_decent_mld.py:546-559)."""
    sim = make_sim(spec)
    li = spec.get("leader_index") or 0
    # Q7: viol_counter is a CLASS attribute shared by every env instance and the scripts pickle entry [0];
    # emptied here so that entry [0] is this run's (env.py:24-25,115)
    sys.modules["env"].PlatoonEnv.viol_counter.clear()
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.chdir(d)
        try:
            with contextlib.redirect_stdout(io.StringIO()), oracle_backend():
                REF[ctrl].simulate(sim, save=True, plot=False, seed=seed, leader_index=li, **kw)
            (f,) = os.listdir(d)
            with open(f, "rb") as fh:
                X, U, R, st, nc, viol, lx = (pickle.load(fh) for _ in range(7))
        finally:
            os.chdir(cwd)
    return dict(X=np.asarray(X, dtype=np.float64), U=np.asarray(U, dtype=np.float64), R=np.asarray(R, dtype=np.float64),
                node_counts=np.asarray(nc, dtype=np.float64), violations=np.asarray(viol, dtype=np.float64),
                leader_x=np.asarray(lx, dtype=np.float64), fname=f)


# (name, controller, sim spec, seed, extra kwargs of simulate)
D3 = dict(task=0, n=3, N=6, ep_len=100)
T5 = dict(task=2, n=5, N=6, ep_len=100)
CASES = []
for s in (1, 2, 3):
    CASES.append((f"decent_default_s{s}", "decent", D3, s, {}))
    CASES.append((f"seq_default_s{s}", "seq", D3, s, {}))
for s in (0, 1, 2):
    CASES.append((f"decent_task2_s{s}", "decent", dict(T5, mass_seed=s), s, {}))
    CASES.append((f"seq_task2_s{s}", "seq", dict(T5, mass_seed=s), s, {}))
CASES.append(("seq_task2_lead3_s0", "seq", dict(T5, mass_seed=0, leader_index=3), 0, {}))
CASES.append(("decent_task2_lead2_s1", "decent", dict(T5, mass_seed=1, leader_index=2), 1, {}))
CASES.append(("decent_default_two_point", "decent", dict(D3, ep_len=40), 1, dict(velocity_estimator="two_point")))
CASES.append(("decent_default_sat", "decent", dict(D3, ep_len=40), 1, dict(velocity_estimator="sat")))
CASES.append(("decent_friction_gear", "decent", dict(task=0, n=3, N=4, ep_len=20, model="pwa_friction"), 2, {}))
CASES.append(("seq_friction_gear", "seq", dict(task=0, n=3, N=4, ep_len=20, model="pwa_friction"), 1, {}))
# config 2 shape: n = 10, N = 6
CASES.append(("decent_task2_C2_n10_N6", "decent", dict(task=2, n=10, N=6, ep_len=60, mass_seed=3), 3, {}))
CASES.append(("seq_task2_C2_n10_N6", "seq", dict(task=2, n=10, N=6, ep_len=60, mass_seed=4), 4, {}))
# config 4 samples: long horizons
CASES.append(("decent_task2_C4_n6_N9", "decent", dict(task=2, n=6, N=9, ep_len=40, mass_seed=5), 5, {}))
CASES.append(("decent_default_C4_n5_N10", "decent", dict(task=0, n=5, N=10, ep_len=25), 6, {}))
# config 1: centralized n = 3, N = 5, 100 steps, both scripted leader trajectories
for s in (0, 1, 2):
    CASES.append((f"cent_default_C1_s{s}", "cent", dict(task=0, n=3, N=5, ep_len=100), s, {}))
    CASES.append((f"cent_task2_C1_s{s}", "cent", dict(task=2, n=3, N=5, ep_len=100, mass_seed=s), s, {}))
CASES.append(("cent_task2_lead1", "cent", dict(task=2, n=3, N=4, ep_len=60, mass_seed=1, leader_index=1), 1, {}))
CASES.append(("cent_friction_gear", "cent", dict(task=0, n=2, N=4, ep_len=20, model="pwa_friction"), 3, {}))
for s in (1, 2, 3):
    CASES.append((f"event_default_s{s}", "event", dict(task=0, n=3, N=5, ep_len=40), s, dict(event_iters=4)))
    CASES.append((f"admm_default_s{s}", "admm", dict(task=0, n=3, N=5, ep_len=30), s, dict(admm_iters=10)))
for s in (0, 1, 2):
    CASES.append((f"event_task2_s{s}", "event", dict(task=2, n=4, N=5, ep_len=60, mass_seed=s), s, dict(event_iters=3)))
    CASES.append((f"admm_task2_s{s}", "admm", dict(task=2, n=4, N=5, ep_len=60, mass_seed=s), s, dict(admm_iters=8)))
CASES.append(("event_task2_lead2_s0", "event", dict(task=2, n=5, N=4, ep_len=60, mass_seed=0, leader_index=2), 0,
              dict(event_iters=3)))
CASES.append(("admm_task2_lead1_s0", "admm", dict(task=2, n=4, N=4, ep_len=60, mass_seed=0, leader_index=1), 0,
              dict(admm_iters=5)))
CASES.append(("event_friction_gear", "event", dict(task=0, n=3, N=3, ep_len=10, model="pwa_friction"), 2, dict(event_iters=2)))
CASES.append(("admm_friction_gear", "admm", dict(task=0, n=3, N=3, ep_len=8, model="pwa_friction"), 2, dict(admm_iters=4)))
# config 3 shape: n = 15, N = 8, 20 ADMM rounds per timestep
CASES.append(("admm_task2_C3_n15_N8", "admm", dict(task=2, n=15, N=8, ep_len=4, mass_seed=7), 7, dict(admm_iters=20)))


def main(only=None):
    out, meta = {}, []
    for name, ctrl, spec, seed, kw in CASES:
        if only and only not in name:
            continue
        t = time.time()
        r = run_reference(ctrl, spec, seed, **kw)
        meta.append(dict(name=name, ctrl=ctrl, spec=spec, seed=seed, kw=kw, fname=r.pop("fname")))
        for k, v in r.items():
            out[f"{name}/{k}"] = v
        print(f"{name:30s} {time.time() - t:6.1f}s  X{r['X'].shape} return {r['R'].sum():.6g} viol {r['violations'].sum():g}",
              flush=True)
    if only is None:
        import json
        out["meta"] = np.array(json.dumps(meta))
        np.savez_compressed(os.path.join(HERE, "fleet_golden.npz"), **out)
        print("wrote fleet_golden.npz:", len(meta), "runs")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None)
