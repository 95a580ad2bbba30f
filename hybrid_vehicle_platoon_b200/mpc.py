"""GPU-backed MPC controllers with the reference's solve interface.

LocalMpcMld mirrors fleet_decent_mld.py:21-223 / fleet_seq_mld.py:21-234 (constructor argument order,
set_x_front / set_x_back / set_leader_x, solve_mpc(state, raises) -> (u0, info)); the mixed-integer
solve that the reference delegates to Gurobi runs in the CUDA branch-and-bound kernel
(csrc/local_miqp.cu).  solve_local_batch() solves the MIQPs of many controllers in ONE launch --
the way the decentralized coordinator batches across vehicles."""
from __future__ import annotations

import numpy as np

from . import api
from ._lib import FRONT, LEADER, OPTIMAL, TRAILER
from .misc import ConstantSpacingPolicy, spacing_params
from .models import mass_of_pwa_system


class _Val:
    """Gurobi-style value accessor (`mpc.x.X`, fleet_naive_admm.py:413-447)."""

    def __init__(self, shape):
        self.X = np.zeros(shape)
        self.shape = shape


class SolverOptions:
    """Solver parameters every controller snapshots when it is constructed -- the analogue of the Gurobi parameters the
    reference sets on its model (mpcs/mpc_gear.py:182-186) or leaves at their defaults.
    mip_gap: relative gap at which a node is pruned (Gurobi MIPGap, 1e-4 there); 0 = proven optimal (default here).
    time_limit: seconds per problem (Gurobi TimeLimit); when it runs out the solve ends with status 9 (TIME_LIMIT),
    which solve_mpc treats like every status other than 2: raise, or zeros + inf cost.  0 = none."""
    mip_gap = 0.0
    time_limit = 0.0


OPTIONS = SolverOptions()


def set_solver_options(mip_gap=None, time_limit=None):
    if mip_gap is not None:
        OPTIONS.mip_gap = float(mip_gap)
    if time_limit is not None:
        OPTIONS.time_limit = float(time_limit)


class LocalMpcMld:
    """A local decentralized MPC for a single vehicle in the platoon (GPU solve)."""

    def __new__(cls, N, pwa_system, spacing_policy=ConstantSpacingPolicy(50), quadratic_cost=True, is_front=False,
                is_leader=False, is_trailer=False, thread_limit=None, accel_cnstr_tightening=0.0,
                real_vehicle_as_reference=False, ctx=None):
        # real_vehicle_as_reference (fleet_seq_mld.py:137,211-219; False in every Sim of the reference) adds a spacing
        # term and a safe-distance row against the leader trajectory: that variant lives in the compiled LOCAL
        # formulation (csrc/pm_build.cu), which solves the same problem as this class otherwise
        # quadratic_cost=False (the 1-norm / MILP variant, fleet_decent_mld.py:75-78) lives there too: LP node problems
        # through the proximal-point wrapper of the compiled-MPC kernel (csrc/pm_types.h)
        if (real_vehicle_as_reference or not quadratic_cost) and cls is LocalMpcMld:
            return LocalMpcGear(N, pwa_system, spacing_policy, quadratic_cost, is_front, is_leader, is_trailer,
                                thread_limit, accel_cnstr_tightening, real_vehicle_as_reference, ctx=ctx)
        return super().__new__(cls)

    def __init__(self, N: int, pwa_system: dict, spacing_policy=ConstantSpacingPolicy(50),
                 quadratic_cost: bool = True, is_front: bool = False, is_leader: bool = False,
                 is_trailer: bool = False, thread_limit=None, accel_cnstr_tightening: float = 0.0,
                 real_vehicle_as_reference: bool = False, ctx=None) -> None:
        self.N, self.n, self.m = N, 1, 1
        self.mass = mass_of_pwa_system(pwa_system)
        self.d0, self.t0 = spacing_params(spacing_policy)
        self.tight = float(accel_cnstr_tightening)
        self.flags = (FRONT if is_front else 0) | (LEADER if is_leader else 0) | (TRAILER if is_trailer else 0)
        self.is_front, self.is_leader, self.is_trailer = is_front, is_leader, is_trailer
        self._xf = np.zeros((2, N + 1)); self._xb = np.zeros((2, N + 1)); self._xl = np.zeros((2, N + 1))
        self.x, self.u = _Val((2, N + 1)), _Val((1, N))
        self.num_bin_vars = 7 * N                     # delta (s, N) of the MLD model
        self._ctx = ctx
        self.mip_gap, self.time_limit_ms = OPTIONS.mip_gap, OPTIONS.time_limit * 1e3

    # parameter setters (fleet_decent_mld.py:210-223)
    def set_leader_x(self, leader_x):
        self._xl = np.array(leader_x, dtype=np.float64).reshape(2, self.N + 1)

    def set_x_front(self, x_front):
        self._xf = np.array(x_front, dtype=np.float64).reshape(2, self.N + 1)

    def set_x_back(self, x_back):
        self._xb = np.array(x_back, dtype=np.float64).reshape(2, self.N + 1)

    def solve_mpc(self, state, raises: bool = True):
        return solve_local_batch([self], [state], raises=raises, ctx=self._ctx)[0]


def solve_local_batch(mpcs, states, raises: bool = True, ctx=None):
    """Solve the local MIQPs of `mpcs` (same N, spacing policy and tightening) for `states` in one
    kernel launch.  Returns [(u0, info), ...] exactly as each mpc.solve_mpc(state) would."""
    m0 = mpcs[0]
    N = m0.N
    if any((m.N, m.d0, m.t0, m.tight, m.mip_gap, m.time_limit_ms) !=
           (N, m0.d0, m0.t0, m0.tight, m0.mip_gap, m0.time_limit_ms) for m in mpcs):
        raise ValueError("solve_local_batch needs controllers with identical N / spacing / tightening / solver options")
    x0 = np.stack([np.asarray(s, dtype=np.float64).reshape(2) for s in states])
    r = api.local_miqp(N, np.array([m.flags for m in mpcs], np.int32), np.array([m.mass for m in mpcs]), x0,
                       np.stack([m._xf for m in mpcs]), np.stack([m._xb for m in mpcs]),
                       np.stack([m._xl for m in mpcs]), d0=m0.d0, t0=m0.t0, tight=m0.tight, mip_gap=m0.mip_gap,
                       time_limit_ms=m0.time_limit_ms, ctx=ctx)
    run_time = r["run_time"]
    out = []
    for i, m in enumerate(mpcs):
        if r["status"][i] == OPTIMAL:
            u, x, cost = r["u"][i].reshape(1, N), r["x"][i], float(r["obj"][i])
        else:
            if raises:
                raise RuntimeError(f"Infeasible problem encountered (status {int(r['status'][i])}).")
            u, x, cost = np.zeros((1, N)), np.zeros((2, N + 1)), float("inf")
        m.x.X, m.u.X = x, u
        info = {"x": x, "u": u, "cost": cost, "run_time": run_time, "nodes": int(r["nodes"][i]),
                "bin_vars": m.num_bin_vars, "status": int(r["status"][i]), "modes": r["modes"][i]}
        out.append((u[:, [0]], info))
    return out


# ==============================================================================================
# Controllers solved by the compiled-MPC kernel (csrc/pm_kernel.cu through hvp_mpc_*)
# ==============================================================================================
from ._lib import (MODEL_FRICTION_GEAR, MPC_ADMM, MPC_CENT, MPC_EVENT, MPC_GADMM, MPC_LOCAL, NO_LEADER,  # noqa: E402
                   REAL_VEHICLE_REF)
from .models import model_of_pwa_system  # noqa: E402

_handles: dict = {}


def _handle(key, ctx):
    """One device-resident compiled formulation per distinct (kind, model, sizes, policy) -- controllers
    with the same structure (e.g. all interior vehicles of a platoon) share it and batch together."""
    k = (key, id(ctx))
    if k not in _handles:
        kind, model, nl, N, flags, li, nf, nb, d0, t0, tight, rho, one_norm, gap, tl = key
        _handles[k] = api.CompiledMpc(kind, N, n_local=nl, model=model, flags=flags, leader_index=li, n_front=nf,
                                      n_behind=nb, d0=d0, t0=t0, tight=tight, rho=rho, one_norm=one_norm,
                                      mip_gap=gap, time_limit_ms=tl, ctx=ctx)
    return _handles[k]


class _CompiledController:
    """Common part of the controllers below: parameter blocks, the solve_mpc contract of dmpcpwa's
    MpcMld and of MpcGear.solve_mpc (mpcs/mpc_gear.py:116-135)."""

    def _setup(self, kind, N, systems, spacing_policy, quadratic_cost, *, flags=0, leader_index=0, n_front=0,
               n_behind=0, tight=0.0, rho=0.5, ctx=None):
        if not quadratic_cost and kind not in (MPC_CENT, MPC_LOCAL):
            raise NotImplementedError("the 1-norm cost (MILP) is built for the centralized and the per-vehicle local "
                                      "controllers (SURVEY.md 8f rank 3); event-based / ADMM variants are not")
        mm = [model_of_pwa_system(sy) for sy in systems]
        if len({m for m, _ in mm}) != 1:
            raise ValueError("all vehicles of one controller must use the same model type")
        self.model = mm[0][0]
        self.masses = np.array([m for _, m in mm], dtype=np.float64)
        self.N, self.nl = int(N), len(systems)
        self.n = self.nl                      # the reference's attribute names: n vehicles, m inputs
        self.m = self.nl
        self.discrete_gears = self.model == MODEL_FRICTION_GEAR
        d0, t0 = spacing_params(spacing_policy)
        self._key = (kind, self.model, self.nl, self.N, int(flags), int(leader_index), int(n_front), int(n_behind),
                     d0, t0, float(tight), float(rho), not quadratic_cost, OPTIONS.mip_gap, OPTIONS.time_limit * 1e3)
        self._ctx = ctx
        self._cm = _handle(self._key, ctx)
        self._blk = 2 * (self.N + 1)
        self._params = np.zeros(self._cm.n_param)
        self.x, self.u = _Val((2 * self.nl, self.N + 1)), _Val((self.nl, self.N))
        if self.discrete_gears:
            self.u_g = _Val((self.nl, self.N))
            self.gears_pred = 6 * np.ones((self.nl, self.N))
        # binaries of the MLD model: delta (s, N) per subsystem (+ sigma (6, nu, N) with gears)
        self.num_bin_vars = (8 if self.discrete_gears else 7) * self.N * self.nl
        self._fixed_modes = None

    def _set_block(self, b, arr):
        self._params[b * self._blk:(b + 1) * self._blk] = np.asarray(arr, dtype=np.float64).reshape(-1)

    def _x0(self, state):
        return np.asarray(state, dtype=np.float64).reshape(self.nl, 2)

    def solve_mpc(self, state, raises: bool = True):
        return solve_compiled_batch([self], [state], raises=raises)[0]

    def _post(self, r, i):
        """Hook for subclasses to store extra decision variables (copies)."""


def solve_compiled_batch(mpcs, states, raises: bool = True):
    """Solve the MIQPs of `mpcs` for `states`, one kernel launch per distinct formulation.
    Returns [(u0, info), ...] exactly as each mpc.solve_mpc(state, raises) would."""
    groups: dict = {}
    for i, m in enumerate(mpcs):
        groups.setdefault((m._key, id(m._ctx), m._fixed_modes is not None), []).append(i)
    out = [None] * len(mpcs)
    for idx in groups.values():
        cm = mpcs[idx[0]]._cm
        fm = None
        if mpcs[idx[0]]._fixed_modes is not None:
            fm = np.stack([mpcs[i]._fixed_modes for i in idx])
        r = cm.solve(np.stack([mpcs[i]._x0(states[i]) for i in idx]), np.stack([mpcs[i].masses for i in idx]),
                     np.stack([mpcs[i]._params for i in idx]), fixed_modes=fm)
        for j, i in enumerate(idx):
            m = mpcs[i]
            N, nl = m.N, m.nl
            ok = r["status"][j] == OPTIMAL
            if ok:
                u, x, cost = r["u"][j].reshape(nl, N), r["x"][j].reshape(2 * nl, N + 1), float(r["obj"][j])
                gears = cm.gears(r["modes"][j]).astype(np.float64).reshape(nl, N)
            else:
                if raises:
                    if m.discrete_gears:      # mpcs/mpc_gear.py:127-128
                        raise RuntimeWarning(f"gear mpc for state {states[i]} is infeasible.")
                    raise RuntimeError(f"Infeasible problem encountered (status {int(r['status'][j])}).")
                u, x, cost = np.zeros((nl, N)), np.zeros((2 * nl, N + 1)), float("inf")
                gears = 6 * np.ones((nl, N))  # mpcs/mpc_gear.py:129-131
            m.x.X, m.u.X = x, u
            info = {"x": x, "u": u, "cost": cost, "run_time": r["run_time"], "nodes": int(r["nodes"][j]),
                    "bin_vars": m.num_bin_vars, "status": int(r["status"][j]), "modes": r["modes"][j]}
            m._post(r, j)
            if m.discrete_gears:              # [u_g ; gears]  (mpcs/mpc_gear.py:133-135)
                m.u_g.X, m.gears_pred = u, gears
                info["u"] = np.vstack((u, gears))
                out[i] = (np.vstack((u[:, [0]], gears[:, [0]])), info)
            else:
                out[i] = (u[:, [0]], info)
    return out


class MpcMldCent(_CompiledController):
    """Centralized platoon MPC (mpcs/cent_mld.py:9-182; with a pwa_friction system list it is
    MpcGearCent, fleet_cent_mld.py:25-52)."""

    def __init__(self, n: int, N: int, pwa_systems, spacing_policy=ConstantSpacingPolicy(50), leader_index: int = 0,
                 quadratic_cost: bool = True, thread_limit=None, accel_cnstr_tightening: float = 0.0,
                 real_vehicle_as_reference: bool = False, ctx=None) -> None:
        if len(pwa_systems) != n:
            raise ValueError(f"expected {n} systems, got {len(pwa_systems)}")
        if leader_index != 0 and real_vehicle_as_reference:
            raise NotImplementedError("Not implemented for real vehicle with leader not 0.")   # cent_mld.py:63-66
        self._setup(MPC_CENT, N, pwa_systems, spacing_policy, quadratic_cost,
                    flags=REAL_VEHICLE_REF if real_vehicle_as_reference else 0, leader_index=leader_index,
                    tight=accel_cnstr_tightening, ctx=ctx)

    def set_leader_traj(self, leader_traj):                   # cent_mld.py:179-182
        self._set_block(0, leader_traj)


MpcGearCent = MpcMldCent      # the model type of the system dicts selects the gear formulation


class LocalMpcGear(_CompiledController):
    """Per-vehicle local MPC with discrete gears (fleet_decent_mld.py:226-253 / fleet_seq_mld.py:237-264);
    also accepts a pwa_gear system, for which it solves the same problem as LocalMpcMld."""

    def __init__(self, N: int, system: dict, spacing_policy=ConstantSpacingPolicy(50), quadratic_cost: bool = True,
                 is_front: bool = False, is_leader: bool = False, is_trailer: bool = False, thread_limit=None,
                 accel_cnstr_tightening: float = 0.0, real_vehicle_as_reference: bool = False, ctx=None) -> None:
        flags = (FRONT if is_front else 0) | (LEADER if is_leader else 0) | (TRAILER if is_trailer else 0) \
            | (REAL_VEHICLE_REF if real_vehicle_as_reference else 0)
        self._setup(MPC_LOCAL, N, [system], spacing_policy, quadratic_cost, flags=flags, tight=accel_cnstr_tightening,
                    ctx=ctx)

    def set_x_front(self, x_front):
        self._set_block(0, x_front)

    def set_x_back(self, x_back):
        self._set_block(1, x_back)

    def set_leader_x(self, leader_x):
        self._set_block(2, leader_x)


class EventLocalMpc(_CompiledController):
    """Event-based local MPC over the vehicle and up to one neighbour each side
    (fleet_event_based.py:26-376: LocalMpc / LocalMpcGear).  Local state = [x_front, x_me, x_back]."""

    def __init__(self, N: int, systems, num_vehicles_in_front: int, num_vehicles_behind: int,
                 spacing_policy=ConstantSpacingPolicy(50), rel_leader_index=None, quadratic_cost: bool = True,
                 thread_limit=None, accel_cnstr_tightening: float = 0.0, ctx=None) -> None:
        if rel_leader_index is not None and rel_leader_index not in (-1, 0, 1):
            raise ValueError(f"rel leader index must be -1, 0, or 1. Got {rel_leader_index}.")
        self._setup(MPC_EVENT, N, systems, spacing_policy, quadratic_cost,
                    leader_index=NO_LEADER if rel_leader_index is None else rel_leader_index,
                    n_front=num_vehicles_in_front, n_behind=num_vehicles_behind, tight=accel_cnstr_tightening, ctx=ctx)

    def set_leader_x(self, leader_x):
        self._set_block(0, leader_x)

    def set_x_f2(self, x_f2):
        self._set_block(1, x_f2)

    def set_x_b2(self, x_b2):
        self._set_block(2, x_b2)

    def solve_mpc(self, state, raises: bool = False):          # fleet_event_based.py:329 (raises defaults False)
        return solve_compiled_batch([self], [state], raises=raises)[0]

    def eval_cost(self, x, u, discrete_gears: bool = False) -> float:
        """Cost of a given (x, u) guess with x[:, :N] and u pinned (fleet_event_based.py:308-327);
        inf if the guess is not feasible for the MLD model."""
        return float(eval_compiled_batch([self], [x], [u])[0])


def eval_compiled_batch(mpcs, xs, us):
    """eval_cost of many event-based controllers, one launch per distinct formulation."""
    groups: dict = {}
    for i, m in enumerate(mpcs):
        groups.setdefault((m._key, id(m._ctx)), []).append(i)
    out = np.empty(len(mpcs))
    for idx in groups.values():
        cm = mpcs[idx[0]]._cm
        xg = np.stack([np.asarray(xs[i], dtype=np.float64).reshape(mpcs[i].nl, 2, mpcs[i].N + 1) for i in idx])
        ug = np.stack([np.asarray(us[i], dtype=np.float64).reshape(mpcs[i].nl, mpcs[i].N) for i in idx])
        c = cm.eval_cost(np.stack([mpcs[i].masses for i in idx]), np.stack([mpcs[i]._params for i in idx]), xg, ug)
        out[idx] = c
    return out


class LocalMpcADMM(_CompiledController):
    """Local MPC of the naive (non-convex) ADMM scheme with free copies of the neighbours' states
    (fleet_naive_admm.py:24-290: LocalMpcADMM / LocalMpcGear)."""

    def __init__(self, N: int, pwa_system: dict, rho: float, spacing_policy=ConstantSpacingPolicy(50),
                 quadratic_cost: bool = True, is_front: bool = False, is_leader: bool = False,
                 is_trailer: bool = False, thread_limit=None, accel_cnstr_tightening: float = 0.0, ctx=None) -> None:
        flags = (FRONT if is_front else 0) | (LEADER if is_leader else 0) | (TRAILER if is_trailer else 0)
        self.rho = rho
        self.is_front, self.is_trailer = is_front, is_trailer
        self._setup(MPC_ADMM, N, [pwa_system], spacing_policy, quadratic_cost, flags=flags,
                    tight=accel_cnstr_tightening, rho=rho, ctx=ctx)
        if not is_front:
            self.x_front = _Val((2, N + 1))
        if not is_trailer:
            self.x_back = _Val((2, N + 1))

    def set_leader_x(self, leader_x):
        self._set_block(0, leader_x)

    def set_front_vars(self, y_front, z_front):                # fleet_naive_admm.py:239-245
        self._set_block(1, y_front); self._set_block(2, z_front)

    def set_back_vars(self, y_back, z_back):                   # fleet_naive_admm.py:247-253
        self._set_block(3, y_back); self._set_block(4, z_back)

    def _post(self, r, j):
        e, N, o = r["extra"][j], self.N, 0
        ok = r["status"][j] == OPTIMAL
        if not self.is_front:
            self.x_front.X = e[o:o + self._blk].reshape(2, N + 1) if ok else np.zeros((2, N + 1))
            o += self._blk
        if not self.is_trailer:
            self.x_back.X = e[o:o + self._blk].reshape(2, N + 1) if ok else np.zeros((2, N + 1))


class GAdmmLocalMpc(_CompiledController):
    """Fixed-sequence local QP of the switching ADMM scheme (fleet_g_admm.py:22-205 LocalMpc on dmpcpwa
    MpcSwitching / dmpcrl MpcAdmm).  The augmented state is [copies in front..., own state, copies behind...];
    y, z are (2 * n_aug, N+1).  set_sequence fixes the PWA region per stage; solve() returns the QP solution."""

    rho = 0.5

    def __init__(self, N: int, pwa_system: dict, num_neighbours: int, my_index: int,
                 spacing_policy=ConstantSpacingPolicy(50), leader: bool = False, rho: float = 0.5, ctx=None) -> None:
        self.rho = rho
        self.my_index, self.num_neighbours, self.leader = my_index, num_neighbours, leader
        self._setup(MPC_GADMM, N, [pwa_system], spacing_policy, True, flags=LEADER if leader else 0,
                    n_front=my_index, n_behind=num_neighbours - my_index, rho=rho, ctx=ctx)
        self.n_aug = num_neighbours + 1
        self._fixed_modes = np.zeros((1, N), np.int32)
        self.x_c = _Val((2 * num_neighbours, N + 1))

    def set_sequence(self, seq):
        self._fixed_modes = np.asarray(seq, dtype=np.int32).reshape(1, self.N)

    def set_leader_traj(self, x_ref):
        self._set_block(0, x_ref)

    def set_consensus(self, y, z):
        na = self.n_aug
        self._params[self._blk:self._blk * (1 + na)] = np.asarray(y, dtype=np.float64).reshape(-1)
        self._params[self._blk * (1 + na):self._blk * (1 + 2 * na)] = np.asarray(z, dtype=np.float64).reshape(-1)

    def _post(self, r, j):
        ok = r["status"][j] == OPTIMAL
        self.x_c.X = r["extra"][j].reshape(2 * self.num_neighbours, self.N + 1) if ok else np.zeros(self.x_c.shape)

    def augmented(self):
        """Augmented state (2 n_aug, N+1) of the last solve, in G-map order."""
        mi = self.my_index
        return np.vstack((self.x_c.X[:2 * mi], self.x.X, self.x_c.X[2 * mi:]))
