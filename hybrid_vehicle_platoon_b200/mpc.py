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


class LocalMpcMld:
    """A local decentralized MPC for a single vehicle in the platoon (GPU solve)."""

    def __init__(self, N: int, pwa_system: dict, spacing_policy=ConstantSpacingPolicy(50),
                 quadratic_cost: bool = True, is_front: bool = False, is_leader: bool = False,
                 is_trailer: bool = False, thread_limit=None, accel_cnstr_tightening: float = 0.0,
                 real_vehicle_as_reference: bool = False, ctx=None) -> None:
        if not quadratic_cost:
            raise NotImplementedError("1-norm cost (MILP) is a SURVEY.md 8f row, not built yet")
        if real_vehicle_as_reference:
            raise NotImplementedError("real_vehicle_as_reference is not built on the GPU path yet")
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
    if any((m.N, m.d0, m.t0, m.tight) != (N, m0.d0, m0.t0, m0.tight) for m in mpcs):
        raise ValueError("solve_local_batch needs controllers with identical N / spacing / tightening")
    x0 = np.stack([np.asarray(s, dtype=np.float64).reshape(2) for s in states])
    r = api.local_miqp(N, np.array([m.flags for m in mpcs], np.int32), np.array([m.mass for m in mpcs]), x0,
                       np.stack([m._xf for m in mpcs]), np.stack([m._xb for m in mpcs]),
                       np.stack([m._xl for m in mpcs]), d0=m0.d0, t0=m0.t0, tight=m0.tight, ctx=ctx)
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
