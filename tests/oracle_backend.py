"""Test helper: run the product's HOST classes (env, MPC, agents, coordinators) on the CPU oracle
instead of the CUDA library by swapping the array-level entry points.  Test code only -- this
is how the closed-loop checker is built, never a product path."""
import contextlib
import itertools

import numpy as np

from oracle import oracle as O
from hybrid_vehicle_platoon_b200 import api, mpc as _mpc


def _rollout_step(x, u, gear=None, mass=None, leader=None, *, d0=50.0, t0=0.0, leader_index=0, d_safe=25.0,
                  quadratic=True, real_ref=False, ctx=None):
    return O.env_step(x, u, gear, mass, leader, d0, t0, leader_index, d_safe, quadratic, real_ref)


def _local_miqp(N, flags, mass, x0, xf=None, xb=None, xl=None, *, d0=50.0, t0=0.0, tight=0.0, max_nodes=0,
                mip_gap=0.0, time_limit_ms=0.0, ctx=None):
    r = O.local_miqp(N, flags, mass, x0, xf, xb, xl, d0=d0, t0=t0, tight=tight)
    return dict(u=r["u"], x=r["x"], modes=r["modes"], obj=r["obj"], status=r["status"],
                nodes=r["leaves"].astype(np.int32), qp_iters=np.zeros(len(r["obj"]), np.int32), run_time=0.0)


def oracle_eval_cost(kind, nl, N, mass, params, xg, ug, tol=1e-6, **kw):
    """Independent restatement of eval_cost (fleet_event_based.py:308-327) on the oracle's dense
    input-space QP: modes from consistency of the pinned data, objective + L1 penalties at u."""
    mass = np.broadcast_to(np.asarray(mass, dtype=np.float64), (nl,))
    tabs = [O.mode_table(kw.get("model", 0), mass[i]) for i in range(nl)]
    cands = []
    for i in range(nl):
        a, b, c, lo, hi, _ = tabs[i]
        row = []
        for k in range(N):
            v = xg[i, 1, k]
            ok = [r for r in range(len(a)) if lo[r] - tol <= v <= hi[r] + tol and
                  (k == N - 1 or abs(xg[i, 1, k + 1] - (a[r] * v + b[r] * ug[i, k] + c[r])) <= tol)]
            if k < N - 1 and abs(xg[i, 0, k + 1] - xg[i, 0, k] - v) > tol:
                return np.inf
            if not ok:
                return np.inf
            row.append(ok if k == N - 1 else ok[:1])
        cands.append(row)
    best = np.inf
    for last in itertools.product(*[cands[i][N - 1][:4] for i in range(nl)]):
        modes = np.array([[cands[i][k][0] for k in range(N - 1)] + [last[i]] for i in range(nl)], np.int32)
        q = O.mpc_build_qp(kind, nl, N, xg[:, :, 0], mass, params, modes, **kw)
        if q is None:
            continue
        H, g, c0, A, bb, w = q
        z = ug.reshape(-1)
        s = A @ z - bb
        hard = ~np.isfinite(w)
        if (s[hard] > tol).any():
            continue
        f = 0.5 * z @ H @ z + g @ z + c0 + (w[~hard] * np.maximum(s[~hard], 0)).sum()
        best = min(best, f)
    return best


class _OracleCompiledMpc:
    """api.CompiledMpc with the same attributes, solved by oracle/hvp_oracle_mpc.c."""

    def __init__(self, kind, N, *, n_local=1, model=0, flags=0, leader_index=0, n_front=0, n_behind=0, d0=50.0,
                 t0=0.0, tight=0.0, rho=0.5, max_nodes=0, one_norm=False, mip_gap=0.0, time_limit_ms=0.0, ctx=None):
        assert not one_norm, "the C oracle restates the 2-norm formulations; the 1-norm oracle is tests/mld_bigm.py (milp)"
        self.kind, self.N, self.n_local, self.model = kind, N, n_local, model
        self.kw = dict(model=model, flags=flags, leader_index=leader_index, n_front=n_front, n_behind=n_behind,
                       d0=d0, t0=t0, tight=tight, rho=rho)
        self.n_param, self.n_extra = O.mpc_dims(kind, n_local, N, flags, n_front, n_behind)
        self.mode_gear = O.mode_table(model)[5]
        self.n_modes = len(self.mode_gear)

    def solve(self, x0, mass, params, fixed_modes=None):
        nl = self.n_local
        x0 = np.asarray(x0, dtype=np.float64).reshape(-1, nl, 2)
        B = x0.shape[0]
        r = O.mpc_solve(self.kind, nl, self.N, x0, mass, params, fixed_modes=fixed_modes, method=1, **self.kw)
        return dict(u=r["u"], x=r["x"], extra=r["extra"], modes=r["modes"], obj=r["obj"], status=r["status"],
                    nodes=np.maximum(r["nodes"], r["leaves"]).astype(np.int32), qp_iters=np.zeros(B, np.int32),
                    run_time=0.0)

    def eval_cost(self, mass, params, xg, ug):
        nl, N = self.n_local, self.N
        xg = np.asarray(xg, dtype=np.float64).reshape(-1, nl, 2, N + 1)
        B = xg.shape[0]
        ug = np.asarray(ug, dtype=np.float64).reshape(B, nl, N)
        mass = np.broadcast_to(np.asarray(mass, dtype=np.float64), (B, nl))
        params = np.asarray(params, dtype=np.float64).reshape(B, -1)
        return np.array([oracle_eval_cost(self.kind, nl, N, mass[i], params[i], xg[i], ug[i], **self.kw)
                         for i in range(B)])

    def gears(self, modes):
        return self.mode_gear[np.clip(modes, 0, self.n_modes - 1)]


@contextlib.contextmanager
def oracle_backend():
    saved = api.rollout_step, api.local_miqp, api.CompiledMpc, dict(_mpc._handles)
    api.rollout_step, api.local_miqp, api.CompiledMpc = _rollout_step, _local_miqp, _OracleCompiledMpc
    _mpc._handles.clear()
    try:
        yield
    finally:
        api.rollout_step, api.local_miqp, api.CompiledMpc = saved[:3]
        _mpc._handles.clear()
        _mpc._handles.update(saved[3])
