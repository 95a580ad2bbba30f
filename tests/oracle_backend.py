"""Test helper: run the product's HOST classes (env, MPC, agents, coordinators) on the CPU oracle
instead of the CUDA library by swapping the two array-level entry points.  Test code only -- this
is how the closed-loop checker is built, never a product path."""
import contextlib

import numpy as np

from oracle import oracle as O
from hybrid_vehicle_platoon_b200 import api


def _rollout_step(x, u, gear=None, mass=None, leader=None, *, d0=50.0, t0=0.0, leader_index=0, d_safe=25.0,
                  quadratic=True, real_ref=False, ctx=None):
    return O.env_step(x, u, gear, mass, leader, d0, t0, leader_index, d_safe, quadratic, real_ref)


def _local_miqp(N, flags, mass, x0, xf=None, xb=None, xl=None, *, d0=50.0, t0=0.0, tight=0.0, max_nodes=0,
                ctx=None):
    r = O.local_miqp(N, flags, mass, x0, xf, xb, xl, d0=d0, t0=t0, tight=tight)
    return dict(u=r["u"], x=r["x"], modes=r["modes"], obj=r["obj"], status=r["status"],
                nodes=r["leaves"].astype(np.int32), qp_iters=np.zeros(len(r["obj"]), np.int32), run_time=0.0)


@contextlib.contextmanager
def oracle_backend():
    saved = api.rollout_step, api.local_miqp
    api.rollout_step, api.local_miqp = _rollout_step, _local_miqp
    try:
        yield
    finally:
        api.rollout_step, api.local_miqp = saved
