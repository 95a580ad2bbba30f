"""GPU: closed-loop platoon trajectories over the scripted leader trajectories
(misc/leader_trajectory.py), GPU controllers + GPU env vs the same host loop on the oracle.
Tolerance: BASELINE.json's 1e-4 on trajectories (asserted much tighter)."""
import numpy as np
import pytest

import hybrid_vehicle_platoon_b200 as hvp
from oracle_backend import oracle_backend

pytestmark = pytest.mark.gpu


def _compare(sim, controller, seed, ep_len, leader_index=0, **kw):
    gpu = hvp.simulate(sim, controller=controller, seed=seed, ep_len=ep_len, leader_index=leader_index, **kw)
    with oracle_backend():
        ref = hvp.simulate(sim, controller=controller, seed=seed, ep_len=ep_len, leader_index=leader_index, **kw)
    dX = np.abs(gpu["X"] - ref["X"]).max()
    dU = np.abs(gpu["U"] - ref["U"]).max()
    assert dX < 1e-6 and dU < 1e-6, (dX, dU)
    np.testing.assert_allclose(gpu["R"], ref["R"], rtol=1e-8)
    np.testing.assert_array_equal(gpu["violations"], ref["violations"])
    return gpu, ref


@pytest.mark.parametrize("controller", ["decent", "seq"])
def test_default_sim_constant_leader(hvp_ctx, controller):
    gpu, _ = _compare(hvp.Sim(), controller, seed=2, ep_len=60)
    assert gpu["R"][-1, 0, 0] < gpu["R"][0, 0, 0]
    assert (gpu["node_counts"][:60] > 0).all() and (gpu["solve_times"][:60] > 0).all()


@pytest.mark.parametrize("controller,li", [("decent", 0), ("seq", 0), ("seq", 3)])
def test_task2_stop_and_go(hvp_ctx, controller, li):
    """Sim_n_task_2: heterogeneous masses, time-headway spacing, stop-and-go leader 20->10->30."""
    _compare(hvp.Sim_n_task_2(4, seed=0, leader_index=li), controller, seed=0, ep_len=70, leader_index=li)


def test_decent_velocity_estimators(hvp_ctx):
    for est in ("two_point", "sat"):
        _compare(hvp.Sim(), "decent", seed=1, ep_len=25, velocity_estimator=est)


def test_batched_env_matches_single(hvp_ctx):
    import torch
    rng = np.random.default_rng(0)
    n, B = 5, 64
    env = hvp.BatchedPlatoonEnv(n, masses=rng.uniform(700, 1000, (B, n)), spacing_policy=hvp.ConstantTimePolicy(10, 3))
    v = rng.uniform(6, 33, (B, n)); p = 3000 - np.cumsum(rng.uniform(30, 120, (B, n)), 1)
    x = np.empty((B, 2 * n)); x[:, 0::2] = p; x[:, 1::2] = v
    u = rng.uniform(-1, 1, (B, n)); leader = np.stack([p[:, 0], np.full(B, 20.0)], 1)
    dev = lambda a: torch.from_numpy(a).cuda()
    xo, c, vi, e = env.step(dev(x), dev(u), dev(leader))
    torch.cuda.synchronize()
    ref = hvp.rollout_step(x, u, None, env.mass.cpu().numpy(), leader, d0=10.0, t0=3.0, ctx=hvp_ctx)
    np.testing.assert_array_equal(xo.cpu().numpy(), ref[0])
    np.testing.assert_array_equal(c.cpu().numpy(), ref[1])
    np.testing.assert_array_equal(e.cpu().numpy(), ref[3])
