"""CPU tests of the host-side fleet logic (coordinators, agents, result schema) for the centralized,
naive-ADMM, event-based and g-ADMM controllers, with the CUDA entry points swapped for the CPU oracle
(tests/oracle_backend.py).  The same closed loops run on the GPU in tests/test_gpu_fleets.py."""
import numpy as np
import pytest

import hybrid_vehicle_platoon_b200 as hvp
from hybrid_vehicle_platoon_b200.misc import (ConstantSpacingPolicy, ConstantTimePolicy, Sim,
                                              StopAndGoLeaderTrajectory)
from oracle_backend import oracle_backend


class SmallSim(Sim):
    def __init__(self, n=3, N=3, ep_len=5, model="pwa_gear", headway=False, masses=None):
        self.n, self.N, self.ep_len = n, N, ep_len
        self.vehicle_model_type = model
        self.id = f"small_n_{n}_N_{N}"
        self.masses = masses
        self.spacing_policy = ConstantTimePolicy(10, 3) if headway else ConstantSpacingPolicy(50)
        self.leader_trajectory = StopAndGoLeaderTrajectory(p=3000, vh=20, vl=14, vf=24, v_change_steps=[2, 4],
                                                           trajectory_len=ep_len + 50, ts=1)


def _check_schema(out, n, ep_len, gears=False):
    assert out["X"].shape == (ep_len + 1, 2 * n)
    assert out["U"].shape == (ep_len, 2 * n if gears else n)
    assert out["R"].shape == (ep_len, 1, 1)
    assert np.isfinite(out["X"]).all() and np.isfinite(out["R"]).all()
    assert out["leader_x"].shape == (2, ep_len + 50)
    assert (np.abs(out["U"][:, :n]) <= 1 + 1e-9).all()
    if gears:
        g = out["U"][:, n:]
        assert ((g >= 1) & (g <= 6) & (g == np.round(g))).all()


def test_cent_closed_loop_schema():
    with oracle_backend():
        out = hvp.fleet_cent_mld.simulate(SmallSim(3, 3, 5), seed=1)
    _check_schema(out, 3, 5)
    assert out["solve_times"].shape == (5, 1) and (out["node_counts"] >= 1).all()
    # first column of the state history is the reference's int-truncated reset state (Q1)
    assert (out["X"][0] == np.round(out["X"][0])).all()


def test_cent_gear_closed_loop_schema():
    with oracle_backend():
        out = hvp.fleet_cent_mld.simulate(SmallSim(2, 3, 4, model="pwa_friction"), seed=3)
    _check_schema(out, 2, 4, gears=True)


def test_admm_closed_loop_and_consensus():
    with oracle_backend():
        out = hvp.fleet_naive_admm.simulate(SmallSim(3, 3, 3, headway=True), admm_iters=4, seed=2)
    _check_schema(out, 3, 3)


def test_event_closed_loop():
    with oracle_backend():
        out = hvp.fleet_event_based.simulate(SmallSim(4, 3, 3), event_iters=2, seed=4)
    _check_schema(out, 4, 3)


def test_gadmm_closed_loop():
    with oracle_backend():
        out = hvp.fleet_g_admm.simulate(SmallSim(3, 4, 3), admm_iters=6, seed=5)
    _check_schema(out, 3, 3)
    assert out["node_counts"] == 0 and len(out["solve_times"]) == 3


def test_cent_solve_mpc_contract():
    """MpcMldCent.solve_mpc -> (u0 (n,1), info) with info['x'] (2n,N+1), info['u'] (n,N), cost, nodes; infeasible
    + raises=False -> zeros and cost = inf."""
    with oracle_backend():
        platoon = hvp.Platoon(2)
        mpc = hvp.MpcMldCent(2, 3, platoon.get_vehicle_system_dicts(1))
        mpc.set_leader_traj(np.stack([3000 + 20.0 * np.arange(4), np.full(4, 20.0)]))
        u0, info = mpc.solve_mpc(np.array([[2990.0], [18.0], [2900.0], [22.0]]))
        assert u0.shape == (2, 1) and info["x"].shape == (4, 4) and info["u"].shape == (2, 3)
        assert np.isfinite(info["cost"]) and info["nodes"] >= 1
        assert np.allclose(info["x"][:, 0], [2990, 18, 2900, 22])
        with pytest.raises(RuntimeError):
            mpc.solve_mpc(np.array([[2990.0], [70.0], [2900.0], [22.0]]))
        u0, info = mpc.solve_mpc(np.array([[2990.0], [70.0], [2900.0], [22.0]]), raises=False)
        assert (u0 == 0).all() and np.isinf(info["cost"])


def test_gear_controller_contract():
    """MpcGear.solve_mpc returns [u_g ; gears] and info['u'] = vstack(u_g, gears) (mpc_gear.py:116-135)."""
    with oracle_backend():
        sysd = hvp.PwaFrictionVehicle(900.0).get_discrete_system(1)
        mpc = hvp.LocalMpcGear(4, sysd, is_front=True, is_leader=True, is_trailer=True)
        mpc.set_leader_x(np.stack([3000 + 15.0 * np.arange(5), np.full(5, 15.0)]))
        u0, info = mpc.solve_mpc(np.array([[2995.0], [12.0]]))
        assert u0.shape == (2, 1) and info["u"].shape == (2, 4)
        assert mpc.gears_pred.shape == (1, 4) and 1 <= u0[1, 0] <= 6
        with pytest.raises(RuntimeWarning):
            mpc.solve_mpc(np.array([[2995.0], [80.0]]))
        u0, info = mpc.solve_mpc(np.array([[2995.0], [80.0]]), raises=False)
        assert u0[0, 0] == 0 and u0[1, 0] == 6 and np.isinf(info["cost"])
