"""CPU: the host-side mirror of the reference interface (env reset quirks, model tables, agent
loop, wrappers, coordinators) -- with the array-level entry points served by the oracle."""
import numpy as np
import pytest

import hybrid_vehicle_platoon_b200 as hvp
from hybrid_vehicle_platoon_b200.models import mass_of_pwa_system
from oracle_backend import oracle_backend


def test_model_tables_match_reference(golden):
    for key, m in (("m800", 800.0), ("m914", 914.5568099117258)):
        s = hvp.PwaGearVehicle(m).get_discrete_system(1.0)
        got = np.vstack([[A[1, 1] for A in s["A"]], [B[1, 0] for B in s["B"]], [c[1, 0] for c in s["c"]]])
        np.testing.assert_array_equal(got, golden[f"pwa_{key}"][:3])
        np.testing.assert_array_equal(np.concatenate([s["E"].ravel(), s["G"].ravel()]), golden[f"pwa_{key}_DEFG"])
        assert mass_of_pwa_system(s) == pytest.approx(m, rel=1e-14)
    v = hvp.PwaGearVehicle(800)
    np.testing.assert_array_equal(v.v_gear_lim, golden["v_gear_lim"])
    assert [v.get_gear_from_velocity(x) for x in golden["gearmap_in"][:8]] == list(golden["gearmap_out"][:8])
    np.testing.assert_allclose([v.get_u_for_constant_vel(x) for x in golden["constvel_in"]], golden["constvel_out"], rtol=1e-15)
    np.testing.assert_array_equal(hvp.Sim_n_task_2(4, seed=0).leader_trajectory.get_leader_trajectory(), golden["task2_leader"])
    np.testing.assert_array_equal(hvp.Sim().leader_trajectory.get_leader_trajectory(), golden["const_leader"])
    np.testing.assert_array_equal(hvp.Sim_n_task_2(4, seed=0).masses, golden["task2_masses"])


def test_env_reset_quirks(golden):
    """Q1 (int64 truncated state) and Q2 (legacy global RNG, 100 + 99 draws) -- env.py:70-116."""
    for n in (3, 5, 10, 15):
        for s in range(4):
            seed = int(np.random.SeedSequence(s).generate_state(1)[0])
            env = hvp.PlatoonEnv(n, hvp.Platoon(n, "pwa_gear"), 150)
            x0, info = env.reset(seed=seed)
            assert x0.dtype == np.int64 and x0.shape == (2 * n, 1) and info == {}
            np.testing.assert_array_equal(x0.ravel(), golden[f"reset_n{n}"][s])
    for tag, pol in (("const", hvp.ConstantSpacingPolicy(50)), ("headway", hvp.ConstantTimePolicy(10, 3))):
        env = hvp.PlatoonEnv(4, hvp.Platoon(4, "pwa_gear"), 150, spacing_policy=pol, start_from_platoon=True)
        x0, _ = env.reset(seed=1)
        np.testing.assert_array_equal(x0.ravel(), golden[f"reset_platoon_{tag}"])


def test_env_step_interface(golden):
    with oracle_backend():
        env = hvp.PlatoonEnv(3, hvp.Platoon(3, "pwa_gear"), 150)
        env.reset(seed=2968811710)
        for t in range(2):
            x, r, term, trunc, info = env.step(np.array([[0.3], [-0.2], [0.1]]))
            assert x.shape == (6, 1) and r.shape == (1, 1) and term is False and trunc is False
            np.testing.assert_array_equal(x.ravel(), golden["kat0_x"][t + 1])
            assert r[0, 0] == golden["kat0_r"][t]
        assert env.step_counter == 2
        with pytest.raises(ValueError):
            env.step(np.zeros((4, 1)))
        with pytest.raises(RuntimeError):          # gear 1 at 21 m/s: outside the traction curve (Q6)
            env.step(np.array([[0.1], [0.1], [0.1], [1], [1], [1]]))


@pytest.mark.parametrize("controller", ["decent", "seq"])
def test_closed_loop_runs_on_oracle(controller):
    with oracle_backend():
        out = hvp.simulate(hvp.Sim(), controller=controller, seed=1, ep_len=30)
    assert out["X"].shape == (31, 6) and out["U"].shape == (30, 3) and out["R"].shape == (30, 1, 1)
    assert out["solve_times"].shape == (30, 1) and out["leader_x"].shape[0] == 2
    assert np.isfinite(out["X"]).all()
    # platoon converges towards the leader: tracking cost decreases
    assert out["R"][-1, 0, 0] < out["R"][0, 0, 0]


def test_agent_prediction_shift():
    a = hvp.MldAgent(None)
    assert a.get_predicted_state(True) is None
    a.x_pred = np.arange(8.0).reshape(2, 4)
    np.testing.assert_array_equal(a.get_predicted_state(True), [[1, 2, 3, 3], [5, 6, 7, 7]])
