"""GPU closed-loop parity of the centralized / ADMM / event-based / g-ADMM fleet controllers: the same
simulate() is run on the CUDA library and on the CPU oracle backend; the platoon trajectories over the
scripted (stop-and-go) leader trajectory must agree to 1e-6 (BASELINE.json asks 1e-4)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle_backend import oracle_backend
from test_host_fleets import SmallSim


def _both(fn, *a, **kw):
    import hybrid_vehicle_platoon_b200  # noqa: F401
    g = fn(*a, **kw)
    with oracle_backend():
        o = fn(*a, **kw)
    return g, o


def _agree(g, o, tol=1e-6):
    assert np.abs(g["X"] - o["X"]).max() < tol, np.abs(g["X"] - o["X"]).max()
    assert np.abs(g["U"] - o["U"]).max() < tol
    assert np.allclose(g["R"], o["R"], rtol=1e-8, atol=1e-6)
    assert (np.asarray(g["violations"]) == np.asarray(o["violations"])).all()


def test_cent_closed_loop_parity():
    import hybrid_vehicle_platoon_b200 as hvp
    g, o = _both(hvp.fleet_cent_mld.simulate, SmallSim(3, 5, 12), seed=1)
    _agree(g, o)
    assert (g["node_counts"] >= 1).all() and (g["solve_times"] > 0).all()


def test_cent_headway_hetero_parity():
    import hybrid_vehicle_platoon_b200 as hvp
    g, o = _both(hvp.fleet_cent_mld.simulate, SmallSim(3, 4, 8, headway=True, masses=[720.0, 850.0, 990.0]), seed=7)
    _agree(g, o)


def test_cent_gear_closed_loop_parity():
    import hybrid_vehicle_platoon_b200 as hvp
    g, o = _both(hvp.fleet_cent_mld.simulate, SmallSim(2, 4, 8, model="pwa_friction"), seed=3)
    _agree(g, o)


def test_decent_gear_closed_loop_parity():
    import hybrid_vehicle_platoon_b200 as hvp
    from hybrid_vehicle_platoon_b200.agents import simulate
    g, o = _both(simulate, SmallSim(3, 4, 6, model="pwa_friction"), "decent", seed=2, mpc_class=hvp.LocalMpcGear)
    _agree(g, o)


def test_admm_closed_loop_parity():
    import hybrid_vehicle_platoon_b200 as hvp
    g, o = _both(hvp.fleet_naive_admm.simulate, SmallSim(4, 4, 5, headway=True), admm_iters=5, seed=2)
    _agree(g, o)


def test_event_closed_loop_parity():
    import hybrid_vehicle_platoon_b200 as hvp
    g, o = _both(hvp.fleet_event_based.simulate, SmallSim(4, 3, 6), event_iters=3, seed=4)
    _agree(g, o)


def test_gadmm_closed_loop_parity():
    import hybrid_vehicle_platoon_b200 as hvp
    g, o = _both(hvp.fleet_g_admm.simulate, SmallSim(3, 5, 5), admm_iters=10, seed=5)
    _agree(g, o)
