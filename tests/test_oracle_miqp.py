"""CPU: certify the oracle's MIQP side.  The reference's solver (Gurobi via dmpcpwa) is absent and
ships no golden vectors, so the pins are (i) KKT optimality certificates of every leaf QP,
(ii) HiGHS on the explicit-slack formulation, (iii) exhaustive vs reachability-pruned
enumeration.  "parity unpinned" w.r.t. Gurobi itself."""
import numpy as np
import pytest

from gen_cases import platoon_local_problems
from qp_certify import highs_qp, kkt_residuals, objective


def _cases(seed, n_scen, N, **kw):
    rng = np.random.default_rng(seed)
    return platoon_local_problems(rng, n_scen, 10, N, **kw)


@pytest.mark.parametrize("N,stress,t0,d0", [(6, False, 0.0, 50.0), (6, True, 3.0, 10.0), (4, True, 0.0, 50.0)])
def test_leaf_qp_kkt_certificate(oracle, N, stress, t0, d0):
    c = _cases(11 + N, 12, N, stress=stress, hetero=True)
    r = oracle.local_miqp(N, c["flags"], c["mass"], c["x0"], c["xf"], c["xb"], c["xl"], d0=d0, t0=t0)
    assert (r["status"] == 2).all()
    rng = np.random.default_rng(5)
    for i in range(len(r["obj"])):
        modes = r["modes"][i].copy()
        if i % 2:   # also certify a non-optimal leaf
            k = rng.integers(1, N)
            modes[k:] = np.clip(modes[k:] + rng.choice([-1, 1]), 0, 6)
        H, g, c0, A, b, w = oracle.local_build_qp(N, c["flags"][i], c["mass"][i], c["x0"][i], c["xf"][i],
                                                  c["xb"][i], c["xl"][i], modes, d0=d0, t0=t0)
        st, x, lam, obj, it = oracle.qp_solve(H, g, c0, A, b, w)
        if st != 0:
            ok, _, _ = highs_qp(H, g, c0, A, b, w)
            assert not ok
            continue
        k = kkt_residuals(H, g, A, b, w, x, lam)
        assert max(k.values()) < 1e-8, k
        assert abs(objective(H, g, c0, A, b, w, x) - obj) <= 1e-10 * abs(obj)
        if i % 2 == 0:
            assert abs(obj - r["obj"][i]) <= 1e-10 * abs(obj)
        if i % 5 == 0:
            ok, xh, oh = highs_qp(H, g, c0, A, b, w)
            assert ok and abs(oh - obj) <= 1e-9 * abs(obj) and np.abs(xh - x).max() < 1e-5


def test_pruned_equals_exhaustive(oracle):
    c = _cases(3, 6, 4, stress=True)
    a = oracle.local_miqp(4, c["flags"], c["mass"], c["x0"], c["xf"], c["xb"], c["xl"], exhaustive=True)
    b = oracle.local_miqp(4, c["flags"], c["mass"], c["x0"], c["xf"], c["xb"], c["xl"])
    assert (a["status"] == b["status"]).all()
    np.testing.assert_array_equal(a["obj"], b["obj"])
    np.testing.assert_array_equal(a["modes"], b["modes"])
    assert (a["leaves"] >= b["leaves"]).all()


def test_infeasible_and_boundary(oracle):
    N = 6
    k = np.arange(N + 1)
    xl = np.stack([3000 + 20.0 * k, np.full(N + 1, 20.0)])[None]
    # v0 so low that v_1 >= 3.94 is unreachable within the acceleration limit -> infeasible
    r = oracle.local_miqp(N, 7, 800.0, np.array([[3000.0, 1.0]]), None, None, xl)
    assert r["status"][0] == 3 and np.isinf(r["obj"][0])
    # v0 exactly on a region edge (closed on both sides, models.py:430-444): both regions are
    # admissible for stage 0 and the optimum is the better of the two
    edge = oracle.pwa_gear_system(800.0)[4][2]
    r = oracle.local_miqp(N, 7, 800.0, np.array([[3000.0, edge]]), None, None, xl)
    lo = oracle.local_miqp(N, 7, 800.0, np.array([[3000.0, edge - 1e-9]]), None, None, xl)
    hi = oracle.local_miqp(N, 7, 800.0, np.array([[3000.0, edge + 1e-9]]), None, None, xl)
    assert r["status"][0] == 2 and r["modes"][0, 0] in (2, 3)
    assert lo["modes"][0, 0] == 2 and hi["modes"][0, 0] == 3
    assert abs(r["obj"][0] - min(lo["obj"][0], hi["obj"][0])) < 1e-4


def test_cpu_bnb_port_matches_enumeration(oracle):
    """bench.py's CPU arm (oracle/hvp_cpu_bnb.cpp: the product's branch-and-bound compiled for the host) returns the
    optimum of the independent exhaustive enumeration on the bench distribution and on the stress distribution."""
    from gen_cases import platoon_local_problems
    rng = np.random.default_rng(11)
    for N, stress, d0, t0 in ((6, False, 50.0, 0.0), (6, True, 10.0, 3.0), (4, True, 50.0, 0.0), (8, True, 10.0, 3.0)):
        c = platoon_local_problems(rng, 40, 10, N, 2, stress, True)
        a = (N, c["flags"], c["mass"], c["x0"], c["xf"], c["xb"], c["xl"])
        r = oracle.local_miqp_bnb(*a, d0=d0, t0=t0, threads=2)
        ro = oracle.local_miqp(*a, d0=d0, t0=t0)
        assert (r["status"] == ro["status"]).all()
        ok = ro["status"] == 2
        rel = np.abs(r["obj"][ok] - ro["obj"][ok]) / np.maximum(1.0, np.abs(ro["obj"][ok]))
        assert rel.max() < 1e-7, rel.max()
        assert r["nodes"][ok].mean() < ro["leaves"][ok].mean()
