"""CPU: certify the oracle for the compiled-MPC formulations (centralized, event-based, naive ADMM, g-ADMM,
discrete gears).  Gurobi / dmpcpwa are absent and the reference ships no golden vectors ("parity unpinned",
DESIGN.md 3), so the pins are: (i) KKT optimality certificates and HiGHS second opinions on the fixed-mode QPs,
(ii) branch-and-bound == exhaustive enumeration, (iii) the LOCAL formulation against the first-generation local
oracle, (iv) structural identities between formulations (a centralized problem with n = 1 IS the leader's local
problem; an event problem without neighbours-of-neighbours and leader IS a 2-vehicle centralized one without
the leader term)."""
import numpy as np
import pytest

import gen_mpc_cases as G
from gen_cases import platoon_local_problems
from qp_certify import highs_qp, kkt_residuals, objective


def _certify(oracle, kind, nl, N, x0, mass, params, modes, **kw):
    q = oracle.mpc_build_qp(kind, nl, N, x0, mass, params, modes, **kw)
    if q is None:
        return None
    H, g, c0, A, b, w = q
    st, x, lam, obj, it = oracle.qp_solve(H, g, c0, A, b, w)
    ok, xh, oh = highs_qp(H, g, c0, A, b, w)
    if st != 0:
        assert not ok
        return None
    k = kkt_residuals(H, g, A, b, w, x, lam)
    assert max(k.values()) < 1e-7, k
    assert abs(objective(H, g, c0, A, b, w, x) - obj) <= 1e-9 * max(1.0, abs(obj))
    assert ok and abs(oh - obj) <= 1e-6 * max(1.0, abs(obj)), (oh, obj)      # HiGHS own accuracy ~1e-7
    return obj


@pytest.mark.parametrize("kind,nl,N,kw", [
    (G.CENT, 3, 4, {}), (G.CENT, 2, 5, dict(t0=3.0, d0=10.0, leader_index=1)),
    (G.EVENT, 3, 4, dict(n_front=2, n_behind=2, leader_index=-100)),
    (G.EVENT, 2, 5, dict(n_front=0, n_behind=2, leader_index=0, t0=3.0, d0=10.0)),
    (G.ADMM, 1, 5, dict(flags=0, rho=0.5)), (G.ADMM, 1, 4, dict(flags=G.FRONT | G.LEADER, rho=0.7)),
    (G.GADMM, 1, 5, dict(n_front=1, n_behind=1, rho=0.5)), (G.GADMM, 1, 4, dict(n_front=0, n_behind=1, flags=G.LEADER, rho=0.5)),
    (G.CENT, 2, 3, dict(model=1)), (G.LOCAL, 1, 5, dict(model=1, flags=G.TRAILER)),
])
def test_optimal_leaf_is_certified(oracle, kind, nl, N, kw):
    rng = np.random.default_rng(900 + kind * 7 + nl + N)
    B = 6
    if kind == G.CENT:
        x0, params = G.cent_cases(rng, B, nl, N, stress=True)
    elif kind == G.EVENT:
        x0, params = G.event_cases(rng, B, kw["n_front"], kw["n_behind"], N, stress=True)
    elif kind == G.ADMM:
        x0, params = G.admm_cases(rng, B, N, stress=True)
    elif kind == G.GADMM:
        x0, params, _ = G.gadmm_cases(rng, B, kw["n_front"], kw["n_behind"], N)
    else:
        full = G.platoon_states(rng, B, 3, stress=True)
        x0 = full[:, 1:2]
        params = np.concatenate([G.const_vel(a, N).reshape(B, -1) for a in (full[:, 0], full[:, 2], full[:, 1])], axis=1)
    mass = rng.uniform(700, 1000, (B, nl))
    r = oracle.mpc_solve(kind, nl, N, x0, mass, params, method=1, **kw)
    n_ok = 0
    for i in range(B):
        if r["status"][i] != 2:
            continue
        obj = _certify(oracle, kind, nl, N, x0[i], mass[i], params[i], r["modes"][i], **kw)
        assert obj is not None and abs(obj - r["obj"][i]) <= 1e-9 * max(1.0, abs(obj))
        n_ok += 1
        # a neighbouring (non-optimal) sequence is certified too and is not better
        alt = r["modes"][i].copy()
        alt[0, -1] = min(alt[0, -1] + 1, (12 if kw.get("model") else 7) - 1)
        o2 = _certify(oracle, kind, nl, N, x0[i], mass[i], params[i], alt, **kw)
        if o2 is not None:
            assert o2 >= r["obj"][i] - 1e-7 * max(1.0, abs(o2))
    assert n_ok >= B // 2


@pytest.mark.parametrize("kind,nl,N,kw", [
    (G.CENT, 2, 4, {}), (G.CENT, 3, 3, dict(t0=3.0, d0=10.0)),
    (G.EVENT, 3, 3, dict(n_front=1, n_behind=2, leader_index=1)),
    (G.ADMM, 1, 5, dict(flags=G.TRAILER, rho=0.5)), (G.CENT, 2, 3, dict(model=1)),
])
def test_branch_and_bound_equals_enumeration(oracle, kind, nl, N, kw):
    rng = np.random.default_rng(950 + kind + nl * N)
    B = 10
    if kind == G.CENT:
        x0, params = G.cent_cases(rng, B, nl, N, stress=True)
    elif kind == G.EVENT:
        x0, params = G.event_cases(rng, B, kw["n_front"], kw["n_behind"], N, stress=True)
    else:
        x0, params = G.admm_cases(rng, B, N, stress=True)
    a = oracle.mpc_solve(kind, nl, N, x0, 800.0, params, method=0, **kw)
    b = oracle.mpc_solve(kind, nl, N, x0, 800.0, params, method=1, **kw)
    assert (a["status"] == b["status"]).all()
    ok = a["status"] == 2
    assert np.allclose(a["obj"][ok], b["obj"][ok], rtol=1e-10)
    uniq = ok & (a["second"] - a["obj"] > 1e-6 * np.maximum(1.0, np.abs(a["obj"])))
    assert (a["modes"][uniq] == b["modes"][uniq]).all()
    assert np.abs(a["u"][uniq] - b["u"][uniq]).max() < 1e-7
    assert (a["leaves"] >= b["leaves"]).all()


def test_local_formulation_matches_first_oracle(oracle):
    rng = np.random.default_rng(4)
    N = 5
    c = platoon_local_problems(rng, 4, 6, N, stress=True, hetero=True)
    old = oracle.local_miqp(N, c["flags"], c["mass"], c["x0"], c["xf"], c["xb"], c["xl"], d0=10.0, t0=3.0)
    for i in range(len(c["flags"])):
        params = np.concatenate([c[k][i].ravel() for k in ("xf", "xb", "xl")])
        r = oracle.mpc_solve(G.LOCAL, 1, N, c["x0"][i][None], c["mass"][i:i + 1], params[None], flags=int(c["flags"][i]),
                             d0=10.0, t0=3.0, method=1)
        assert r["status"][0] == old["status"][i]
        if old["status"][i] == 2:
            assert abs(r["obj"][0] - old["obj"][i]) <= 1e-9 * abs(old["obj"][i])
            assert np.abs(r["u"][0, 0] - old["u"][i]).max() < 1e-7


def test_structural_identities(oracle):
    rng = np.random.default_rng(8)
    N, B = 4, 8
    # (a) centralized with one vehicle == local problem of a vehicle that is front, leader and trailer
    x0, params = G.cent_cases(rng, B, 1, N, stress=True)
    a = oracle.mpc_solve(G.CENT, 1, N, x0, 900.0, params, method=0)
    lp = np.concatenate([np.zeros((B, 4 * (N + 1))), params], axis=1)
    b = oracle.mpc_solve(G.LOCAL, 1, N, x0, 900.0, lp, flags=G.FRONT | G.LEADER | G.TRAILER, method=0)
    assert (a["status"] == b["status"]).all() and np.allclose(a["obj"], b["obj"], rtol=1e-12, equal_nan=True)
    # (b) event problem "me + one behind", leader = me  ==  centralized n = 2 with leader_index 0
    x0, params = G.cent_cases(rng, B, 2, N, stress=True)
    a = oracle.mpc_solve(G.CENT, 2, N, x0, 800.0, params, method=0)
    ep = np.concatenate([params, np.zeros((B, 4 * (N + 1)))], axis=1)
    b = oracle.mpc_solve(G.EVENT, 2, N, x0, 800.0, ep, n_front=0, n_behind=1, leader_index=0, method=0)
    assert (a["status"] == b["status"]).all() and np.allclose(a["obj"], b["obj"], rtol=1e-12, equal_nan=True)
    assert np.allclose(a["u"], b["u"], atol=1e-9)


def test_gear_mode_table(oracle):
    """12 modes = 2 friction regions x 6 gears; validity interval = friction region intersected with the gear
    window (mpc_gear.py:101-110, models.py:77-85, :276-282)."""
    a, b, c, lo, hi, gear = oracle.mode_table(1, 800.0)
    assert len(a) == 12 and list(gear) == [1, 2, 3, 4, 5, 6] * 2
    vl = np.array([3.94, 5.43, 7.56, 9.96, 13.70, 19.10]); vh = np.array([9.46, 13.04, 18.15, 23.90, 32.93, 45.84])
    alpha = 45.84 / 2
    assert np.allclose(lo[:6], vl) and np.allclose(hi[:6], np.minimum(vh, alpha))
    assert np.allclose(lo[6:], np.maximum(vl, alpha)) and np.allclose(hi[6:], vh)
    assert np.allclose(b[:6], np.array([4057, 2945, 2116, 1607, 1166, 838]) / 800.0)
    assert (lo[:3] <= hi[:3]).all() and lo[6] > hi[6]          # gear 1 never meets the upper friction region


def test_qp_tie_between_partial_and_full_step_regression(oracle):
    """A row that reaches its boundary exactly at the end of a PARTIAL dual step must join the active set with the
    multiplier it accumulated; an earlier version left the loop and lost lam_p * n_p (non-stationary 'optimum',
    a branch-and-bound miss found by scripts/stress_parity.py seed 1).  The node relaxations along the optimal
    path of that problem must be non-decreasing and agree with HiGHS."""
    rng = np.random.default_rng(1)
    B = 192
    for (n, N) in ((3, 5), (2, 6), (4, 4)):
        rng.choice([0.0, 3.0]); li = int(rng.integers(0, n))
        G.cent_cases(rng, B, n, N, stress=False, leader_index=li)
        rng.uniform(700, 1000, (B, n))
    for (nf, nb, rl, N) in ((2, 2, -100, 5), (1, 2, 1, 5), (0, 1, 0, 6)):
        nl = (nf > 0) + 1 + (nb > 0)
        x0, params = G.event_cases(rng, B, nf, nb, N, stress=False)
        mass = rng.uniform(700, 1000, (B, nl))
    w, N, nl = 158, 6, 2
    kw = dict(leader_index=0, n_front=0, n_behind=1)
    best = np.array([[3, 2, 2, 2, 2, 3], [0, 1, 1, 2, 2, 3]], np.int32)
    prev = -np.inf
    for depth in range(0, nl * N + 1):
        fm = np.full((nl, N), -1, np.int32)
        for d in range(depth):
            fm[d % nl, d // nl] = best[d % nl, d // nl]
        H, g, c0, A, b, wm = oracle.mpc_build_qp(G.EVENT, nl, N, x0[w], mass[w], params[w], fm, **kw)
        st, x, lam, obj, it = oracle.qp_solve(H, g, c0, A, b, wm)
        assert st == 0
        assert max(kkt_residuals(H, g, A, b, wm, x, lam).values()) < 1e-7
        ok, xh, oh = highs_qp(H, g, c0, A, b, wm)
        assert ok and abs(oh - obj) <= 1e-6 * abs(obj)
        assert obj >= prev - 1e-6
        prev = obj
    a = oracle.mpc_solve(G.EVENT, nl, N, x0[w][None], mass[w][None], params[w][None], method=1, **kw)
    e = oracle.mpc_solve(G.EVENT, nl, N, x0[w][None], mass[w][None], params[w][None], method=0, **kw)
    assert abs(a["obj"][0] - e["obj"][0]) <= 1e-9 * e["obj"][0] and (a["modes"] == e["modes"]).all()
