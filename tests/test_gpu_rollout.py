"""GPU: rollout/cost kernel through the C ABI vs (i) the reference-generated golden vectors and
(ii) the oracle on seeded inputs, plus size-independent properties at bench size."""
import numpy as np
import pytest

import hybrid_vehicle_platoon_b200 as hvp

pytestmark = pytest.mark.gpu


def _states(rng, B, n):
    v = rng.uniform(5.0, 34.0, (B, n))
    gaps = rng.uniform(15.0, 160.0, (B, n))
    p = 3000.0 - np.cumsum(gaps, 1) + gaps[:, :1]
    x = np.empty((B, 2 * n)); x[:, 0::2] = p; x[:, 1::2] = v
    return x, rng.uniform(-1, 1, (B, n)), np.stack([p[:, 0] + rng.uniform(-30, 30, B), rng.uniform(10, 30, B)], 1)


def test_golden_reference_steps(hvp_ctx, golden):
    g = golden
    nbit = ntot = 0
    for ci, (n, het, d0, t0, given, li, quad, rr) in enumerate(g["cfg_meta"]):
        x, u, gear, mass, L = (g[f"c{ci}_{k}"] for k in ("x", "u", "gear", "mass", "leader"))
        xo, c, v, e = hvp.rollout_step(x, u, gear if given else None, mass if het else None, L, d0=d0,
                                       t0=t0, leader_index=int(li), quadratic=bool(quad),
                                       real_ref=bool(rr), ctx=hvp_ctx)
        assert (e == 0).all()
        np.testing.assert_array_equal(v, g[f"c{ci}_viol"])
        # float tolerance: same operation order as numpy, no FMA -> 1e-15 relative (<= 1 ulp)
        np.testing.assert_allclose(xo, g[f"c{ci}_x_new"], rtol=1e-15, atol=0)
        np.testing.assert_allclose(c, g[f"c{ci}_r"], rtol=1e-15, atol=0)
        nbit += (xo == g[f"c{ci}_x_new"]).sum() + (c == g[f"c{ci}_r"]).sum()
        ntot += xo.size + c.size
    assert nbit >= ntot - 3


def test_golden_kats_and_sequence(hvp_ctx, golden):
    g = golden
    for tag in ("kat0", "kat1"):
        x = g[f"{tag}_x"][0:1]
        for t in range(2):
            xo, c, v, e = hvp.rollout_step(x, np.array([[0.3, -0.2, 0.1]]), None, None,
                                           np.array([[3000.0 + 20.0 * t, 20.0]]), ctx=hvp_ctx)
            np.testing.assert_array_equal(xo[0], g[f"{tag}_x"][t + 1])
            assert c[0] == g[f"{tag}_r"][t]
            x = xo
    xo, c, v, e = hvp.rollout_step(g["task2_x0"][None], np.array([[0.5, -0.5, 0.25, 1.0]]),
                                   np.array([[4, 4, 2, 4]], np.int32), g["task2_masses"],
                                   g["task2_leader"][:, 0][None], d0=10.0, t0=3.0, ctx=hvp_ctx)
    np.testing.assert_array_equal(xo[0], g["task2_x1"])
    np.testing.assert_allclose(c[0], g["task2_r"][0], rtol=1e-15)
    x = g["seq_x"][0:1]
    for t in range(40):   # closed sequence: chain the GPU's own states, compare trajectory
        xo, c, v, e = hvp.rollout_step(x, g["seq_u"][t][None], None, g["seq_masses"],
                                       g["seq_leader"][:, t][None], d0=10.0, t0=3.0, ctx=hvp_ctx)
        assert e[0] == 0 and v[0] == (g["seq_viol"][t] == 100)
        np.testing.assert_allclose(xo[0], g["seq_x"][t + 1], rtol=1e-14, atol=0)
        np.testing.assert_allclose(c[0], g["seq_r"][t], rtol=1e-14)
        x = xo


def test_error_codes_match_reference(hvp_ctx, golden):
    for (v, j, u), k in zip(golden["err_in"], golden["err_kind"]):
        xo, _, _, e = hvp.rollout_step(np.array([[100.0, v]]), np.array([[u]]),
                                       np.array([[int(j)]], np.int32), None, np.zeros((1, 2)), ctx=hvp_ctx)
        assert (e[0] & 255) == k
        assert np.isnan(xo).all()


@pytest.mark.parametrize("n,B,per_mass,given", [(10, 5000, True, True), (3, 777, False, False),
                                                 (15, 1001, True, False), (1, 300, False, False),
                                                 (64, 130, True, True)])
def test_oracle_parity_random(hvp_ctx, oracle, n, B, per_mass, given):
    rng = np.random.default_rng(n * 1000 + B)
    x, u, leader = _states(rng, B, n)
    mass = rng.uniform(700, 1000, (B, n)) if per_mass else None
    gear = None
    if given:
        gear = np.array([[oracle.gear_from_velocity(v) for v in row] for row in x[:, 1::2]], np.int32)
        gear = np.clip(gear + rng.integers(-1, 2, gear.shape), 0, 7).astype(np.int32)  # some invalid
    li = int(rng.integers(0, n))
    out = hvp.rollout_step(x, u, gear, mass, leader, d0=10.0, t0=3.0, leader_index=li, ctx=hvp_ctx)
    ref = oracle.env_step(x, u, gear, mass, leader, 10.0, 3.0, li)
    np.testing.assert_array_equal(out[3], ref[3])          # error codes (first exception)
    np.testing.assert_array_equal(out[2], ref[2])          # violations
    ok = ref[3] == 0
    np.testing.assert_array_equal(out[0][ok], ref[0][ok])  # bit-exact states
    np.testing.assert_array_equal(out[1], ref[1])          # bit-exact costs
    assert np.isnan(out[0][~ok]).all()


def test_empty_and_bad_args(hvp_ctx):
    xo, c, v, e = hvp.rollout_step(np.zeros((0, 6)), np.zeros((0, 3)), None, None, np.zeros((0, 2)), ctx=hvp_ctx)
    assert xo.shape == (0, 6) and c.shape == (0,)
    with pytest.raises(RuntimeError, match="leader_index"):
        hvp.rollout_step(np.ones((2, 6)) * 10, np.zeros((2, 3)), None, None, np.zeros((2, 2)),
                         leader_index=5, ctx=hvp_ctx)


def test_full_size_properties(hvp_ctx, oracle):
    """C5 size (1M scenario-steps, n=10): translation invariance of the dynamics in p, cost
    independence of later vehicles, and agreement with the oracle on a random subset."""
    rng = np.random.default_rng(5)
    B, n = 1 << 20, 10
    x, u, leader = _states(rng, B, n)
    a = hvp.rollout_step(x, u, None, None, leader, ctx=hvp_ctx)
    x2 = x.copy(); x2[:, 0::2] += 128.0
    l2 = leader.copy(); l2[:, 0] += 128.0
    b = hvp.rollout_step(x2, u, None, None, l2, ctx=hvp_ctx)
    ok = a[3] == 0
    assert ok.mean() > 0.9
    np.testing.assert_array_equal(a[0][ok][:, 1::2], b[0][ok][:, 1::2])     # velocities identical
    np.testing.assert_allclose(b[0][ok][:, 0::2] - 128.0, a[0][ok][:, 0::2], rtol=0, atol=1e-9)
    np.testing.assert_array_equal(a[2], b[2])
    idx = rng.choice(B, 4096, replace=False)
    ref = oracle.env_step(x[idx], u[idx], None, None, leader[idx])
    np.testing.assert_array_equal(a[3][idx], ref[3])
    okk = ref[3] == 0
    np.testing.assert_array_equal(a[0][idx][okk], ref[0][okk])
    np.testing.assert_array_equal(a[1][idx], ref[1])
