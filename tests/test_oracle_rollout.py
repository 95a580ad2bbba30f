"""CPU: the C oracle's rollout/cost restatement against golden vectors produced by the reference's
own env.py/models.py (tests/golden/make_rollout_golden.py).  Pins the oracle."""
import numpy as np


def _run_cfg(O, g, ci):
    n, het, d0, t0, given, li, quad, rr = g["cfg_meta"][ci]
    x, u, gear, mass, L = (g[f"c{ci}_{k}"] for k in ("x", "u", "gear", "mass", "leader"))
    out = O.env_step(x, u, gear if given else None, mass if het else None, L, d0, t0, int(li), 25.0,
                     bool(quad), bool(rr))
    return out, g[f"c{ci}_x_new"], g[f"c{ci}_r"], g[f"c{ci}_viol"], gear


def test_random_steps_match_reference(oracle, golden):
    nbit = ntot = 0
    for ci in range(len(golden["cfg_meta"])):
        (xo, c, v, e), xr, rr, vr, gear = _run_cfg(oracle, golden, ci)
        assert (e == 0).all()
        assert (v == vr).all()
        # floating point: reference order of operations reproduced; tolerance 1e-15 relative
        # (one value in 12k differs by 1 ulp: numpy's pow/BLAS path), bit-exact otherwise
        np.testing.assert_allclose(xo, xr, rtol=1e-15, atol=0)
        np.testing.assert_allclose(c, rr, rtol=1e-15, atol=0)
        nbit += (xo == xr).sum() + (c == rr).sum()
        ntot += xo.size + c.size
    assert nbit >= ntot - 3


def test_derived_gears(oracle, golden):
    for v, gg in zip(golden["gearmap_in"], golden["gearmap_out"]):
        assert oracle.gear_from_velocity(v) == gg
    for ci in range(len(golden["cfg_meta"])):
        if golden["cfg_meta"][ci][4] == 0:
            x, gear = golden[f"c{ci}_x"], golden[f"c{ci}_gear"]
            got = np.array([[oracle.gear_from_velocity(v) for v in row[1::2]] for row in x])
            assert (got == gear).all()


def test_kats(oracle, golden):
    g = golden
    assert list(g["seedseq"]) == [2968811710, 1835504127, 2834126987, 1576890651]
    for tag in ("kat0", "kat1"):
        xs, rs = g[f"{tag}_x"], g[f"{tag}_r"]
        x = xs[0:1]
        for t in range(2):
            u = np.array([[0.3, -0.2, 0.1]])
            leader = np.array([[3000.0 + 20.0 * t, 20.0]])
            xo, c, v, e = oracle.env_step(x, u, None, None, leader)
            np.testing.assert_array_equal(xo[0], xs[t + 1])
            assert c[0] == rs[t]
            x = xo
    # SURVEY.md 4.3 literal values
    np.testing.assert_array_equal(g["kat0_x"][0], [3000, 21, 2931, 22, 2808, 10])
    assert abs(g["kat0_r"][0] - 5704.74) < 1e-9 and abs(g["kat0_r"][1] - 7502.87189225) < 1e-7
    # task-2: heterogeneous masses + explicit gears + headway spacing
    xo, c, v, e = oracle.env_step(g["task2_x0"][None], np.array([[0.5, -0.5, 0.25, 1.0]]),
                                  np.array([[4, 4, 2, 4]], np.int32), g["task2_masses"],
                                  g["task2_leader"][:, 0][None], 10.0, 3.0)
    np.testing.assert_array_equal(xo[0], g["task2_x1"])
    np.testing.assert_allclose(c[0], g["task2_r"][0], rtol=1e-15)


def test_tables(oracle, golden):
    for (v, j), t in zip(golden["traction_in"], golden["traction_out"]):
        assert oracle.traction(v, int(j)) == (t, 0)
    for key, m in (("m800", 800.0), ("m914", 914.5568099117258)):
        a, b, c, lo, hi = oracle.pwa_gear_system(m)
        ref = golden[f"pwa_{key}"]
        np.testing.assert_array_equal(np.vstack([a, b, c, lo, hi]), ref)


def test_error_codes(oracle, golden):
    for (v, j, u), k in zip(golden["err_in"], golden["err_kind"]):
        _, _, _, e = oracle.env_step(np.array([[100.0, v]]), np.array([[u]]),
                                     np.array([[int(j)]], np.int32), None, np.zeros((1, 2)))
        assert (e[0] & 255) == k


def test_closed_sequence(oracle, golden):
    g = golden
    x = g["seq_x"][0:1]
    for t in range(40):
        xo, c, v, e = oracle.env_step(x, g["seq_u"][t][None], None, g["seq_masses"],
                                      g["seq_leader"][:, t][None], 10.0, 3.0)
        assert e[0] == 0
        np.testing.assert_allclose(xo[0], g["seq_x"][t + 1], rtol=1e-15, atol=0)
        np.testing.assert_allclose(c[0], g["seq_r"][t], rtol=1e-15)
        assert v[0] == (g["seq_viol"][t] == 100)
        x = g["seq_x"][t + 1][None]   # re-anchor on the reference's state
