"""The MIQP half of the hot path against an INDEPENDENT formulation and an independent solver.

tests/mld_bigm.py writes the reference's MLD model out explicitly (delta, z, big-M rows: SURVEY.md 8a A1; cost and rows
of LocalMpcMld fleet_decent_mld.py:61-208 and MpcMldCent mpcs/cent_mld.py:48-177) and lets HiGHS solve it: for the
2-norm cost by brute force over ALL s^N mode sequences with the binaries fixed (HiGHS does convex QPs, not MIQPs), for
the 1-norm cost through `scipy.optimize.milp`.  Nothing is shared with the velocity-space restatement that the
oracle, the CPU port and the CUDA kernels have in common -- neither the variables (x, u, z, delta vs velocities only),
nor the search (exhaustive vs reachability-pruned / branch and bound), nor the QP method (HiGHS vs dual active set).
Big-M side effects on inactive regions are part of the model here, so agreement also shows they are inactive on these
inputs.

CPU tests pin the oracle; GPU tests pin the kernels directly (through the C ABI).  Sizes are what brute force allows:
N = 3 (343 sequences, ~1 s per problem), N = 4 (2401), centralized n = 2, N = 2 (2401 combinations)."""
import numpy as np
import pytest

import mld_bigm as MB
import gen_mpc_cases as G
from gen_cases import platoon_local_problems
from hybrid_vehicle_platoon_b200.models import Platoon

# HiGHS stops its QP at ~1e-9 in the objective; with unit curvature in u (Q_u = 1) that leaves u accurate to about
# sqrt(1e-9) ~ 3e-5 .. 1e-4 -- the bar on u is HiGHS's accuracy, not ours (objectives agree to 1e-12 typically)
TOL_OBJ, TOL_U = 1e-8, 3e-4


def _local_cases(seed, n_scen, n, N, stress, hetero, li):
    rng = np.random.default_rng(seed)
    return platoon_local_problems(rng, n_scen, n, N, li, stress, hetero)


def _bigm_local(c, i, N, d0, t0, tight=0.0):
    fl, m = int(c["flags"][i]), float(c["mass"][i])
    sysd = Platoon(1, "pwa_gear", masses=[m]).get_vehicle_system_dicts(1.0)[0]
    M, x, u, dl = MB.build_local(sysd, N, c["x0"][i], c["xf"][i], c["xb"][i], c["xl"][i], is_front=bool(fl & 1),
                                 is_leader=bool(fl & 2), is_trailer=bool(fl & 4), d0=d0, t0=t0, tight=tight)
    best, bm, second, feas, bx = MB.enumerate_miqp(M, [dl])
    return best, (None if bm is None else np.array(bm[0])), second, (None if bx is None else bx[u].ravel())


def _check_local(r, c, idx, N, d0, t0, tight=0.0):
    worst = 0.0
    for i in idx:
        best, bm, second, bu = _bigm_local(c, i, N, d0, t0, tight)
        if not np.isfinite(best):
            assert r["status"][i] == 3, (i, r["status"][i])
            continue
        assert r["status"][i] == 2
        rel = abs(r["obj"][i] - best) / max(1.0, abs(best))
        worst = max(worst, rel)
        assert rel < TOL_OBJ, (i, r["obj"][i], best)
        if second - best > 1e-6 * max(1.0, abs(best)):            # unique optimum: same modes, same inputs
            assert (r["modes"][i] == bm).all(), (i, r["modes"][i], bm)
            assert np.abs(r["u"][i] - bu).max() < TOL_U
    return worst


LOCAL_SETS = [  # (seed, scenarios, n, N, stress, hetero, leader, d0, t0, tight, problems checked)
    (1, 3, 4, 3, False, False, 0, 50.0, 0.0, 0.0, range(0, 12)),
    (2, 3, 4, 3, True, True, 2, 10.0, 3.0, 0.0, range(0, 12)),
    (3, 2, 3, 3, True, True, 1, 50.0, 0.0, 0.05, range(0, 6)),
    (4, 1, 3, 4, True, True, 0, 10.0, 3.0, 0.0, range(0, 3)),
]


@pytest.mark.parametrize("spec", LOCAL_SETS[:3], ids=["bench_N3", "stress_headway_N3", "tightening_N3"])
def test_oracle_local_miqp_matches_bigm_bruteforce(oracle, spec):
    seed, S, n, N, stress, hetero, li, d0, t0, tight, idx = spec
    c = _local_cases(seed, S, n, N, stress, hetero, li)
    r = oracle.local_miqp(N, c["flags"], c["mass"], c["x0"], c["xf"], c["xb"], c["xl"], d0=d0, t0=t0, tight=tight)
    _check_local(r, c, idx, N, d0, t0, tight)


def test_cpu_port_matches_bigm_bruteforce(oracle):
    """bench.py's CPU arm (the branch and bound with sibling bounds and warm starts) against the same model."""
    seed, S, n, N, stress, hetero, li, d0, t0, tight, idx = LOCAL_SETS[1]
    c = _local_cases(seed, S, n, N, stress, hetero, li)
    r = oracle.local_miqp_bnb(N, c["flags"], c["mass"], c["x0"], c["xf"], c["xb"], c["xl"], d0=d0, t0=t0, threads=2)
    _check_local(r, c, idx, N, d0, t0)


def _bigm_cent(x0, params, n, N, d0=50.0, t0=0.0):
    systems = Platoon(n, "pwa_gear", masses=[800.0] * n).get_vehicle_system_dicts(1.0)
    M, xs, us, ds = MB.build_cent(systems, N, x0, params.reshape(2, N + 1), d0=d0, t0=t0)
    best, bm, second, feas, bx = MB.enumerate_miqp(M, ds)
    u = None if bx is None else np.stack([bx[ui].ravel() for ui in us])
    return best, bm, second, u


def test_oracle_centralized_matches_bigm_bruteforce(oracle):
    rng = np.random.default_rng(7)
    x0, params = G.cent_cases(rng, 3, 2, 2, stress=True)
    r = oracle.mpc_solve(oracle.CENT, 2, 2, x0, 800.0, params, method=0)
    for b in range(3):
        best, bm, second, u = _bigm_cent(x0[b], params[b], 2, 2)
        assert r["status"][b] == 2 and abs(r["obj"][b] - best) / max(1.0, abs(best)) < TOL_OBJ, (r["obj"][b], best)
        if second - best > 1e-6 * max(1.0, abs(best)):
            assert (r["modes"][b].reshape(2, 2) == np.array(bm)).all()
            assert np.abs(r["u"][b].reshape(2, 2) - u).max() < TOL_U


def test_milp_one_norm_model_is_consistent_with_bruteforce():
    """The 1-norm model (quadratic_cost=False) handed to scipy.optimize.milp returns the optimum of the brute force
    over the same model's mode sequences (HiGHS LP per sequence): the MILP oracle of the 1-norm variant is sound."""
    c = _local_cases(9, 1, 3, 3, True, True, 1)
    for i in range(3):
        fl, m = int(c["flags"][i]), float(c["mass"][i])
        sysd = Platoon(1, "pwa_gear", masses=[m]).get_vehicle_system_dicts(1.0)[0]
        M, x, u, dl = MB.build_local(sysd, 3, c["x0"][i], c["xf"][i], c["xb"][i], c["xl"][i], is_front=bool(fl & 1),
                                     is_leader=bool(fl & 2), is_trailer=bool(fl & 4), quadratic=False)
        ok, xs, obj = MB.solve_milp(M)
        best, bm, second, feas, bx = MB.enumerate_miqp(M, [dl])
        assert ok and abs(obj - best) < 1e-6 * max(1.0, abs(best)), (obj, best)


# ---- the CUDA kernels against the same independent model ------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("spec", LOCAL_SETS, ids=["bench_N3", "stress_headway_N3", "tightening_N3", "stress_headway_N4"])
def test_gpu_local_miqp_matches_bigm_bruteforce(hvp_ctx, spec):
    import hybrid_vehicle_platoon_b200 as hvp
    seed, S, n, N, stress, hetero, li, d0, t0, tight, idx = spec
    c = _local_cases(seed, S, n, N, stress, hetero, li)
    r = hvp.local_miqp(N, c["flags"], c["mass"], c["x0"], c["xf"], c["xb"], c["xl"], d0=d0, t0=t0, tight=tight, ctx=hvp_ctx)
    _check_local(r, c, idx, N, d0, t0, tight)


@pytest.mark.gpu
def test_gpu_compiled_local_and_centralized_match_bigm_bruteforce(hvp_ctx):
    import hybrid_vehicle_platoon_b200 as hvp
    rng = np.random.default_rng(7)
    x0, params = G.cent_cases(rng, 3, 2, 2, stress=True)
    mpc = hvp.CompiledMpc(G.CENT, 2, n_local=2, ctx=hvp_ctx)
    r = mpc.solve(x0, 800.0, params)
    for b in range(3):
        best, bm, second, u = _bigm_cent(x0[b], params[b], 2, 2)
        assert r["status"][b] == 2 and abs(r["obj"][b] - best) / max(1.0, abs(best)) < TOL_OBJ
        if second - best > 1e-6 * max(1.0, abs(best)):
            assert (r["modes"][b].reshape(2, 2) == np.array(bm)).all()
            assert np.abs(r["u"][b].reshape(2, 2) - u).max() < TOL_U
