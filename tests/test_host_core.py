"""CPU: the product's per-thread branch-and-bound core (csrc/miqp_core.cuh) compiled for the host
(tests/host_harness) against the oracle's exhaustive leaf enumeration.  Checks the algorithm
(velocity-space QP, bounded-dual active set, relaxed-tail bounds) without a GPU; the GPU tests
then check the same code as it actually ships (CUDA)."""
import numpy as np
import pytest

import harness
from gen_cases import platoon_local_problems


@pytest.mark.parametrize("N,stress,hetero,d0,t0,li,n_scen", [
    (6, False, False, 50.0, 0.0, 0, 40),
    (6, True, True, 10.0, 3.0, 0, 40),
    (4, True, False, 50.0, 0.0, 2, 30),
    (5, True, True, 10.0, 3.0, 9, 30),
    (8, True, True, 10.0, 3.0, 1, 6),
])
def test_core_matches_oracle(oracle, N, stress, hetero, d0, t0, li, n_scen):
    rng = np.random.default_rng(100 + N + li)
    c = platoon_local_problems(rng, n_scen, 10, N, li, stress, hetero)
    args = (N, c["flags"], c["mass"], c["x0"], c["xf"], c["xb"], c["xl"])
    ro = oracle.local_miqp(*args, d0=d0, t0=t0)
    rh = harness.local_miqp(*args, d0=d0, t0=t0)
    np.testing.assert_array_equal(rh["status"], ro["status"])
    ok = ro["status"] == 2
    np.testing.assert_allclose(rh["obj"][ok], ro["obj"][ok], rtol=1e-9)
    assert np.abs(rh["u"][ok] - ro["u"][ok]).max() < 1e-7
    assert np.abs(rh["x"][ok] - ro["x"][ok]).max() < 1e-6
    uniq = ok & ((ro["second"] - ro["obj"]) > 1e-6 * np.abs(ro["obj"]))
    np.testing.assert_array_equal(rh["modes"][uniq], ro["modes"][uniq])
    assert rh["nodes"].mean() < ro["leaves"].mean()


def test_core_infeasible(oracle):
    N = 6
    k = np.arange(N + 1)
    xl = np.stack([3000 + 20.0 * k, np.full(N + 1, 20.0)])[None]
    z = np.zeros((1, 2, N + 1))
    r = harness.local_miqp(N, np.array([7]), np.array([800.0]), np.array([[3000.0, 1.0]]), z, z, xl)
    assert r["status"][0] == 3 and np.isinf(r["obj"][0])


@pytest.mark.parametrize("N,after,t0", [(6, 1, 0.0), (6, 4, 0.0), (5, 2, 3.0), (8, 3, 0.0)])
def test_flat_subtree_adoption_partitions_the_tree(N, after, t0):
    """Tail of a flat-kernel launch (flat_core.cuh open_level / adopt_prefix), emulated on the host: an owner that
    hands open branches to fresh solvers, plus those solvers, find exactly the optimum of the plain search."""
    import harness as H
    rng = np.random.default_rng(500 + N + after)
    cs = platoon_local_problems(rng, 40, 10, N, stress=True)
    d0 = 10.0 if t0 else 50.0
    plain = H.local_miqp(N, cs["flags"], cs["mass"], cs["x0"], cs["xf"], cs["xb"], cs["xl"], d0=d0, t0=t0, flat=True)
    r = H.flat_adopt(N, cs["flags"], cs["mass"], cs["x0"], cs["xf"], cs["xb"], cs["xl"], after_nodes=after, d0=d0, t0=t0)
    ok = plain["status"] == 2
    assert ok.sum() > 300 and (r["adopters"] > 0).mean() > 0.3          # the split really happened
    assert np.isinf(r["obj"][~ok]).all()
    assert np.allclose(r["obj"][ok], plain["obj"][ok], rtol=1e-9, atol=0)


@pytest.mark.parametrize("N,M,stress,t0", [(6, 8, True, 0.0), (6, 3, False, 0.0), (8, 8, True, 3.0), (4, 5, True, 0.0)])
def test_coop_split_search_partitions_the_tree(N, M, stress, t0):
    """Latency path (csrc/local_miqp.cu coop_split_kernel): M workers search one tree -- the same root relaxation and
    first dive on every worker, the sub-trees off the dive path dealt round robin.  Emulated on the host WITHOUT
    incumbent exchange: the best worker's objective is the plain search's optimum (nothing is lost), and the longest
    worker is shorter than the plain search wherever the tree has something to share."""
    rng = np.random.default_rng(300 + N + M)
    d0 = 10.0 if t0 else 50.0
    c = platoon_local_problems(rng, 30, 10, N, 0, stress, True)
    args = (c["flags"], c["mass"], c["x0"], c["xf"], c["xb"], c["xl"])
    plain = harness.local_miqp(N, *args, d0=d0, t0=t0, coop=True)
    split = harness.coop_split(N, M, *args, d0=d0, t0=t0)
    ok = plain["status"] == 2
    assert np.isinf(split["obj"][~ok]).all()
    np.testing.assert_allclose(split["obj"][ok], plain["obj"][ok], rtol=1e-9)
    big = ok & (plain["nodes"] >= 12)
    assert big.any() and (split["nodes_max"][big] < plain["nodes"][big]).mean() > 0.9
    print(f"N={N} M={M}: plain nodes mean {plain['nodes'][ok].mean():.1f} p99 {np.percentile(plain['nodes'][ok], 99):.0f}; "
          f"longest worker mean {split['nodes_max'][ok].mean():.1f} p99 {np.percentile(split['nodes_max'][ok], 99):.0f}; "
          f"total {split['nodes_sum'][ok].mean():.1f}")
