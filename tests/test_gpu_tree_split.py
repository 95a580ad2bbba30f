"""GPU: one MIQP tree split over several devices (include/hvp.h hvp_mpc_solve_shard_dev + dist.solve_tree_split).
The ranks are played by threads on one GPU (dist.ThreadRanks); the union of the shares must reproduce the plain
solve (and the oracle), the wave-A bound must be valid, and every rank must end with the same answer."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle as O
import gen_mpc_cases as G


def _split(hvp, mpc, x0, params, world, **kw):
    import torch
    from hybrid_vehicle_platoon_b200 import dist as D
    dev = torch.device("cuda", 0)
    B, nl = x0.shape[0], mpc.n_local
    tx0 = torch.as_tensor(np.ascontiguousarray(x0), device=dev)
    tm = torch.full((B, nl), 800.0, dtype=torch.float64, device=dev)
    tp = torch.as_tensor(np.ascontiguousarray(params), device=dev)
    outs = D.ThreadRanks(world).run(lambda r, w, ar: D.solve_tree_split(mpc, tx0, tm, tp, rank=r, world=w, allreduce=ar, **kw))
    torch.cuda.synchronize()
    return [{k: v.cpu().numpy() for k, v in o.items()} for o in outs]


@pytest.mark.parametrize("n,N,world,groups,depth", [(3, 5, 2, 8, 0), (3, 5, 4, 3, 6), (2, 4, 8, 1, 3), (4, 4, 2, 16, 8)])
def test_cent_tree_split_matches_plain_solve_and_oracle(n, N, world, groups, depth):
    import hybrid_vehicle_platoon_b200 as hvp
    rng = np.random.default_rng(900 + n * 10 + N + world)
    B = 24
    x0, params = G.cent_cases(rng, B, n, N, stress=True)
    mpc = hvp.api.CompiledMpc(G.CENT, N, n_local=n)
    plain = mpc.solve(x0, 800.0, params)
    outs = _split(hvp, mpc, x0, params, world, groups=groups, prefix_depth=depth, wave_budget=4)
    r = outs[0]
    for o in outs[1:]:                       # every rank holds the same answer
        for k in ("obj", "u", "x", "modes", "status", "winner", "nodes"):
            assert np.array_equal(o[k], r[k]), k
    assert (r["status"] == plain["status"]).all()
    ok = plain["status"] == 2
    assert np.allclose(r["obj"][ok], plain["obj"][ok], rtol=1e-9, atol=1e-9)
    assert np.abs(r["u"][ok] - plain["u"][ok]).max() < 1e-6
    assert (r["bound_after_wave_a"][ok] >= r["obj"][ok] - 1e-9 * np.abs(r["obj"][ok])).all()   # a bound, never below the optimum
    assert (r["winner"][ok] < world).all() and (r["winner"][~ok] == world).all()
    ro = O.mpc_solve(O.CENT, n, N, x0, 800.0, params, method=1)
    assert (r["status"] == ro["status"]).all()
    assert np.allclose(r["obj"][ok], ro["obj"][ok], rtol=1e-8, atol=1e-7)
    uniq = ok & (ro["second"] - ro["obj"] > 1e-6 * np.maximum(1.0, np.abs(ro["obj"])))
    assert (r["modes"][uniq] == ro["modes"][uniq]).all()


def test_shard_shares_partition_the_tree():
    """With no bound exchanged, the minimum over the ranks' shares is the optimum, and a share searched alone with
    the optimum as incumbent finds nothing better (status 3)."""
    import torch
    import hybrid_vehicle_platoon_b200 as hvp
    from hybrid_vehicle_platoon_b200 import dist as D
    rng = np.random.default_rng(77)
    n, N, B, world = 3, 4, 16, 3
    x0, params = G.cent_cases(rng, B, n, N, stress=True)
    mpc = hvp.api.CompiledMpc(G.CENT, N, n_local=n)
    plain = mpc.solve(x0, 800.0, params)
    dev = torch.device("cuda", 0)
    tx0 = torch.as_tensor(x0, device=dev); tp = torch.as_tensor(params, device=dev)
    tm = torch.full((B, n), 800.0, dtype=torch.float64, device=dev)
    objs = []
    for r in range(world):
        o = D.shard_wave(mpc, tx0, tm, tp, r, world, groups=4, prefix_depth=5, node_budget=0, incumbent=None)
        objs.append(o["obj"].cpu().numpy())
    best = np.min(np.stack(objs), axis=0)
    ok = plain["status"] == 2
    assert np.allclose(best[ok], plain["obj"][ok], rtol=1e-9, atol=1e-9) and np.isinf(best[~ok]).all()
    inc = torch.as_tensor(np.where(ok, plain["obj"], np.inf), device=dev)
    for r in range(world):
        o = D.shard_wave(mpc, tx0, tm, tp, r, world, groups=4, prefix_depth=5, node_budget=0, incumbent=inc)
        assert np.isinf(o["obj"].cpu().numpy()).all()
        assert (o["status"].cpu().numpy() == 3).all()


def test_shard_argument_checks():
    import torch
    import hybrid_vehicle_platoon_b200 as hvp
    from hybrid_vehicle_platoon_b200 import dist as D
    mpc = hvp.api.CompiledMpc(G.CENT, 3, n_local=2)
    rng = np.random.default_rng(1)
    x0, params = G.cent_cases(rng, 2, 2, 3)
    dev = torch.device("cuda", 0)
    a = (torch.as_tensor(x0, device=dev), torch.full((2, 2), 800.0, dtype=torch.float64, device=dev),
         torch.as_tensor(params, device=dev))
    with pytest.raises(RuntimeError):
        D.shard_wave(mpc, *a, 2, 2, groups=4, prefix_depth=0, node_budget=0, incumbent=None)     # rank outside world
    with pytest.raises(RuntimeError):
        D.shard_wave(mpc, *a, 0, 1, groups=0, prefix_depth=0, node_budget=0, incumbent=None)     # no groups


def test_tree_split_bench_problem_matches_oracle():
    """The size bench.py's tree-split leg runs (centralized n = 8, N = 6: 48 variables, thousands of nodes) against the
    oracle's branch and bound -- VERDICT r01: that leg was only ever checked for `optimal_frac`."""
    import hybrid_vehicle_platoon_b200 as hvp
    rng = np.random.default_rng(1234 + 9)
    n, N, B = 8, 6, 2
    x0, params = G.cent_cases(rng, B, n, N)
    mpc = hvp.api.CompiledMpc(G.CENT, N, n_local=n)
    plain = mpc.solve(x0, 800.0, params)
    outs = _split(hvp, mpc, x0, params, 2, groups=64, prefix_depth=20, wave_budget=32)
    ro = O.mpc_solve(O.CENT, n, N, x0, 800.0, params, method=1)
    assert (ro["status"] == 2).all() and (plain["status"] == 2).all() and (outs[0]["status"] == 2).all()
    for r in (plain, outs[0], outs[1]):
        assert np.allclose(r["obj"], ro["obj"], rtol=1e-8, atol=1e-7), (r["obj"], ro["obj"])
        uniq = ro["second"] - ro["obj"] > 1e-6 * np.maximum(1.0, np.abs(ro["obj"]))
        assert (r["modes"][uniq] == ro["modes"][uniq]).all()
        assert np.abs(r["u"][uniq] - ro["u"][uniq]).max() < 1e-5


@pytest.mark.gpu
def test_adoption_stress_campaign_vs_oracle():
    """The sub-tree pass with adoption (pm_kernel.cu try_donate / pm_wait_job) under the conditions that exercise it
    hardest: a node budget of 3 sends nearly every tree to the sub-tree pass, 3 static workers per tree leave most of
    the grid waiting, and no minimum size for a donated sub-tree -- every formulation of scripts/stress_parity.py
    (centralized, event-based, ADMM, gear models; random roles, horizons and spacing policies) against the oracle's
    branch and bound: statuses equal, objectives 5e-7, modes and inputs where the optimum is unique.  The knobs are read
    once per process, hence the subprocess."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, HVP_MPC_BUDGET="3", HVP_MPC_SPLIT_M="3", HVP_MPC_ADOPT_FREE="0", HVP_MPC_ADOPT="1")
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "stress_parity.py"), "11", "64"], env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "failures: 0" in r.stdout
