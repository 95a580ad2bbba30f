"""Dev helper: the on-device closed loops with and without CUDA-graph replay (HVP_SWEEP_GRAPH), equality + wall clock.
usage: python scripts/time_sweeps.py [decent|admm|gadmm|seq|event ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hybrid_vehicle_platoon_b200 as hvp
from hybrid_vehicle_platoon_b200 import sweep as SW
from hybrid_vehicle_platoon_b200.misc import StopAndGoLeaderTrajectory, ConstantVelocityLeaderTrajectory

which = sys.argv[1:] or ["decent", "admm", "gadmm", "seq", "event"]
ctx = hvp.Context(0)


def states(rng, S, n, vlo, vhi, glo, ghi):
    v = np.floor(rng.uniform(vlo, vhi, (S, n))); gaps = rng.uniform(glo, ghi, (S, n))
    p = np.floor(3000.0 - np.cumsum(gaps, 1) + gaps[:, :1])
    x0 = np.empty((S, 2 * n)); x0[:, 0::2] = p; x0[:, 1::2] = v
    return x0


def ab(name, make, x0, lx, T, units, small):
    res = {}
    for g in ("0", "1"):
        os.environ["HVP_SWEEP_GRAPH"] = g
        sw = make()
        sw.run(x0[:small], lx, min(T, 3))
        t0 = time.perf_counter(); out = sw.run(x0, lx, T); dt = time.perf_counter() - t0
        res[g] = (out, dt)
        print(f"{name} graph={g}: {dt*1e3:.1f} ms  {units/dt/1e6:.3f} M units/s", flush=True)
    a, b = res["0"][0], res["1"][0]
    for k in a:
        same = np.array_equal(a[k], b[k], equal_nan=True) if a[k].dtype.kind == "f" else np.array_equal(a[k], b[k])
        if not same:
            d = np.nanmax(np.abs(a[k].astype(np.float64) - b[k].astype(np.float64)))
            print(f"   {k}: differs, max |d| {d:.3e}")
    print(f"   speed-up {res['0'][1] / res['1'][1]:.2f}x", flush=True)


rng = np.random.default_rng(1237)
if "decent" in which:
    S, T, n, N = 4096, 20, 10, 6
    lx = StopAndGoLeaderTrajectory(p=3000, vh=20, vl=10, vf=30, v_change_steps=[5, 12], trajectory_len=T + N + 10, ts=1).get_leader_trajectory()
    ab("decent", lambda: SW.BatchedDecentSweep(n, N, ctx=ctx), states(rng, S, n, 5, 35, 60, 160), lx, T, S * T * n, 256)
if "seq" in which:
    S, T, n, N = 4096, 10, 10, 6
    lx = StopAndGoLeaderTrajectory(p=3000, vh=20, vl=10, vf=30, v_change_steps=[5, 12], trajectory_len=T + N + 10, ts=1).get_leader_trajectory()
    ab("seq", lambda: SW.BatchedSeqSweep(n, N, ctx=ctx), states(rng, S, n, 5, 35, 60, 160), lx, T, S * T * n, 256)
if "admm" in which:
    S, T, n, N, it = 1024, 3, 15, 8, 20
    lx = ConstantVelocityLeaderTrajectory(p=3000, v=20, trajectory_len=T + N + 10, ts=1).get_leader_trajectory()
    ab("admm", lambda: SW.BatchedAdmmSweep(n, N, admm_iters=it, rho=0.5, ctx=ctx), states(rng, S, n, 10, 30, 60, 140), lx, T, S * T * it * n, 64)
if "gadmm" in which:
    S, T, n, N, it = 1024, 2, 15, 8, 100
    lx = ConstantVelocityLeaderTrajectory(p=3000, v=20, trajectory_len=T + N + 10, ts=1).get_leader_trajectory()
    ab("gadmm", lambda: SW.BatchedGAdmmSweep(n, N, admm_iters=it, rho=0.5, ctx=ctx), states(rng, S, n, 12, 28, 60, 120), lx, T, S * it * (1 + 2 * (T - 1)) * n, 16)
if "event" in which:
    S, T, n, N = 1024, 6, 6, 5
    lx = StopAndGoLeaderTrajectory(p=3000, vh=20, vl=10, vf=30, v_change_steps=[3, 5], trajectory_len=T + N + 10, ts=1).get_leader_trajectory()
    ab("event", lambda: SW.BatchedEventSweep(n, N, event_iters=4, ctx=ctx), states(rng, S, n, 10, 30, 60, 140), lx, T, S * T * n * 4, 64)
if "gfused" in which:      # fused glue kernel vs torch glue (both graphed)
    S, T, n, N, it = 1024, 2, 15, 8, 100
    lx = ConstantVelocityLeaderTrajectory(p=3000, v=20, trajectory_len=T + N + 10, ts=1).get_leader_trajectory()
    x0 = states(rng, S, n, 12, 28, 60, 120)
    os.environ["HVP_SWEEP_GRAPH"] = "1"
    res = {}
    for fused in (False, True):
        sw = SW.BatchedGAdmmSweep(n, N, admm_iters=it, rho=0.5, ctx=ctx, fused=fused)
        sw.run(x0[:16], lx, 1)
        t0 = time.perf_counter(); out = sw.run(x0, lx, T); dt = time.perf_counter() - t0
        res[fused] = out
        print(f"gadmm fused={fused}: {dt*1e3:.1f} ms  {S*it*(1+2*(T-1))*n/dt/1e6:.3f} M QPs/s solved {out['solved'].mean():.4f}", flush=True)
    for k in res[False]:
        a, b = res[False][k], res[True][k]
        if not np.array_equal(a, b):
            print(f"   {k}: differs, max |d| {np.nanmax(np.abs(a.astype(np.float64) - b.astype(np.float64))):.3e}")
if "hint" in which:        # MIP start from the previous timestep (modes_hint) on / off
    S, T, n, N = 4096, 20, 10, 6
    lx = StopAndGoLeaderTrajectory(p=3000, vh=20, vl=10, vf=30, v_change_steps=[5, 12], trajectory_len=T + N + 10, ts=1).get_leader_trajectory()
    x0 = states(rng, S, n, 5, 35, 60, 160)
    res = {}
    for h in (False, True):
        sw = SW.BatchedDecentSweep(n, N, ctx=ctx, use_hint=h)
        sw.run(x0[:256], lx, 3)
        t0 = time.perf_counter(); out = sw.run(x0, lx, T); dt = time.perf_counter() - t0
        res[h] = out
        print(f"decent hint={h}: {dt*1e3:.1f} ms  {S*T*n/dt/1e6:.3f} M solves/s  nodes {out['nodes'].mean():.2f}  optimal {np.mean(out['status']==2):.4f}", flush=True)
    print("   max |dX|", np.abs(res[False]["X"] - res[True]["X"]).max(), " max |dU|", np.abs(res[False]["U"] - res[True]["U"]).max())
