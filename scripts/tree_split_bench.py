"""One LARGE centralized MIQP tree searched by all GPUs of the box (SURVEY.md 8e "Collective").
Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
             scripts/tree_split_bench.py --vehicles 8 --horizon 6 --problems 4
(or plain `python scripts/tree_split_bench.py` for one GPU).  Every rank holds the same problems; the incumbent
bound is exchanged with an NCCL allreduce(min); times are CUDA events, max over ranks.  Prints one JSON line."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--vehicles", type=int, default=8)
    ap.add_argument("--horizon", type=int, default=6)
    ap.add_argument("--problems", type=int, default=4)
    ap.add_argument("--groups", type=int, default=256)
    ap.add_argument("--depth", type=int, default=0)
    ap.add_argument("--wave-budget", type=int, default=16)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--seed", type=int, default=5)
    ap.add_argument("--plain", action="store_true", help="also time the single-device solve (rank 0)")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import hybrid_vehicle_platoon_b200 as hvp
    from hybrid_vehicle_platoon_b200 import dist as D
    import gen_mpc_cases as G
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ctx = hvp.Context(local)
    rng = np.random.default_rng(a.seed)
    x0, params = G.cent_cases(rng, a.problems, a.vehicles, a.horizon, stress=False)
    mpc = hvp.api.CompiledMpc(G.CENT, a.horizon, n_local=a.vehicles, ctx=ctx)
    tx0 = torch.as_tensor(x0, device=dev); tp = torch.as_tensor(params, device=dev)
    tm = torch.full((a.problems, a.vehicles), 800.0, dtype=torch.float64, device=dev)
    times = []
    out = None
    for rep in range(a.reps + 1):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = D.solve_tree_split(mpc, tx0, tm, tp, groups=a.groups, prefix_depth=a.depth, wave_budget=a.wave_budget)
        e1.record()
        torch.cuda.synchronize()
        t = D.max_over_ranks([e0.elapsed_time(e1)], device=dev)[0]
        if rep > 0:
            times.append(t)
    plain = None
    if a.plain and rank == 0:
        r = mpc.solve(x0, 800.0, params)
        t0 = r["run_time"] * 1e3
        r = mpc.solve(x0, 800.0, params)
        plain = dict(ms=r["run_time"] * 1e3, first_ms=t0, obj=r["obj"].tolist(), nodes=r["nodes"].tolist(),
                     status=r["status"].tolist())
    if rank == 0:
        print(json.dumps(dict(what="tree split of centralized MIQPs over GPUs", n=a.vehicles, N=a.horizon, variables=a.vehicles * a.horizon,
                              problems=a.problems, world=world, groups_per_gpu=a.groups, wave_budget=a.wave_budget,
                              ms=float(np.median(times)), ms_all=times, obj=out["obj"].cpu().tolist(),
                              bound_after_wave_a=out["bound_after_wave_a"].cpu().tolist(),
                              status=out["status"].cpu().tolist(), nodes=out["nodes"].cpu().tolist(),
                              winner=out["winner"].cpu().tolist(), plain=plain)))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
