"""Dev helper: bench.py's tree-split leg alone (one GPU, or under torchrun), over groups per problem / prefix depth."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hybrid_vehicle_platoon_b200 as hvp
import bench
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
ctx = hvp.Context(local)
CASES = [tuple(int(v) for v in c.split(":")) for c in os.environ.get("CASES", "8:6:20,10:6:30").split(",")]
for n, N, depth in CASES:
    for groups in (int(g) for g in os.environ.get("NGROUPS", "256").split(",")):
        r = bench.tree_split_leg(ctx, dev, n=n, N=N, depth=depth, groups=groups, problems=int(os.environ.get('PROBLEMS', '4')), stress=bool(int(os.environ.get('STRESS', '0'))), seed=int(os.environ.get('SEED', '5')))
        if rank == 0:
            print(f"n={n} N={N} depth={depth} groups={groups} world={world}: {r['ms']:.1f} ms  nodes/solve {r['nodes_per_solve']:.0f}  opt {r['optimal_frac']}", flush=True)
if world > 1:
    dist.destroy_process_group()
