"""Dev helper: the compiled-MPC legs of bench.py (cent, ADMM local, g-ADMM QP, 1-norm) and the closed-loop ADMM leg in
isolation, for A/B runs over the library's environment knobs (HVP_MPC_SPLIT_M, HVP_MPC_BUDGET, HVP_MPC_ADOPT)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hybrid_vehicle_platoon_b200 as hvp
import bench
dev = torch.device("cuda", 0)
ctx = hvp.Context(0)
flushbuf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
legs = bench.compiled_mpc_legs(hvp, torch, dev, torch.cuda.current_stream().cuda_stream, flushbuf, steps=3, with_cpu=False)
for k, v in legs.items():
    print(k[:28], round(v["value"]), "nodes %.1f" % v["nodes_per_solve"], "opt %.4f" % v["optimal_frac"], flush=True)
if "admm" in sys.argv[1:]:
    r = bench.admm_loop_leg(ctx)
    print("closed_loop_admm", round(r["solves_per_s"]), r["seconds_runs"])
