"""Dev helper: the closed-loop legs of bench.py in isolation (env toggles: HVP_SWEEP_GRAPH, HVP_SWEEP_FORK)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hybrid_vehicle_platoon_b200 as hvp
import bench
ctx = hvp.Context(0)
for name in [a for a in sys.argv[1:] if a != "mixed"] or ([] if "mixed" in sys.argv[1:] else ["decent", "admm", "gadmm"]):
    fn = {"decent": bench.closed_loop_leg, "admm": bench.admm_loop_leg, "gadmm": bench.gadmm_loop_leg}[name]
    for rep in range(2):
        r = fn(ctx)
        print(name, rep, f"graph={os.environ.get('HVP_SWEEP_GRAPH','1')} fork={os.environ.get('HVP_SWEEP_FORK','1')}",
              {k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items() if k in ("seconds", "solves_per_s", "qp_solves_per_s")}, flush=True)
if "mixed" in sys.argv[1:]:
    import torch
    for rep in range(2):
        r = bench.mixed_sweep_leg(ctx, 0, 1, torch.device("cuda", 0))
        print("mixed", rep, f"sibling={os.environ.get('HVP_MPC_SIBLING','1')}", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items() if k in ("seconds", "solves_per_s", "value")}, flush=True)
