"""Dev helper: p50/p99 latency of one scenario-timestep (10 MIQPs) through the host call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import hybrid_vehicle_platoon_b200 as hvp
from gen_cases import platoon_local_problems
ctx = hvp.Context(0)
for scen in (1, 8, 64, 512):
    one = platoon_local_problems(np.random.default_rng(77), scen, 10, 6)
    lat = []
    for i in range(700):
        t0 = time.perf_counter()
        r = hvp.local_miqp(6, one["flags"], one["mass"], one["x0"], one["xf"], one["xb"], one["xl"], ctx=ctx)
        lat.append(time.perf_counter() - t0)
    lat = np.array(lat[100:]) * 1e3
    print(f"kernel={os.environ.get('HVP_LOCAL_KERNEL','auto')} scenarios={scen} p50={np.percentile(lat,50):.3f} ms p99={np.percentile(lat,99):.3f} ms kernel_ms={ctx.last_kernel_ms():.3f}")
