"""Dev helper: latency of one scenario-timestep (10 MIQPs) through the host call, a different scenario per call:
wall time of the call next to the device time of the kernel inside it."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import hybrid_vehicle_platoon_b200 as hvp
from gen_cases import platoon_local_problems
ctx = hvp.Context(0)
n, N = 10, 6
pool = platoon_local_problems(np.random.default_rng(77), 512, n, N)
sl = lambda a, j, k: a[j * n:(j + k) * n]
for scen in (1, 4):
    lat, ker, nodes = [], [], []
    for i in range(2200):
        j = (i * scen) % (512 - scen)
        t0 = time.perf_counter()
        r = hvp.local_miqp(N, sl(pool["flags"], j, scen), sl(pool["mass"], j, scen), sl(pool["x0"], j, scen),
                           sl(pool["xf"], j, scen), sl(pool["xb"], j, scen), sl(pool["xl"], j, scen), ctx=ctx)
        lat.append(time.perf_counter() - t0); ker.append(ctx.last_kernel_ms()); nodes.append(r["nodes"].max())
    lat = np.array(lat[200:]) * 1e3; ker = np.array(ker[200:]); nodes = np.array(nodes[200:])
    print(f"splitM={os.environ.get('HVP_COOP_SPLIT_M','dflt')} kernel={os.environ.get('HVP_LOCAL_KERNEL','auto')} scenarios={scen} wall p50={np.percentile(lat,50):.3f} p99={np.percentile(lat,99):.3f} ms | "
          f"kernel p50={np.percentile(ker,50):.3f} p99={np.percentile(ker,99):.3f} ms | worst tree p50={np.percentile(nodes,50):.0f} p99={np.percentile(nodes,99):.0f} nodes")
