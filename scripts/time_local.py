"""Dev helper: time the per-vehicle local MIQP kernel on the bench batch (device-resident, CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import hybrid_vehicle_platoon_b200 as hvp
from hybrid_vehicle_platoon_b200 import api
from gen_cases import platoon_local_problems

S = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
N, n = 6, 10
b = platoon_local_problems(np.random.default_rng(1235), S, n, N)
B = S * n
dev = torch.device("cuda", 0); ctx = hvp.Context(0)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
d = {k: t(v) for k, v in b.items()}
u = torch.empty((B, N), dtype=torch.float64, device=dev); x = torch.empty((B, 2, N + 1), dtype=torch.float64, device=dev)
mo = torch.empty((B, N), dtype=torch.int32, device=dev); ob = torch.empty(B, dtype=torch.float64, device=dev)
st = torch.empty(B, dtype=torch.int32, device=dev); no = torch.empty(B, dtype=torch.int32, device=dev)
it = torch.empty(B, dtype=torch.int32, device=dev)
desc = api.local_desc(N); stream = torch.cuda.current_stream().cuda_stream
ms = []
for i in range(8):
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    api.local_miqp_device(desc, B, d["flags"], d["mass"], d["x0"], d["xf"], d["xb"], d["xl"], u, x, mo, ob, st, no, it, ctx=ctx, stream=stream)
    e.record(); torch.cuda.synchronize()
    if i >= 3: ms.append(a.elapsed_time(e))
print(f"NODE_BATCH={os.environ.get('HVP_NODE_BATCH','dflt')} HULL={os.environ.get('HVP_FLAT_HULL','dflt')} ms={np.mean(ms):.3f} solves/s={B/np.mean(ms)*1e3:.0f} "
      f"nodes={no.double().mean().item():.2f} iters={it.double().mean().item():.2f} ok={(st==2).all().item()} objsum={ob.sum().item():.6f}")
