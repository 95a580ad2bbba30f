"""Dev experiment: does ORDERING the batch so that a warp's 32 problems are alike raise lane utilisation?
usage: python scripts/time_local_sorted.py S mode   (mode: none | role | role_v | role_v_err)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hybrid_vehicle_platoon_b200 as hvp
from hybrid_vehicle_platoon_b200 import api
from hybrid_vehicle_platoon_b200.synth_local import platoon_local_problems, EDGES

S = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
mode = sys.argv[2] if len(sys.argv) > 2 else "none"
N, n = 6, 10
b = platoon_local_problems(np.random.default_rng(1235), S, n, N)
B = S * n
if mode != "none":
    reg = np.digitize(b["x0"][:, 1], EDGES)
    # sign / size of the position error the vehicle wants to close: front gap vs desired 50 m
    err = np.where(b["flags"] & 2, b["xl"][:, 0, 0] - b["x0"][:, 0], b["xf"][:, 0, 0] - b["x0"][:, 0] - 50.0)
    eb = np.digitize(err, [-40, -10, 0, 10, 40])
    dv = np.where(b["flags"] & 2, b["xl"][:, 1, 0], b["xf"][:, 1, 0]) - b["x0"][:, 1]
    db = np.digitize(dv, [-8, -2, 2, 8])
    key = b["flags"].astype(np.int64)
    if mode in ("role_v", "role_v_err"):
        key = key * 8 + reg
    if mode == "role_v_err":
        key = (key * 8 + eb) * 8 + db
    perm = np.argsort(key, kind="stable")
    b = {k: v[perm] for k, v in b.items()}
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
print(f"order={mode} ms={np.mean(ms):.3f} solves/s={B/np.mean(ms)*1e3:.0f} nodes={no.double().mean().item():.2f} iters={it.double().mean().item():.2f} "
      f"ok={(st==2).all().item()} objsum={ob.sum().item():.6f}")
