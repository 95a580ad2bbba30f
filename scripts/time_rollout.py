"""Dev helper: time the rollout kernel on 1 Mi scenario-steps (n = 10) with CUDA events."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import hybrid_vehicle_platoon_b200 as hvp
from hybrid_vehicle_platoon_b200 import api

ctx = hvp.Context(0); dev = torch.device("cuda", 0)
rb, n = 1 << 20, 10
rng = np.random.default_rng(4321)
v = rng.uniform(6, 33, (rb, n)); gaps = rng.uniform(30, 150, (rb, n)); p = 3000.0 - np.cumsum(gaps, 1)
xs = np.empty((rb, 2 * n)); xs[:, 0::2] = p; xs[:, 1::2] = v
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
rx = t(xs); ru = t(rng.uniform(-1, 1, (rb, n)))
rg = t(np.clip(np.digitize(v, [9.235, 12.855, 16.93, 23.315, 32.47]) + 1, 1, 6).astype(np.int32))
rm = t(rng.uniform(700, 1000, (rb, n))); rl = t(np.stack([p[:, 0] + 5, np.full(rb, 20.0)], 1))
rxo = torch.empty_like(rx); rc = torch.empty(rb, dtype=torch.float64, device=dev)
rv = torch.empty(rb, dtype=torch.uint8, device=dev); re_ = torch.empty(rb, dtype=torch.int32, device=dev)
desc = api.env_desc(n, mass_per_scenario=True); st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ms = []
for i in range(23):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); api.rollout_step_device(desc, rb, rx, ru, rg, rm, rl, rxo, rc, rv, re_, ctx=ctx, stream=st); b.record()
    flush.fill_(1); torch.cuda.synchronize()
    if i >= 3:
        ms.append(a.elapsed_time(b))
print("rollout ms mean %.4f min %.4f  GB/s %.0f" % (np.mean(ms), np.min(ms), rb * 549 / np.mean(ms) / 1e6))
