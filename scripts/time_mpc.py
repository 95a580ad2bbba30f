"""Dev helper: time the compiled-MPC kernel on a batch (device-resident inputs, CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import hybrid_vehicle_platoon_b200 as hvp
import gen_mpc_cases as G

kind = sys.argv[1] if len(sys.argv) > 1 else "cent"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
rng = np.random.default_rng(5)
fm = None
if kind == "cent":
    n, N = 3, 5
    x0, params = G.cent_cases(rng, B, n, N); mpc = hvp.api.CompiledMpc(G.CENT, N, n_local=n); nl = n
elif kind == "cent10":
    n, N = 10, 6
    x0, params = G.cent_cases(rng, B, n, N); mpc = hvp.api.CompiledMpc(G.CENT, N, n_local=n); nl = n
elif kind == "event":
    N = 6; x0, params = G.event_cases(rng, B, 2, 2, N); mpc = hvp.api.CompiledMpc(G.EVENT, N, n_local=3, leader_index=G.O_NO if hasattr(G, "O_NO") else -100, n_front=2, n_behind=2); nl = 3
elif kind == "admm":
    N = 8; x0, params = G.admm_cases(rng, B, N); mpc = hvp.api.CompiledMpc(G.ADMM, N, rho=0.5); nl = 1
elif kind == "gadmm":
    N = 8; x0, params, fm = G.gadmm_cases(rng, B, 1, 1, N); mpc = hvp.api.CompiledMpc(G.GADMM, N, n_front=1, n_behind=1, rho=0.5); nl = 1
dev = torch.device("cuda", 0)
t = lambda a, dt=torch.float64: torch.from_numpy(np.ascontiguousarray(a)).to(dev).to(dt)
dx0, dm, dp = t(x0), t(np.full((B, nl), 800.0)), t(params)
dfm = None if fm is None else t(fm, torch.int32)
N = mpc.N
u = torch.empty(B, nl, N, dtype=torch.float64, device=dev); x = torch.empty(B, nl, 2, N + 1, dtype=torch.float64, device=dev)
ex = torch.empty(B, max(mpc.n_extra, 1), dtype=torch.float64, device=dev); mo = torch.empty(B, nl, N, dtype=torch.int32, device=dev)
ob = torch.empty(B, dtype=torch.float64, device=dev); stt = torch.empty(B, dtype=torch.int32, device=dev)
no = torch.empty(B, dtype=torch.int32, device=dev); it = torch.empty(B, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
ms = []
for i in range(6):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); mpc.solve_device(B, dx0, dm, dp, dfm, u, x, ex, mo, ob, stt, no, it, stream=st); b.record()
    torch.cuda.synchronize()
    if i >= 2: ms.append(a.elapsed_time(b))
print(f"{kind}: B={B} nv={mpc.n_var} smem/warp={mpc.smem_bytes} ms={np.mean(ms):.3f} solves/s={B/np.mean(ms)*1e3:.0f} "
      f"nodes/solve={no.float().mean().item():.1f} iters/solve={it.float().mean().item():.1f} "
      f"status2={(stt==2).float().mean().item():.3f} max_nodes={no.max().item()}")
