"""Kernel-only timing of the ADMM solve for the library named by $F110_LIB (tuning builds)."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
M = importlib.import_module("f110-mpc_b200")
W = importlib.import_module("f110-mpc_b200.workloads")
N = int(os.environ.get("TUNE_N", "30")); B = int(os.environ.get("TUNE_B", "4096"))
recs = W.tracking_batch(B, N, seed=4096)
dev = torch.device("cuda:0")
kw = dict(warm_start=0)
if os.environ.get("TUNE_FIXED"):      # fixed work per QP: exactly 50 iterations, one final check (for what-if experiments)
    kw.update(max_iter=int(os.environ.get("TUNE_ITERS", "50")), check_termination=0, adaptive_rho=0)
RATE = float(os.environ["TUNE_RATE"]) if os.environ.get("TUNE_RATE") else None   # steering-rate rows: max step (rad)
sol = M.MpcSolver(M.default_config(N, 0, rate_delta=RATE), M.default_settings(**kw), max_batch=B)
r = torch.from_numpy(recs).to(dev)
u0 = torch.empty(B, 2, dtype=torch.float64, device=dev); st = torch.empty(B, dtype=torch.int32, device=dev); it = torch.empty(B, dtype=torch.int32, device=dev)
s = torch.cuda.current_stream().cuda_stream
for _ in range(3): sol.solve_device(r, None, None, u0, st, it, None, None, stream=s)
torch.cuda.synchronize()
ts = []
for _ in range(10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); sol.solve_device(r, None, None, u0, st, it, None, None, stream=s); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print("%-40s N=%d B=%d  min %.4f ms  med %.4f ms  iters_sum %d  u0sum %.9f" % (os.path.basename(M.LIB_PATH), N, B, min(ts), float(np.median(ts)), int(it.sum()), float(u0.sum())))
