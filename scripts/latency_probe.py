"""B = 1 latency: host call vs kernel alone (warm-started sequence)."""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
M = importlib.import_module("f110-mpc_b200")
W = importlib.import_module("f110-mpc_b200.workloads")
N = 30
recs = W.tracking_batch(512, N, seed=4096)
dev = torch.device("cuda:0")
sol = M.MpcSolver(M.default_config(N), M.default_settings(warm_start=1), max_batch=1)
lat = []
for i in range(400):
    r = recs[i * 13 % 512: i * 13 % 512 + 1]
    t0 = time.perf_counter(); out = sol.solve_host(r, want_xy=False); lat.append((time.perf_counter() - t0) * 1e6)
lat = np.array(lat[50:]); print("solve_host B=1 (u0 only): p50 %.1f p90 %.1f us  iters %s" % (np.percentile(lat, 50), np.percentile(lat, 90), out["iters"]))
lat = []
for i in range(400):
    r = recs[i * 13 % 512: i * 13 % 512 + 1]
    t0 = time.perf_counter(); out = sol.solve_host(r, want_xy=True); lat.append((time.perf_counter() - t0) * 1e6)
lat = np.array(lat[50:]); print("solve_host B=1 (x, y too): p50 %.1f p90 %.1f us" % (np.percentile(lat, 50), np.percentile(lat, 90)))
d = torch.from_numpy(recs).to(dev)
u0 = torch.empty(1, 2, dtype=torch.float64, device=dev); st = torch.empty(1, dtype=torch.int32, device=dev); it = torch.empty(1, dtype=torch.int32, device=dev)
s = torch.cuda.current_stream().cuda_stream
ts, its = [], []
for i in range(300):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rr = d[i * 13 % 512: i * 13 % 512 + 1]
    a.record(); sol.solve_device(rr, None, None, u0, st, it, None, None, stream=s); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3); its.append(int(it[0]))
print("kernel alone B=1: p50 %.1f us, iters median %d" % (np.percentile(ts[50:], 50), np.median(its)))
t0 = time.perf_counter()
for i in range(200):
    torch.cuda.synchronize()
print("bare synchronize: %.1f us" % ((time.perf_counter() - t0) / 200 * 1e6))
# fixed-work kernels: where a lone warp's time goes
for mi in (1, 2, 26, 51):
    so = M.MpcSolver(M.default_config(N), M.default_settings(warm_start=0, max_iter=mi, check_termination=0, adaptive_rho=0), max_batch=1)
    ts = []
    for i in range(60):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); so.solve_device(d[i:i + 1], None, None, u0, st, it, None, None, stream=s); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    print("kernel alone, exactly %d iterations: p50 %.1f us" % (mi, np.percentile(ts[10:], 50)))
for sc in (0, 10):
    so = M.MpcSolver(M.default_config(N), M.default_settings(warm_start=0, max_iter=1, check_termination=0, adaptive_rho=0, scaling=sc), max_batch=1)
    ts = []
    for i in range(60):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); so.solve_device(d[i:i + 1], None, None, u0, st, it, None, None, stream=s); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    print("1 iteration, %d Ruiz passes: p50 %.1f us" % (sc, np.percentile(ts[10:], 50)))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for i in range(60):
    a.record(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
print("empty event pair: %.1f us" % np.percentile(ts, 50))
