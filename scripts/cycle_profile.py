"""Profiling target: a few device cycles (f110_cycle_device) for 4096 scenes."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
M = importlib.import_module("f110-mpc_b200")
W = importlib.import_module("f110-mpc_b200.workloads")
S = 4096
dev = torch.device("cuda:0")
poses, yaws, scans = W.scene_batch(S, seed=20240906)
table = np.ascontiguousarray(W.traj_table(steer_discrete=19)[:, :, :2])
xy, _ = W.skirk_waypoints()
sol = M.MpcSolver(M.default_config(30), M.default_settings(warm_start=0), max_batch=S)
cc = M.default_cycle_config()
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
d_pose, d_scan, d_tab, d_wp = t(poses), t(scans), t(table), t(xy)
u0 = torch.empty(S, 2, dtype=torch.float64, device=dev); st = torch.empty(S, dtype=torch.int32, device=dev)
it = torch.empty(S, dtype=torch.int32, device=dev); ch = torch.empty(S, dtype=torch.int32, device=dev)
for _ in range(3):
    sol.cycle_device(cc, d_pose, d_scan, None, d_tab, d_wp, u0, st, it, ch)
torch.cuda.synchronize()
print("ok", int((st == 1).sum()))
