"""Small end-to-end exercise of every kernel (for compute-sanitizer)."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
M = importlib.import_module("f110-mpc_b200")
W = importlib.import_module("f110-mpc_b200.workloads")
for N in (7, 30, 31, 50, 100):
    recs = W.tracking_batch(24, N, seed=N, gaps=True)
    for gm, ws in ((0, 0), (1, 1)):
        sol = M.MpcSolver(M.default_config(N, gm), M.default_settings(warm_start=ws), max_batch=24)
        g = sol.solve_host(recs)
        g = sol.solve_host(recs[:3])
        print(N, gm, ws, np.bincount(g["iters"] // 25))
poses, yaws, scans = W.scene_batch(12, seed=1)
table = np.ascontiguousarray(W.traj_table(steer_discrete=19)[:, :, :2])
xy, _ = W.skirk_waypoints()
for mode in (0, 1, 2):
    sol = M.MpcSolver(M.default_config(30), M.default_settings(warm_start=0), max_batch=12 * 20)
    r = sol.cycle_host(M.default_cycle_config(qp_mode=mode), poses, scans, None, table, xy)
    print("cycle", mode, (r["status"] == 1).sum())
mpc = M.HostMPC(30)
ref = np.zeros((50, 3)); ref[:, 0] = np.linspace(0, 2.2, 50)
print(mpc.update(np.zeros(3), [4.5, 0.0], ref)["status"])
print("done")
