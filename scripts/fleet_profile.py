"""Profiling target: a few ticks of the batched closed loop (f110_fleet_run) for 4096 cars (ncu launch-list target)."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
M = importlib.import_module("f110-mpc_b200")
W = importlib.import_module("f110-mpc_b200.workloads")
cars = int(os.environ.get("FLEET_CARS", "4096"))
table = np.ascontiguousarray(W.traj_table(steer_discrete=19)[:, :, :2])
xy, _ = W.skirk_waypoints()
sol = M.MpcSolver(M.default_config(30), M.default_settings(warm_start=1), max_batch=cars)
fleet = M.Fleet(sol, M.default_cycle_config(qp_mode=0), cars, table, xy, drive_every=2, scan_every=4, dt_tick=0.01)
p, y, s = W.scene_batch(cars, seed=20240908)
fleet.reset(np.stack([p[:, 0], p[:, 1], y], axis=1), s)
li, ld = fleet.run(int(os.environ.get("FLEET_TICKS", "12")), log=True)
print("ok", int((li[:, :, 0] == M.Fleet.CONTROL).sum()))
