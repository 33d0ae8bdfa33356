"""e2e throughput of the asynchronous cycle entry against the number of cycles in flight (f110_cycle_set_depth) and, with
F110_CYCLE_ORDERED=1, with the solves of consecutive cycles forced into submission order (the round-2a behaviour).
TUNE_GAP=1: config 3 with the gap rows on (1024 scenes, qp_mode 0, gap_mode 1) — one QP of the batch runs to max_iter."""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
M = importlib.import_module("f110-mpc_b200")
W = importlib.import_module("f110-mpc_b200.workloads")
gap = bool(os.environ.get("TUNE_GAP"))
S = 1024 if gap else 205
P = 20
poses, _, scans = W.scene_batch(S, seed=7 if gap else 11)
table = np.ascontiguousarray(W.traj_table(steer_discrete=19)[:, :, :2])
wp = np.ascontiguousarray(W.skirk_waypoints()[0], dtype=np.float32)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
h_pose, h_scan = pin(poses), pin(np.ascontiguousarray(scans, dtype=np.float32))
nq = S if gap else S * P
cc = M.default_cycle_config(qp_mode=0 if gap else 2, use_half_spaces=1)
for depth in (1, 2, 3, 4):
    sol = M.MpcSolver(M.default_config(30, 1 if gap else 0), M.default_settings(warm_start=0), max_batch=nq)
    sol.set_cycle_depth(depth)
    out = {"u0": pin(np.empty((nq, 2))), "status": pin(np.empty(nq, dtype=np.int32)), "iters": pin(np.empty(nq, dtype=np.int32)),
           "chosen": pin(np.empty(S, dtype=np.int32)), "valid": pin(np.empty((S, P), dtype=np.uint8))}
    def loop(n):
        pend = []
        for _ in range(n):
            if len(pend) == depth:
                sol.cycle_wait(pend.pop(0), out=out)
            pend.append(sol.cycle_submit(cc, h_pose, h_scan, None, table, wp))
        for t in pend:
            sol.cycle_wait(t, out=out)
    loop(5)
    torch.cuda.synchronize()
    n = 30 if gap else 200
    t0 = time.perf_counter(); loop(n); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("depth %d  %s  %.4f ms/cycle  %.3f M solves/s  (max iters %d, solved %.3f)" % (depth, "ordered" if os.environ.get("F110_CYCLE_ORDERED") else "overlap",
          dt / n * 1e3, nq * n / dt / 1e6, int(out["iters"].max()), float((out["status"] == 1).mean())))
