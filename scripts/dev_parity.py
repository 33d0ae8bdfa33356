"""Developer check (GPU box): CUDA solve vs the CPU oracle on a seeded batch, with diagnostics."""
import importlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
M = importlib.import_module("f110-mpc_b200")
W = importlib.import_module("f110-mpc_b200.workloads")
from oracle import oracle_py as O


def run(N, B, eps, gap_mode=0, gaps=False):
    recs = W.tracking_batch(B, N, gaps=gaps)
    cfg = M.default_config(N, gap_mode)
    st = M.default_settings(eps_abs=eps, eps_rel=eps, warm_start=0)
    sol = M.MpcSolver(cfg, st, max_batch=B)
    t0 = time.time()
    g = sol.solve_host(recs)
    t1 = time.time()
    ob = O.MpcBatch(O.default_cfg(N, gap_mode), O.default_settings(eps_abs=eps, eps_rel=eps, warm_start=0), B)
    o = ob.solve(recs)
    ok = np.isfinite(o["x"]).all(axis=1) & np.isfinite(g["x"]).all(axis=1)
    dx = np.abs(g["x"][ok] - o["x"][ok]).max() if ok.any() else float("nan")
    dy = np.abs(g["y"][ok] - o["y"][ok]).max() if ok.any() else float("nan")
    print("N=%d B=%d eps=%g gap_mode=%d: max|dx|=%.3e max|dy|=%.3e iters_equal=%d/%d status_equal=%d/%d gpu_host_call=%.1f ms  cpu=%.1f ms (%d thr)"
          % (N, B, eps, gap_mode, dx, dy, (g["iters"] == o["iters"]).sum(), B, (g["status"] == o["status"]).sum(), B,
             (t1 - t0) * 1e3, o["seconds"] * 1e3, ob.threads))
    print("   gpu iters hist", np.bincount(g["iters"] // 25)[:8], " oracle", np.bincount(o["iters"] // 25)[:8],
          " status gpu", sorted(set(g["status"].tolist())), "oracle", sorted(set(o["status"].tolist())))
    return dx, dy


if __name__ == "__main__":
    for N in (30, 10, 20, 31, 5):
        for eps in (1e-3, 1e-4, 1e-6):
            run(N, 256, eps)
    run(30, 256, 1e-4, gap_mode=1, gaps=True)
    run(30, 4096, 1e-3)
