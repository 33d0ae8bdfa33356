"""Generate the committed fixtures under tests/golden/ (run HERE, where /root/reference exists).

The reference holds no tests / golden vectors, so fixtures are of two kinds:
  * data derived from the reference's own files: the float-parsed first two columns of csv/skirk.csv
    (Trajectory::ReadCSV, trajectory.cpp:28-32) and the 10 unused mini-paths of csv/local_traj_50.csv;
  * outputs of the CPU oracle (oracle/) on seeded inputs — regression pins for oracle and CUDA path alike.
PARITY UNPINNED by the reference: these vectors pin our restatement, not the OSQP binary.
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_py as O  # noqa: E402

W = importlib.import_module("f110-mpc_b200.workloads")
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    # --- reference data, as the reference parses it (stof on the first two columns) ---
    sk = np.loadtxt(os.path.join(REF, "csv", "skirk.csv"), delimiter=",")
    skirk_xy = sk[:, :2].astype(np.float32)
    lt = np.loadtxt(os.path.join(REF, "csv", "local_traj_50.csv"), delimiter=",")
    # CSV is y-forward (rows 0,50,.. are path origins); swap to the base_link x-forward frame: (x, y) = (y_csv, -x_csv)
    local10 = np.stack([lt[:, 1], -lt[:, 0]], axis=1).reshape(10, 50, 2)
    np.savez_compressed(os.path.join(OUT, "reference_data.npz"), skirk_xy=skirk_xy, local_traj10_xy=local10,
                        skirk_heading=sk[:, 3])
    # --- oracle outputs on seeded QPs ---
    for N, B, gap_mode in ((30, 32, 0), (10, 16, 0), (30, 32, 1)):
        recs = W.tracking_batch(B, N, seed=20240900 + N + gap_mode, gaps=bool(gap_mode))
        for eps in (1e-3, 1e-4):
            mb = O.MpcBatch(O.default_cfg(N, gap_mode), O.default_settings(eps_abs=eps, eps_rel=eps, warm_start=0), B, 1)
            r = mb.solve(recs)
            np.savez_compressed(os.path.join(OUT, "qp_N%d_gap%d_eps%g.npz" % (N, gap_mode, eps)), recs=recs, x=r["x"],
                                y=r["y"], status=r["status"], iters=r["iters"], rho_updates=r["rho_updates"],
                                rho=r["rho"], N=N, gap_mode=gap_mode, eps=eps)
    # --- steering-rate rows (not in the reference): the oracle's own stacking, limits that bind ---
    for N, B, delta in ((30, 24, 0.01), (12, 16, 0.02)):
        recs = W.tracking_batch(B, N, seed=20240950 + N)
        mb = O.MpcBatch(O.default_cfg(N, 0, rate_delta=delta), O.default_settings(eps_abs=1e-4, eps_rel=1e-4, warm_start=0), B, 1)
        r = mb.solve(recs)
        np.savez_compressed(os.path.join(OUT, "qprate_N%d_delta%g.npz" % (N, delta)), recs=recs, x=r["x"], y=r["y"], status=r["status"],
                            iters=r["iters"], N=N, rate_delta=delta, eps=1e-4)
    # --- pipeline pieces ---
    tab = O.traj_table()
    A, Bm, Cv = O.linearize(0.3, 4.5, -0.1, W.DT_F32)
    np.savez_compressed(os.path.join(OUT, "pipeline.npz"), traj_table=tab, lin_A=A, lin_B=Bm, lin_C=Cv)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
