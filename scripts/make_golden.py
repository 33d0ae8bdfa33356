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
    make_fixed_points()
    print("wrote", sorted(os.listdir(OUT)))


def make_fixed_points():
    """Exact optima of sample QPs of every BASELINE.json configuration, from tests/qp_exact.py: the QP assembled in numpy straight
    from the reference's builders (mpc.cpp:208-306, model.cpp:30-59) and solved by a dense active-set method whose answer is
    accepted only if it passes the KKT conditions to 1e-9.  No OSQP iteration is involved, so these vectors pin the FIXED POINT
    of the solve independently of oracle/osqp_restated.hpp and tests/osqp_numpy.py."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import qp_exact as E

    def exact_batch(recs, N, gap_mode, want, state_lim=None):
        X, Y, keep = [], [], []
        for i, r in enumerate(recs):
            if len(keep) == want:
                break
            P, q, A, l, u = E.assemble(r, N, gap_mode, state_lim)
            try:
                x, y, _ = E.solve_exact(P, q, A, l, u, max_changes=150)
            except RuntimeError:
                continue            # infeasible (gap mode 1 on the all-ones stage-0 pair): not a fixed-point sample
            X.append(x); Y.append(y); keep.append(i)
        return recs[keep], np.array(X), np.array(Y)

    def n_active(Y, N):
        return int((np.abs(Y[:, 3 * (N + 1):]) > 1e-9).sum())

    samples = {
        "cfg1_skirk_N30": (W.config1_records(500)[::16], 30, 0),
        "cfg2_minipaths_N30": (W.tracking_batch(32, 30, seed=20240902), 30, 0),
        "cfg3_gap0_N30": (W.tracking_batch(32, 30, seed=20240903, gaps=True), 30, 0),
        "cfg3_gap1_N30": (W.tracking_batch(400, 30, seed=20240903, gaps=True), 30, 1),
        "cfg3_gap2_N30": (W.tracking_batch(32, 30, seed=8, gaps=True), 30, 2),
        "cfg4_lanes_N30": (W.config4_records(64)[::280], 30, 0),
        "cfg5_N10": (W.tracking_batch(24, 10, seed=20240905), 10, 0),
        "cfg5_N20": (W.tracking_batch(24, 20, seed=20240905), 20, 0),
        "cfg5_N50": (W.tracking_batch(16, 50, seed=20240905), 50, 0),
        "cfg5_N100": (W.tracking_batch(8, 100, seed=20240905), 100, 0),
    }
    # half-planes that BIND: a wall across the reference path at its 26th point (the car must brake to stay behind it) — line 2 —
    # and a far-away line 1; gap rows on from stage 1 (gap_mode 2)
    wall = W.tracking_batch(48, 30, seed=20240913)
    for r in wall:
        ref = r[11:].reshape(30, 3)
        f = (ref[29, :2] - ref[0, :2]) / np.linalg.norm(ref[29, :2] - ref[0, :2])
        r[5:8] = (f[0], f[1], 100.0)                       # f.p >= -100: never active
        r[8:11] = (-f[0], -f[1], float(f @ ref[25, :2]))    # -f.p >= -f.ref_25: stay behind the wall
    samples["cfg3_gap2_wall_N30"] = (wall, 30, 2)
    for name, (recs, N, gap_mode) in samples.items():
        recs, X, Y = exact_batch(recs, N, gap_mode, 32)
        assert len(recs) >= 8, name
        np.savez_compressed(os.path.join(OUT, "exact_%s.npz" % name), recs=recs, x=X, y=Y, N=N, gap_mode=gap_mode)
        print("exact_%s: %d QPs, %d active inequality rows, %d of them half-plane rows" %
              (name, len(recs), n_active(Y, N), int((np.abs(Y[:, 3 * (N + 1):5 * (N + 1)]) > 1e-9).sum())))
    # state-box rows (stored by the reference, never stacked: constraints.cpp:14-17, 108-114), d = params.yaml's state_lims = 1 —
    # the box binds: at the 3 m/s speed floor the car travels 0.9 m over the horizon, at 4.5 m/s 1.35 m
    for N, d in ((30, 1.0), (20, 0.7)):
        recs, X, Y = exact_batch(W.tracking_batch(32, N, seed=20240920 + N), N, 0, 32, state_lim=d)
        assert len(recs) >= 8
        np.savez_compressed(os.path.join(OUT, "exactbox_N%d_d%g.npz" % (N, d)), recs=recs, x=X, y=Y, N=N, gap_mode=0, state_lim=d)
        print("exactbox_N%d_d%g: %d QPs, %d active state-box rows" % (N, d, len(recs), int((np.abs(Y[:, 7 * N + 5:]) > 1e-9).sum())))


if __name__ == "__main__":
    main()
