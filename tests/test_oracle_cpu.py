"""CPU tests of the oracle (the checker itself): closed-form QPs, KKT residuals, an independent scipy
cross-check, the reference's structural facts and the committed golden vectors.
PARITY UNPINNED by the reference — these tests pin our restatement."""
import os

import numpy as np
import pytest
import scipy.linalg
import scipy.optimize

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def kkt_residuals(P, q, A, l, u, x, y):
    stat = np.abs(P @ x + q + A.T @ y).max()
    Ax = A @ x
    prim = np.abs(Ax - np.clip(Ax, l, u)).max()
    # complementarity: y>0 only at the upper bound, y<0 only at the lower bound
    comp = max(np.abs(np.minimum(y, 0) * (Ax - l))[l > -1e20].max(initial=0.0),
               np.abs(np.maximum(y, 0) * (u - Ax))[u < 1e20].max(initial=0.0))
    return stat, prim, comp


def test_closed_form_box_qp(oracle):
    # min 1/2 x'x - [1,2]'x  s.t. x1 + x2 = 1, 0 <= x <= 0.7  ->  x = (0.3, 0.7), y = (0.7, 0, 0.6)
    P = np.eye(2); q = np.array([-1.0, -2.0])
    A = np.array([[1.0, 1.0], [1.0, 0.0], [0.0, 1.0]])
    l = np.array([1.0, 0.0, 0.0]); u = np.array([1.0, 0.7, 0.7])
    r = oracle.osqp_dense(P, q, A, l, u, oracle.default_settings(eps_abs=1e-9, eps_rel=1e-9))
    assert r["status"] == oracle.SOLVED
    np.testing.assert_allclose(r["x"], [0.3, 0.7], atol=1e-7)
    np.testing.assert_allclose(r["y"], [0.7, 0.0, 0.6], atol=1e-6)


def test_closed_form_unconstrained_and_infeasible(oracle):
    P = np.diag([2.0, 4.0]); q = np.array([-2.0, -8.0])
    A = np.eye(2); l = np.full(2, -1e30); u = np.full(2, 1e30)
    r = oracle.osqp_dense(P, q, A, l, u, oracle.default_settings(eps_abs=1e-9, eps_rel=1e-9))
    np.testing.assert_allclose(r["x"], [1.0, 2.0], atol=1e-6)
    # x1 >= 1 and x1 <= 0: primal infeasible -> NaN solution like OSQP
    A2 = np.array([[1.0, 0.0], [1.0, 0.0]]); l2 = np.array([1.0, -1e30]); u2 = np.array([1e30, 0.0])
    r2 = oracle.osqp_dense(P, q, A2, l2, u2)
    assert r2["status"] == oracle.PRIMAL_INFEASIBLE and np.isnan(r2["x"]).all()
    # min -x1 with x1 unbounded above: dual infeasible
    r3 = oracle.osqp_dense(np.zeros((2, 2)), np.array([-1.0, 0.0]), np.eye(2), np.array([0.0, 0.0]), np.array([1e30, 1.0]))
    assert r3["status"] == oracle.DUAL_INFEASIBLE


def test_structure_matches_survey(oracle):
    # n = 5N+3, m = 7N+5, nnz P triu 9N+6, nnz A 26N+9, banded nnz(L) = 51N+13
    for N in (10, 20, 30, 50):
        nnzP, nnzA, nnzL = oracle.mpc_nnz(oracle.default_cfg(N))
        assert (nnzP, nnzA, nnzL) == (9 * N + 6, 26 * N + 9, 51 * N + 13)


def test_assembly_layout(oracle, workloads):
    N = 6
    cfg = oracle.default_cfg(N)
    rec = workloads.tracking_batch(1, N, seed=3)[0]
    P, q, A, l, u = oracle.mpc_assemble_dense(cfg, rec)
    n, m, ns = 5 * N + 3, 7 * N + 5, 3 * (N + 1)
    assert P.shape == (n, n) and A.shape == (m, n)
    np.testing.assert_array_equal(np.diag(P)[:ns], np.tile([10.0, 10.0, 0.0], N + 1))
    np.testing.assert_array_equal(np.diag(P)[ns:], np.tile([0.10, 5.0], N))
    Am, Bm, Cv = oracle.linearize(rec[2], rec[3], rec[4], workloads.DT_F32)
    np.testing.assert_array_equal(A[:3, :3], -np.eye(3))                     # -x0 = -x_cur (mpc.cpp:244)
    np.testing.assert_array_equal(A[3:6, 0:3], Am)                            # mpc.cpp:269
    np.testing.assert_array_equal(A[3:6, ns:ns + 2], Bm)                      # mpc.cpp:270
    np.testing.assert_array_equal(A[ns:ns + 2, 0:3], np.ones((2, 3)))         # stage-0 gap rows stay all-ones
    np.testing.assert_array_equal(A[ns + 2:ns + 4, 3:6], [[rec[5], rec[6], 0], [rec[8], rec[9], 0]])
    np.testing.assert_array_equal(l[:3], -rec[:3]); np.testing.assert_array_equal(u[3:6], -Cv)
    assert (l[ns:ns + 2 * (N + 1)] == -1e30).all() and (u[ns:ns + 2 * (N + 1)] == 1e30).all()
    np.testing.assert_array_equal(l[-2:], [3.0, float(np.float32(-0.43))])
    np.testing.assert_array_equal(u[-2:], [4.5, float(np.float32(0.43))])
    # gradient: -Q ref_k, terminal uses ref[N-1]; -R u_des (mpc.cpp:221-229)
    ref = rec[11:].reshape(N, 3)
    np.testing.assert_allclose(q[3 * N:3 * N + 2], -10.0 * ref[N - 1, :2])
    np.testing.assert_allclose(q[ns:ns + 2], [-0.45, -0.0])


def _scipy_reference(P, q, A, l, u, N):
    """Independent solve: eliminate the states through the dynamics rows, solve the 2N-variable box QP."""
    n, ns = 5 * N + 3, 3 * (N + 1)
    Ad = A[:ns]; b = l[:ns]
    Ax_, Au_ = Ad[:, :ns], Ad[:, ns:]
    Tm = -np.linalg.solve(Ax_, Au_)            # x = x_free + Tm u
    xf = np.linalg.solve(Ax_, b)
    Pxx, Puu = P[:ns, :ns], P[ns:, ns:]
    H = Tm.T @ Pxx @ Tm + Puu
    g = Tm.T @ (Pxx @ xf + q[:ns]) + q[ns:]
    L = np.linalg.cholesky(H)
    res = scipy.optimize.lsq_linear(L.T, -scipy.linalg.solve_triangular(L, g, lower=True), bounds=(l[-2 * N:], u[-2 * N:]),
                                    method="bvls", tol=1e-14, max_iter=2000)
    uo = res.x
    return np.concatenate([xf + Tm @ uo, uo])


@pytest.mark.parametrize("N", [5, 10, 30])
def test_oracle_vs_scipy_and_kkt(oracle, workloads, N):
    B = 6
    recs = workloads.tracking_batch(B, N, seed=11 + N)
    cfg = oracle.default_cfg(N)
    r = oracle.MpcBatch(cfg, oracle.default_settings(eps_abs=1e-7, eps_rel=1e-7, warm_start=0), B, 1).solve(recs)
    assert (r["status"] == oracle.SOLVED).all()
    for b in range(B):
        P, q, A, l, u = oracle.mpc_assemble_dense(cfg, recs[b])
        stat, prim, comp = kkt_residuals(P, q, A, l, u, r["x"][b], r["y"][b])
        assert stat < 1e-5 and prim < 1e-6 and comp < 1e-4
        xs = _scipy_reference(P, q, A, l, u, N)
        np.testing.assert_allclose(r["x"][b], xs, atol=2e-5, rtol=1e-5)


def test_default_tolerance_iteration_counts(oracle, workloads):
    # OSQP defaults (eps 1e-3, check every 25): the reference's operating point (mpc.cpp:98-99)
    recs = workloads.tracking_batch(64, 30)
    r = oracle.MpcBatch(oracle.default_cfg(30), oracle.default_settings(warm_start=0), 64, 2).solve(recs)
    assert (r["status"] == oracle.SOLVED).all() and (r["iters"] % 25 == 0).all() and r["iters"].max() <= 100


def test_warm_start_sequence_runs(oracle, workloads):
    recs = workloads.tracking_batch(4, 30, seed=5)
    mb = oracle.MpcBatch(oracle.default_cfg(30), oracle.default_settings(), 4, 1)
    r1 = mb.solve(recs, warm=True)
    recs2 = recs.copy(); recs2[:, 0] += 0.02; recs2[:, 4] = r1["x"][:, 94]
    r2 = mb.solve(recs2, warm=True)
    assert (r2["status"] == oracle.SOLVED).all() and r2["iters"].max() <= r1["iters"].max()
    D, E, c = mb.scaling(0)
    assert D.shape == (153,) and E.shape == (215,) and c > 0


def test_linearize_values(oracle, workloads):
    th, v, de, dt, L = 0.3, 4.5, -0.1, workloads.DT_F32, float(np.float32(0.3302))
    A, B, Cv = oracle.linearize(th, v, de, dt)
    np.testing.assert_allclose(A, [[1, 0, -v * np.sin(th) * dt], [0, 1, v * np.cos(th) * dt], [0, 0, 1]], rtol=1e-15)
    np.testing.assert_allclose(B, [[np.cos(th) * dt, 0], [np.sin(th) * dt, 0], [np.tan(de) * dt / L, v / np.cos(de) ** 2 * dt / L]], rtol=1e-14)
    np.testing.assert_allclose(Cv, [v * th * np.sin(th) * dt, -v * th * np.cos(th) * dt, -de * v / np.cos(de) ** 2 * dt / L], rtol=1e-14)


def test_traj_table_shape_and_spacing(oracle, workloads):
    t = oracle.traj_table()                     # 31 x 50 as shipped
    assert t.shape == (31, 50, 3)
    seg = np.linalg.norm(np.diff(t[:, :, :2], axis=1), axis=2)
    np.testing.assert_allclose(seg, 0.045, rtol=1e-12)
    np.testing.assert_allclose(t[15, :, 1], 0.0, atol=1e-12)   # middle path is straight
    np.testing.assert_allclose(t, workloads.traj_table(), rtol=0, atol=1e-15)


def test_world_to_occupancy_truncation(oracle):
    # (int) truncation toward zero: a point just left of the grid (index in (-1, 0)) counts as in-grid
    grid = np.zeros(100 * 100, dtype=np.float32)
    off = np.zeros(2, dtype=np.float32)
    R = np.array([1.0, 0.0, 0.0, 1.0])
    table = np.zeros((1, 2, 2)); table[0, 0] = (-5.05, 0.0); table[0, 1] = (-5.15, 0.0)
    valid, free, _ = oracle.collision_check(grid, 100, 0.1, off, R, np.zeros(2), table)
    assert free[0] == 1 and valid[0] == 0


def test_fill_grid_stamp(oracle, workloads):
    pose = workloads.yaw_pose(0.0, 0.0, 0.0)
    ranges = np.full(1080, 100.0, dtype=np.float32); ranges[540] = 2.0
    grid, off, blocks = oracle.fill_grid(pose, workloads.SCAN_ANGLE_MIN, workloads.SCAN_ANGLE_MAX, workloads.SCAN_ANGLE_INC, ranges)
    assert blocks == 100 and grid.sum() == 16.0       # one beam -> 4x4 dilation stamp
    np.testing.assert_allclose(off, [0.275, 0.0], atol=1e-7)


def test_find_half_spaces_edge_cases(oracle, workloads):
    amin, amax, inc = workloads.SCAN_ANGLE_MIN, workloads.SCAN_ANGLE_MAX, workloads.SCAN_ANGLE_INC
    r = np.full(1080, 1.0, dtype=np.float32)
    ok, l1, l2, lohi = oracle.find_half_spaces(np.zeros(3), amin, amax, inc, r)
    assert not ok and tuple(lohi) == (-1, -1)            # no gap: the reference reads ranges[-1]; oracle defines "none"
    r[500:560] = 5.0
    ok, l1, l2, lohi = oracle.find_half_spaces(np.zeros(3), amin, amax, inc, r)
    assert ok and tuple(lohi) == (503, 556)              # shrunk by buffer = 3 beams each side
    # both gap edge points lie on the >= 0 side of the other line
    a1 = amin + 503 * inc; a2 = amin + 556 * inc
    p1 = 5.0 * np.array([np.cos(a1), np.sin(a1)]); p2 = 5.0 * np.array([np.cos(a2), np.sin(a2)])
    assert l1[0] * p2[0] + l1[1] * p2[1] + (l1[2] - 0.5) >= 0 and l2[0] * p1[0] + l2[1] * p1[1] + (l2[2] - 0.5) >= 0
    # stale-hi quirk (a13'): a later, narrower gap inherits the old hi for its first beam
    r2 = np.full(1080, 1.0, dtype=np.float32); r2[400:420] = 5.0; r2[600] = 5.0
    ok, _, _, lohi2 = oracle.find_half_spaces(np.zeros(3), amin, amax, inc, r2)
    assert ok and tuple(lohi2) == (403, 416)


def test_best_global_idx_and_selection(oracle, workloads):
    xy, ori = workloads.skirk_waypoints()
    pose = workloads.yaw_pose(float(xy[0, 0]), float(xy[0, 1]), float(ori[1]))
    idx = oracle.best_global_idx(xy, pose, 2.5)
    d = np.hypot(*(xy[idx] - xy[0]))
    assert idx > 0 and abs(d - 2.5) < 0.1
    np.testing.assert_allclose(oracle.waypoint_headings(xy), ori, atol=1e-6)
    valid = np.array([0, 1, 1], dtype=np.uint8)
    endw = np.array([[0, 0], [1, 1], [1, 1]], dtype=np.float32)
    assert oracle.select_best(valid, endw, 0.0, 0.0) == 1     # strict <: first of equal distances wins


@pytest.mark.parametrize("name", sorted(f for f in os.listdir(GOLD) if f.startswith("qp_")))
def test_oracle_reproduces_golden(oracle, name):
    g = np.load(os.path.join(GOLD, name))
    N, gap_mode, eps = int(g["N"]), int(g["gap_mode"]), float(g["eps"])
    B = g["recs"].shape[0]
    r = oracle.MpcBatch(oracle.default_cfg(N, gap_mode), oracle.default_settings(eps_abs=eps, eps_rel=eps, warm_start=0), B, 1).solve(g["recs"])
    np.testing.assert_array_equal(r["status"], g["status"])
    np.testing.assert_array_equal(r["iters"], g["iters"])
    np.testing.assert_allclose(r["x"], g["x"], atol=1e-9, rtol=1e-9, equal_nan=True)
    np.testing.assert_allclose(r["y"], g["y"], atol=1e-7, rtol=1e-9, equal_nan=True)


@pytest.mark.parametrize("N,eps", [(5, 1e-3), (12, 1e-4), (30, 1e-3), (30, 1e-6)])
def test_cpp_restatement_vs_independent_numpy_osqp(oracle, workloads, N, eps):
    # iterate-level cross-check of the C++ sparse restatement against an independent dense numpy statement of OSQP:
    # same iteration counts, same rho updates, same solution
    from osqp_numpy import solve
    B = 5
    recs = workloads.tracking_batch(B, N, seed=900 + N)
    cfg = oracle.default_cfg(N)
    r = oracle.MpcBatch(cfg, oracle.default_settings(eps_abs=eps, eps_rel=eps, warm_start=0), B, 1).solve(recs)
    for b in range(B):
        P, q, A, l, u = oracle.mpc_assemble_dense(cfg, recs[b])
        o = solve(P, q, A, l, u, eps_abs=eps, eps_rel=eps)
        assert o["status"] == r["status"][b] == 1
        assert o["iters"] == r["iters"][b] and o["rho_updates"] == r["rho_updates"][b]
        np.testing.assert_allclose(r["x"][b], o["x"], atol=1e-7, rtol=1e-7)
        np.testing.assert_allclose(r["y"][b], o["y"], atol=1e-5, rtol=1e-6)
        assert abs(r["rho"][b] - o["rho"]) <= 1e-6 * o["rho"]


# ---- steering-rate rows (not in the reference) ------------------------------------------------
def test_rate_rows_structure(oracle, workloads):
    N, D = 12, 0.02
    rec = workloads.tracking_batch(1, N, seed=1)[0]
    cfg0, cfg1 = oracle.default_cfg(N), oracle.default_cfg(N, 0, rate_delta=D)
    P0, q0, A0, l0, u0 = oracle.mpc_assemble_dense(cfg0, rec)
    P1, q1, A1, l1, u1 = oracle.mpc_assemble_dense(cfg1, rec)
    m0 = 7 * N + 5
    assert A1.shape == (8 * N + 5, 5 * N + 3) and oracle.mpc_rows(cfg1) == 8 * N + 5
    np.testing.assert_array_equal(A1[:m0], A0); np.testing.assert_array_equal(P1, P0); np.testing.assert_array_equal(q1, q0)
    np.testing.assert_array_equal(l1[:m0], l0); np.testing.assert_array_equal(u1[:m0], u0)
    S = np.zeros((N, 5 * N + 3))                      # row k: +delta_k - delta_{k-1}
    for k in range(N):
        S[k, 3 * (N + 1) + 2 * k + 1] = 1.0
        if k:
            S[k, 3 * (N + 1) + 2 * (k - 1) + 1] = -1.0
    np.testing.assert_array_equal(A1[m0:], S)
    np.testing.assert_array_equal(l1[m0 + 1:], -D); np.testing.assert_array_equal(u1[m0 + 1:], D)
    assert l1[m0] == rec[4] - D and u1[m0] == rec[4] + D     # row 0 is measured from the steering applied last cycle
    nnzP, nnzA, nnzL = oracle.mpc_nnz(cfg1)
    assert nnzA == 26 * N + 9 + 2 * N - 1


@pytest.mark.parametrize("N,eps", [(5, 1e-4), (30, 1e-3), (30, 1e-5)])
def test_rate_rows_oracle_vs_numpy_osqp_and_kkt(oracle, workloads, N, eps):
    from osqp_numpy import solve
    B, D = 4, 0.01
    recs = workloads.tracking_batch(B, N, seed=950 + N)
    cfg = oracle.default_cfg(N, 0, rate_delta=D)
    r = oracle.MpcBatch(cfg, oracle.default_settings(eps_abs=eps, eps_rel=eps, warm_start=0), B, 1).solve(recs)
    for b in range(B):
        P, q, A, l, u = oracle.mpc_assemble_dense(cfg, recs[b])
        o = solve(P, q, A, l, u, eps_abs=eps, eps_rel=eps)
        assert o["status"] == r["status"][b] == 1
        assert o["iters"] == r["iters"][b] and o["rho_updates"] == r["rho_updates"][b]
        np.testing.assert_allclose(r["x"][b], o["x"], atol=1e-7, rtol=1e-7)
        np.testing.assert_allclose(r["y"][b], o["y"], atol=1e-5, rtol=1e-6)
    if eps <= 1e-5:
        for b in range(B):
            P, q, A, l, u = oracle.mpc_assemble_dense(cfg, recs[b])
            stat, prim, comp = kkt_residuals(P, q, A, l, u, r["x"][b], r["y"][b])
            assert stat < 1e-3 and prim < 1e-4
            steer = r["x"][b][3 * (N + 1) + 1::2]
            assert np.abs(np.diff(np.concatenate([[recs[b, 4]], steer]))).max() <= D + 1e-4


@pytest.mark.parametrize("name", ["qprate_N30_delta0.01.npz", "qprate_N12_delta0.02.npz"])
def test_oracle_reproduces_rate_golden_vectors(oracle, name):
    gd = np.load(os.path.join(os.path.dirname(__file__), "golden", name))
    N, delta, eps = int(gd["N"]), float(gd["rate_delta"]), float(gd["eps"])
    B = gd["recs"].shape[0]
    r = oracle.MpcBatch(oracle.default_cfg(N, 0, rate_delta=delta), oracle.default_settings(eps_abs=eps, eps_rel=eps, warm_start=0), B, 1).solve(gd["recs"])
    np.testing.assert_array_equal(r["status"], gd["status"]); np.testing.assert_array_equal(r["iters"], gd["iters"])
    np.testing.assert_allclose(r["x"], gd["x"], atol=1e-9, rtol=1e-9, equal_nan=True)
