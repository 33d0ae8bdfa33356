"""GPU parity at the STATED size of every BASELINE.json configuration: every QP of the batch against the
CPU oracle — status and iteration count equal, primal / dual within 1e-4 abs + rel, first control within 1e-3 relative.

  config 1  500 warm-started skirk cycles through the C++ MPC::Update (one QP per cycle, mpc.cpp:69-143)
  config 3  1024 scan/state pairs, half-planes from Constraints::FindHalfSpaces, gap modes 0 / 1 / 2
  config 4  7 lanes x 20 mini-paths x 64 scenarios = 8960 QPs
  config 5  4096 QPs at N = 10 / 20 / 50 / 100
(config 2 — 256 scenes x 20 paths with the grid check — is test_cycle_gpu.py / test_gpu_parity.py::test_collision_check_bit_exact;
 its QP-per-surviving-path batch is below.)"""
import numpy as np
import pytest

from test_gpu_parity import assert_solution_parity

pytestmark = pytest.mark.gpu


def _both(pkg, oracle, recs, N, gap_mode=0, eps=1e-3, want_xy=True):
    B = len(recs)
    g = pkg.MpcSolver(pkg.default_config(N, gap_mode), pkg.default_settings(eps_abs=eps, eps_rel=eps, warm_start=0), max_batch=B).solve_host(recs)
    o = oracle.MpcBatch(oracle.default_cfg(N, gap_mode), oracle.default_settings(eps_abs=eps, eps_rel=eps, warm_start=0), B).solve(recs)
    return g, o


def test_config1_500_warm_started_skirk_cycles(pkg, oracle, workloads):
    N, n = 30, 500
    recs = workloads.config1_records(n, N)
    free_scan = np.full(workloads.SCAN_BEAMS, 10.0, dtype=np.float32)
    mpc = pkg.HostMPC(N)
    mpc.update_scan(workloads.SCAN_ANGLE_MIN, workloads.SCAN_ANGLE_MAX, workloads.SCAN_ANGLE_INC, free_scan)
    orc = oracle.MpcBatch(oracle.default_cfg(N), oracle.default_settings(warm_start=1), 1, 1)
    steer = 0.0
    n_warm_faster = 0
    for i in range(n):
        state, ref = recs[i, :3], recs[i, 11:].reshape(N, 3)
        g = mpc.update(state, [4.5, steer], ref)
        rec = np.concatenate([state, [4.5, steer], g["l1"], g["l2"], ref.reshape(-1)])[None, :]
        o = orc.solve(rec, warm=True)
        assert g["status"] == o["status"][0] and g["iters"] == o["iters"][0], i
        np.testing.assert_allclose(g["x"], o["x"][0], atol=1e-4, rtol=1e-4)
        np.testing.assert_allclose(g["y"], o["y"][0], atol=1e-4, rtol=1e-4)
        u0o = o["x"][0][3 * (N + 1):3 * (N + 1) + 2]
        assert (np.abs(g["inputs"][0] - u0o) <= 1e-3 * np.maximum(np.abs(u0o), [0.0, 0.05])).all()
        steer = float(g["inputs"][0, 1])          # the next cycle linearises about the applied steering (project.cpp:170)
        n_warm_faster += g["iters"] <= 25
    assert n_warm_faster > n // 2                  # the warm start does shorten most cycles


def test_config2_qp_per_surviving_path(pkg, oracle, workloads):
    S, P, N = 256, 20, 30
    poses, yaws, scans = workloads.scene_batch(S, seed=20240902)
    table = np.ascontiguousarray(workloads.traj_table(steer_discrete=P - 1)[:, :, :2])
    recs = []
    for s in range(S):
        grid, off, _ = oracle.fill_grid(poses[s], workloads.SCAN_ANGLE_MIN, workloads.SCAN_ANGLE_MAX, workloads.SCAN_ANGLE_INC, scans[s])
        v, _, _ = oracle.collision_check(grid, 100, 0.1, off, oracle.car_to_world_R(poses[s]), poses[s, :2], table)
        for p in np.nonzero(v)[0]:
            ref = np.zeros((N, 3)); ref[:, :2] = workloads.path_to_world(table[p, :N], poses[s, 0], poses[s, 1], yaws[s])
            recs.append(np.concatenate([[poses[s, 0], poses[s, 1], yaws[s]], [4.5, 0.0], [0.3, -0.8, 1.5], [-0.4, 0.7, 2.0], ref.reshape(-1)]))
    recs = np.array(recs)
    assert 1000 < len(recs) < S * P
    g, o = _both(pkg, oracle, recs, N)
    assert_solution_parity(g, o, N)
    np.testing.assert_array_equal(g["iters"], o["iters"])


@pytest.mark.parametrize("gap_mode", [0, 1, 2])
def test_config3_1024_gap_constrained(pkg, oracle, workloads, gap_mode):
    N, B = 30, 1024
    recs = workloads.tracking_batch(B, N, seed=20240903)
    scans = workloads.config3_scans(B)
    n_gap = 0
    for b in range(B):
        ok, l1, l2, _ = oracle.find_half_spaces(recs[b, :3], workloads.SCAN_ANGLE_MIN, workloads.SCAN_ANGLE_MAX, workloads.SCAN_ANGLE_INC, scans[b])
        if ok:
            recs[b, 5:8], recs[b, 8:11] = l1, l2
            n_gap += 1
    assert n_gap > B // 2
    g, o = _both(pkg, oracle, recs, N, gap_mode)
    assert_solution_parity(g, o, N)
    np.testing.assert_array_equal(g["iters"], o["iters"])
    if gap_mode:
        assert (o["status"] == 1).any() and (o["status"] != 1).any()     # half-planes through the car: some QPs are infeasible / hit max_iter


def test_config4_all_8960_qps(pkg, oracle, workloads):
    N = 30
    recs = workloads.config4_records(64, N)
    assert recs.shape[0] == 8960
    g, o = _both(pkg, oracle, recs, N)
    assert_solution_parity(g, o, N)
    np.testing.assert_array_equal(g["iters"], o["iters"])
    assert (g["status"] == 1).all()


@pytest.mark.parametrize("N", [10, 20, 50, 100])
def test_config5_4096_qps_per_horizon(pkg, oracle, workloads, N):
    recs = workloads.tracking_batch(4096, N, seed=20240905)
    g, o = _both(pkg, oracle, recs, N)
    assert_solution_parity(g, o, N)
    np.testing.assert_array_equal(g["iters"], o["iters"])
