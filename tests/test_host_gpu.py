"""GPU tests of the C++ host classes that reach the device through the C ABI: MPC::Update (B = 1, warm started
like the reference's per-cycle OSQP object) and the planning cycle of project.cpp:73-157."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_mpc_update_sequence_matches_oracle(pkg, oracle, workloads):
    N = 30
    amin, amax, inc = workloads.SCAN_ANGLE_MIN, workloads.SCAN_ANGLE_MAX, workloads.SCAN_ANGLE_INC
    scan = np.full(1080, 1.5, dtype=np.float32); scan[450:640] = 8.0          # one wide gap ahead
    mpc = pkg.HostMPC(N)
    mpc.update_scan(amin, amax, inc, scan)                                   # frozen first scan (project.cpp:45-49)
    mb = oracle.MpcBatch(oracle.default_cfg(N), oracle.default_settings(warm_start=1), 1, 1)
    table = workloads.traj_table()
    x, y, yaw, steer = 1.0, -2.0, 0.4, 0.0
    for cycle in range(5):
        ref = np.zeros((50, 3)); ref[:, :2] = workloads.path_to_world(table[18, :, :2], x, y, yaw)
        state = np.array([x + 0.05, y - 0.03, yaw + 0.02])
        g = mpc.update(state, [4.5, steer], ref)
        ok, l1, l2, _ = oracle.find_half_spaces(state, amin, amax, inc, scan)
        assert ok
        np.testing.assert_array_equal(g["l1"], l1); np.testing.assert_array_equal(g["l2"], l2)
        rec = np.concatenate([state, [4.5, steer], l1, l2, ref[:N].reshape(-1)])[None, :]
        o = mb.solve(rec, warm=True)
        assert g["status"] == o["status"][0] == 1 and g["iters"] == o["iters"][0]
        np.testing.assert_allclose(g["x"], o["x"][0], atol=1e-4, rtol=1e-4)
        np.testing.assert_allclose(g["y"], o["y"][0], atol=1e-4, rtol=1e-4)
        u = o["x"][0][3 * (N + 1):].reshape(N, 2)
        assert g["inputs"].shape == (N, 2)                                   # mpc.cpp:145-159
        np.testing.assert_allclose(g["inputs"], u, atol=1e-4, rtol=1e-3)
        steer = float(u[0, 1])
        x += 0.02 * u[0, 0] * np.cos(yaw); y += 0.02 * u[0, 0] * np.sin(yaw)


def test_mpc_keeps_previous_trajectory_on_failure(pkg, oracle, workloads):
    N = 30
    mpc = pkg.HostMPC(N, gap_mode=1)
    table = workloads.traj_table()
    ref = np.zeros((50, 3)); ref[:, :2] = table[15, :, :2]
    good = mpc.update(np.zeros(3), [4.5, 0.0], ref)                          # no scan yet: l1 = l2 = 0, rows vacuous
    assert good["status"] == 1 and len(good["inputs"]) == N
    amin, amax, inc = workloads.SCAN_ANGLE_MIN, workloads.SCAN_ANGLE_MAX, workloads.SCAN_ANGLE_INC
    scan = np.full(1080, 1.0, dtype=np.float32); scan[300:330] = 9.0
    mpc.update_scan(amin, amax, inc, scan)
    # find a state for which the gap-enabled QP is infeasible (the stage-0 all-ones row)
    bad_state = None
    for cand in ([-40.0, -40.0, -3.0], [-30.0, -50.0, 1.0], [-60.0, -10.0, 2.5], [-25.0, -25.0, 0.0]):
        ok, l1, l2, _ = oracle.find_half_spaces(np.array(cand), amin, amax, inc, scan)
        rec = np.concatenate([cand, [4.5, 0.0], l1, l2, ref[:N].reshape(-1)])[None, :]
        o = oracle.MpcBatch(oracle.default_cfg(N, 1), oracle.default_settings(), 1, 1).solve(rec)
        if o["status"][0] != 1:
            bad_state = np.array(cand)
            break
    assert bad_state is not None
    bad = mpc.update(bad_state, [4.5, 0.0], ref)
    assert bad["status"] == o["status"][0] != 1
    np.testing.assert_array_equal(bad["inputs"], good["inputs"])             # mpc.cpp:133-142: previous trajectory kept


@pytest.mark.parametrize("sd", [19, 30])
def test_planning_cycle_matches_oracle(pkg, oracle, workloads, sd):
    amin, amax, inc = workloads.SCAN_ANGLE_MIN, workloads.SCAN_ANGLE_MAX, workloads.SCAN_ANGLE_INC
    xy, _ = workloads.skirk_waypoints()
    poses, yaws, scans = workloads.scene_batch(40, seed=11)
    table = np.ascontiguousarray(oracle.traj_table(steer_discrete=sd)[:, :, :2])
    n_planned = 0
    for s in range(40):
        idx, path, valid, bg = pkg.host_plan(poses[s], amin, amax, inc, scans[s], xy, steer_discrete=sd)
        grid, off, _ = oracle.fill_grid(poses[s], amin, amax, inc, scans[s])
        R = oracle.car_to_world_R(poses[s])
        v, f, e = oracle.collision_check(grid, 100, 0.1, off, R, poses[s, :2], table)
        np.testing.assert_array_equal(valid, v)
        if v.sum() == 0:
            assert idx == -1
            continue
        bgo = oracle.best_global_idx(xy, poses[s], 2.5)
        assert bg == bgo
        idxo = oracle.select_best(v, e, float(xy[bgo, 0]), float(xy[bgo, 1]))
        assert idx == idxo
        # chosen mini-path in the world frame, float-narrowed, ori = 0.0 (project.cpp:145-149)
        _, _, endw = oracle.collision_check(np.zeros(10000, dtype=np.float32), 100, 0.1, off, R, poses[s, :2], table[idx:idx + 1, -1:, :])
        np.testing.assert_array_equal(path[-1, :2].astype(np.float32), endw[0])
        assert (path[:, 2] == 0).all()
        n_planned += 1
    assert n_planned > 5


def test_closed_loop_follows_raceline(pkg, workloads):
    # the project orchestrator (plan -> track -> re-plan near the path end) against
    # Model::simulate_dynamics, free scan, start on the raceline.  The car must make progress along skirk and stay near it.
    xy, ori = workloads.skirk_waypoints()
    scan = np.full(1080, 10.0, dtype=np.float32)
    ticks = 600
    traj, solved, plans = pkg.host_closed_loop(ticks, xy, [float(xy[5, 0]), float(xy[5, 1]), float(ori[5])], workloads.SCAN_ANGLE_MIN,
                                               workloads.SCAN_ANGLE_MAX, workloads.SCAN_ANGLE_INC, scan)
    assert plans >= 3 and solved > 400                       # re-planned several times, an MPC solve on most ticks
    d = np.sqrt(((traj[:, None, :2] - xy[None, :, :].astype(np.float64)) ** 2).sum(-1))
    nearest = d.argmin(axis=1)
    assert d.min(axis=1)[50:].max() < 0.8                    # stays within 0.8 m of the raceline
    progress = (nearest[-1] - nearest[0]) % len(xy)
    assert 150 < progress < 450                              # ~6 s at ~4.4 m/s on a 31.9 m lap of 500 points
    assert (traj[100:, 3] >= 3.0 - 1e-3).all() and (traj[100:, 3] <= 4.5 + 1e-3).all() and (np.abs(traj[:, 4]) <= 0.43 + 1e-3).all()


def test_batch_mpc_over_all_visible_gpus_cpp(pkg, tmp_path):
    # tests/host_multi_gpu.cpp: the C++ BatchMPC over every visible device (f110_mpc_create_multi) against the one-device BatchMPC
    import os
    import subprocess
    pkg.build()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    host = os.path.join(root, "f110-mpc_b200", "host")
    exe = tmp_path / "host_multi_gpu"
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", host, "-I", os.path.join(root, "include"), os.path.join(root, "tests", "host_multi_gpu.cpp"),
                    "-o", str(exe), "-L", os.path.join(root, "f110-mpc_b200"), "-lf110mpc_host", "-lf110mpc_b200",
                    "-Wl,-rpath," + os.path.join(root, "f110-mpc_b200")], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "multi-GPU host checks passed" in r.stdout
