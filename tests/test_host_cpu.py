"""CPU tests of the C++ host classes (f110-mpc_b200/host): the ROS-free mirror of the reference's Model /
Constraints / OccGrid / Transforms / Trajectory / Traj_Plan against the oracle's independent restatement.
Both are compiled with -ffp-contract=off from separately written sources; results must be bit-identical."""
import os

import numpy as np
import pytest


def test_linearize_bit_identical(pkg, oracle, workloads):
    rng = np.random.default_rng(0)
    for _ in range(50):
        th, de = rng.uniform(-3.1, 3.1), rng.uniform(-0.43, 0.43)
        A, B, C = pkg.host_linearize(th, 4.5, de, workloads.DT_F32)
        Ao, Bo, Co = oracle.linearize(th, 4.5, de, workloads.DT_F32)
        np.testing.assert_array_equal(A, Ao); np.testing.assert_array_equal(B, Bo); np.testing.assert_array_equal(C, Co)


@pytest.mark.parametrize("sd,td", [(30, 50), (19, 50), (30, 100)])
def test_traj_table_bit_identical(pkg, oracle, sd, td):
    np.testing.assert_array_equal(pkg.host_traj_table(sd, td), oracle.traj_table(steer_discrete=sd, traj_discrete=td))


def test_grid_fill_and_rotation_bit_identical(pkg, oracle, workloads):
    poses, yaws, scans = workloads.scene_batch(24, seed=3)
    for s in range(24):
        g, off = pkg.host_fill_grid(poses[s], workloads.SCAN_ANGLE_MIN, workloads.SCAN_ANGLE_MAX, workloads.SCAN_ANGLE_INC, scans[s])
        go, offo, _ = oracle.fill_grid(poses[s], workloads.SCAN_ANGLE_MIN, workloads.SCAN_ANGLE_MAX, workloads.SCAN_ANGLE_INC, scans[s])
        np.testing.assert_array_equal(g, go); np.testing.assert_array_equal(off, offo)
        assert 0 < g.sum() < g.size
        np.testing.assert_array_equal(pkg.host_car_to_world_R(poses[s]), oracle.car_to_world_R(poses[s]))


def test_half_spaces_bit_identical(pkg, oracle, workloads):
    rng = np.random.default_rng(1)
    amin, amax, inc = workloads.SCAN_ANGLE_MIN, workloads.SCAN_ANGLE_MAX, workloads.SCAN_ANGLE_INC
    n_ok = 0
    for t in range(40):
        r = rng.uniform(0.5, 2.8, 1080).astype(np.float32)
        for _ in range(rng.integers(0, 4)):
            a = rng.integers(150, 900); w = rng.integers(1, 200)
            r[a:a + w] = rng.uniform(3.5, 10.0)
        st = np.array([rng.uniform(-5, 5), rng.uniform(-5, 5), rng.uniform(-3, 3)])
        ok, l1, l2, lohi = pkg.host_find_half_spaces(st, amin, amax, inc, r)
        oko, l1o, l2o, lohio = oracle.find_half_spaces(st, amin, amax, inc, r)
        assert ok == oko and tuple(lohi) == tuple(lohio)
        if ok:
            n_ok += 1
            np.testing.assert_array_equal(l1, l1o); np.testing.assert_array_equal(l2, l2o)
    assert n_ok > 10


def test_lookahead_index_and_headings(pkg, oracle, workloads):
    xy, ori = workloads.skirk_waypoints()
    for i in range(0, 500, 23):
        pose = workloads.yaw_pose(float(xy[i, 0]), float(xy[i, 1]), float(ori[i]))
        idx, head = pkg.host_best_global_idx(xy, pose)
        assert idx == oracle.best_global_idx(xy, pose, 2.5)
    np.testing.assert_array_equal(head, oracle.waypoint_headings(xy))
    np.testing.assert_allclose(head, ori, atol=1e-6)       # numpy's float32 arctan2 in workloads.py is within an ulp of atan2f


def test_lookahead_scan_closed_form_equals_literal_scan():
    # select_kernel (csrc/pipeline_kernels.cu) replaces the reference's ordered scan with a FLOAT running minimum
    # (trajectory.cpp:103-107) by a closed form; this is the equivalence it relies on, on random and adversarial near-tie inputs
    rng = np.random.default_rng(0)

    def literal(d):
        best, idx = np.float32(3.402823466e+38), -1
        for i, x in enumerate(d):
            if x >= 0 and x < float(best):
                best, idx = np.float32(x), i
        return idx

    def closed(d):
        ahead = d >= 0
        if not ahead.any():
            return -1
        f = d.astype(np.float32)
        F = f[ahead].min()
        C = np.nonzero(ahead & (f == F))[0]
        return max([C.min()] + [i for i in C if d[i] < float(F)])

    for t in range(20000):
        n = rng.integers(1, 40)
        kind = t % 4
        if kind == 0:
            d = rng.uniform(0, 3, n)
        elif kind == 1:
            base = np.float32(rng.uniform(0.01, 3))
            d = float(base) + rng.integers(-6, 7, n) * float(np.spacing(base)) / rng.choice([1, 2, 4, 8])
        elif kind == 2:
            base = np.float32(rng.uniform(0.01, 3))
            d = float(base) + rng.uniform(-1.5, 1.5, n) * float(np.spacing(base))
        else:
            d = np.abs(rng.normal(0, 1e-3, n)); d[rng.integers(0, n)] = 0.0
        d = d.astype(np.float64)
        d[rng.random(n) < 0.2] = -1.0
        assert literal(d) == closed(d)


def test_host_classes_cpp_unit_checks(pkg, tmp_path):
    # tests/host_unit.cpp: the C++ mirror classes exercised from C++ (State, Input, Params::FromYaml, Cost, Constraints, Model,
    # Traj_Plan, OccGrid, Transforms, Trajectory) — no GPU involved
    import subprocess
    pkg.build()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    host = os.path.join(root, "f110-mpc_b200", "host")
    exe = tmp_path / "host_unit"
    subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-I", host, "-I", os.path.join(root, "include"),
                    os.path.join(root, "tests", "host_unit.cpp"), "-o", str(exe), "-L", os.path.join(root, "f110-mpc_b200"),
                    "-lf110mpc_host", "-lf110mpc_b200", "-Wl,-rpath," + os.path.join(root, "f110-mpc_b200")], check=True)
    r = subprocess.run([str(exe), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host unit checks passed" in r.stdout
