"""GPU tests of the device-resident planning + control cycle (f110_cycle_device):
grid fill -> collision check -> look-ahead / best path -> half-planes -> QP -> first control, against the oracle
pipeline.  The fill and gap stages call transcendental functions whose last bit may differ from glibc's, so their
parity is a mismatch COUNT (expected 0 on these inputs, bounded at 1e-4 of the cells); everything downstream of
identical inputs is exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run_cycle(pkg, workloads, S, sd, seed, use_half_spaces=1, N=30):
    dev = torch.device("cuda:0")
    poses, yaws, scans = workloads.scene_batch(S, seed=seed)
    table = np.ascontiguousarray(workloads.traj_table(steer_discrete=sd)[:, :, :2])
    xy, _ = workloads.skirk_waypoints()
    sol = pkg.MpcSolver(pkg.default_config(N), pkg.default_settings(warm_start=0), max_batch=S)
    cc = pkg.default_cycle_config(use_half_spaces=use_half_spaces)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    prev = np.random.default_rng(seed).uniform(-0.3, 0.3, S)
    P = table.shape[0]
    u0 = torch.empty(S, 2, dtype=torch.float64, device=dev); st = torch.empty(S, dtype=torch.int32, device=dev)
    it = torch.empty(S, dtype=torch.int32, device=dev); ch = torch.empty(S, dtype=torch.int32, device=dev)
    va = torch.empty(S, P, dtype=torch.uint8, device=dev)
    sol.cycle_device(cc, t(poses), t(scans), t(prev), t(table), t(xy), u0, st, it, ch, va)
    torch.cuda.synchronize()
    bufs = {k: v.cpu().numpy() for k, v in sol.cycle_buffers(S).items()}
    return dict(poses=poses, yaws=yaws, scans=scans, table=table, xy=xy, prev=prev, u0=u0.cpu().numpy(), status=st.cpu().numpy(),
                iters=it.cpu().numpy(), chosen=ch.cpu().numpy(), valid=va.cpu().numpy(), launches=sol.last_launches, **bufs)


@pytest.mark.parametrize("sd", [19, 30])
def test_cycle_matches_oracle_pipeline(pkg, oracle, workloads, sd):
    S, N = 96, 30
    amin, amax, inc = workloads.SCAN_ANGLE_MIN, workloads.SCAN_ANGLE_MAX, workloads.SCAN_ANGLE_INC
    r = _run_cycle(pkg, workloads, S, sd, seed=31 + sd)
    assert r["launches"] == 5                           # scene prep (fill + gap finder + rotation), check, select, build, solve
    cell_mismatch = 0
    recs_o = []
    n_none = 0
    for s in range(S):
        grid, off, _ = oracle.fill_grid(r["poses"][s], amin, amax, inc, r["scans"][s])
        cell_mismatch += int((grid != r["grid"][s]).sum())
        np.testing.assert_allclose(off, r["offset"][s], rtol=0, atol=5e-7)      # cosf/sinf: last float bit may differ from glibc
        # downstream stages are compared on the DEVICE grid so a (hypothetical) last-bit trig difference cannot cascade
        R = oracle.car_to_world_R(r["poses"][s])
        v, f, e = oracle.collision_check(r["grid"][s], 100, 0.1, r["offset"][s], R, r["poses"][s, :2], r["table"])
        np.testing.assert_array_equal(r["valid"][s], v)
        if v.sum() == 0:
            assert r["chosen"][s] == -1 and r["status"][s] == -10 and np.isnan(r["u0"][s]).all()
            n_none += 1
            recs_o.append(None)
            continue
        bg = oracle.best_global_idx(r["xy"], r["poses"][s], 2.5)
        assert r["best_global"][s] == bg
        pick = oracle.select_best(v, e, float(r["xy"][bg, 0]), float(r["xy"][bg, 1]))
        assert r["chosen"][s] == pick
        state = np.array([r["poses"][s, 0], r["poses"][s, 1], oracle.car_orientation(r["poses"][s])])
        ok, l1, l2, _ = oracle.find_half_spaces(state, amin, amax, inc, r["scans"][s])
        l1l2 = np.concatenate([l1, l2]) if ok else np.zeros(6)
        np.testing.assert_allclose(r["l1l2"][s], l1l2, rtol=2e-6, atol=2e-5)  # float trig: within a few float ulps of the products
        # record: reference of the chosen path in the world frame, float-narrowed (project.cpp:145-149)
        ref = np.zeros((N, 3))
        for k in range(N):
            _, _, ew = oracle.collision_check(np.zeros(10000, dtype=np.float32), 100, 0.1, np.array([1e6, 1e6], dtype=np.float32) * 0 + r["offset"][s],
                                              R, r["poses"][s, :2], r["table"][pick:pick + 1, k:k + 1, :])
            ref[k, :2] = ew[0]
        rec = np.concatenate([state, [4.5, r["prev"][s]], r["l1l2"][s], ref.reshape(-1)])
        np.testing.assert_array_equal(r["recs"][s], rec)
        recs_o.append(rec)
    assert cell_mismatch <= 1e-4 * S * 10000
    assert 0 < n_none < S
    # the QP solved on the device-built records == oracle solve of the same records
    idx = [s for s in range(S) if recs_o[s] is not None]
    o = oracle.MpcBatch(oracle.default_cfg(N), oracle.default_settings(warm_start=0), len(idx)).solve(np.array([recs_o[s] for s in idx]))
    np.testing.assert_array_equal(r["status"][idx], o["status"])
    np.testing.assert_array_equal(r["iters"][idx], o["iters"])
    np.testing.assert_allclose(r["u0"][idx], o["x"][:, 3 * (N + 1):3 * (N + 1) + 2], atol=1e-4, rtol=1e-3)


def test_cycle_without_half_spaces_and_long_horizon(pkg, oracle, workloads):
    r = _run_cycle(pkg, workloads, 40, 19, seed=5, use_half_spaces=0, N=50)
    assert r["launches"] == 5
    ok = r["chosen"] >= 0
    assert ok.any() and (r["recs"][ok][:, 5:11] == 0).all()
    o = oracle.MpcBatch(oracle.default_cfg(50), oracle.default_settings(warm_start=0), int(ok.sum())).solve(r["recs"][ok])
    np.testing.assert_array_equal(r["status"][ok], o["status"])
    np.testing.assert_allclose(r["u0"][ok], o["x"][:, 3 * 51:3 * 51 + 2], atol=1e-4, rtol=1e-3)


@pytest.mark.parametrize("qp_mode", [1, 2])
def test_cycle_qp_per_path_modes(pkg, oracle, workloads, qp_mode):
    # BASELINE config 2: a tracking QP per surviving mini-path (mode 1), or per candidate path (mode 2), built on the device
    S, sd, N = 48, 19, 30
    P = sd + 1
    poses, yaws, scans = workloads.scene_batch(S, seed=77)
    table = np.ascontiguousarray(workloads.traj_table(steer_discrete=sd)[:, :, :2])
    xy, _ = workloads.skirk_waypoints()
    sol = pkg.MpcSolver(pkg.default_config(N), pkg.default_settings(warm_start=0), max_batch=S * P)
    cc = pkg.default_cycle_config(qp_mode=qp_mode)
    prev = np.linspace(-0.2, 0.2, S)
    g = sol.cycle_host(cc, poses, scans, prev, table, xy)
    assert g["u0"].shape == (S * P, 2)
    recs = sol.cycle_buffers(S, S * P)["recs"].cpu().numpy()
    valid = g["valid"].reshape(-1).astype(bool)
    assert 0 < valid.sum() < S * P
    solve = valid if qp_mode == 1 else np.ones(S * P, dtype=bool)
    assert (g["status"][~solve] == -10).all() and np.isnan(g["u0"][~solve]).all() and np.isnan(recs[~solve][:, 3]).all()
    # every built record carries its scene's state / steering and its own path
    for slot in np.nonzero(solve)[0][::7]:
        s_, p_ = divmod(int(slot), P)
        assert recs[slot, 0] == poses[s_, 0] and recs[slot, 4] == prev[s_] and recs[slot, 3] == 4.5
        R = oracle.car_to_world_R(poses[s_])
        _, _, ew = oracle.collision_check(np.zeros(10000, dtype=np.float32), 100, 0.1, np.array([poses[s_, 0], poses[s_, 1]], dtype=np.float32),
                                          R, poses[s_, :2], table[p_:p_ + 1, N - 1:N, :])
        np.testing.assert_array_equal(recs[slot, 11 + 3 * (N - 1): 13 + 3 * (N - 1)].astype(np.float32), ew[0])
    o = oracle.MpcBatch(oracle.default_cfg(N), oracle.default_settings(warm_start=0), int(solve.sum())).solve(recs[solve])
    np.testing.assert_array_equal(g["status"][solve], o["status"])
    np.testing.assert_array_equal(g["iters"][solve], o["iters"])
    np.testing.assert_allclose(g["u0"][solve], o["x"][:, 3 * (N + 1):3 * (N + 1) + 2], atol=1e-4, rtol=1e-3)
    # the selected path of each scene is one of its valid ones
    v2 = g["valid"].astype(bool)
    for s_ in range(S):
        assert (g["chosen"][s_] == -1) == (not v2[s_].any())
        if g["chosen"][s_] >= 0:
            assert v2[s_, g["chosen"][s_]]


def test_device_gap_finder_random_scans(pkg, oracle, workloads):
    # the device gap finder visits stretches of equal predicate instead of single beams: hammer it with random scans
    # (many gaps, one-beam gaps, no gap at all, gaps touching the edge of the field of view) against the literal oracle
    S, N = 512, 30
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(123)
    amin, amax, inc = workloads.SCAN_ANGLE_MIN, workloads.SCAN_ANGLE_MAX, workloads.SCAN_ANGLE_INC
    poses = np.zeros((S, 7)); scans = np.zeros((S, 1080), dtype=np.float32)
    for s_ in range(S):
        poses[s_] = workloads.yaw_pose(rng.uniform(-5, 5), rng.uniform(-5, 5), rng.uniform(-3.1, 3.1))
        kind = s_ % 8
        r = rng.uniform(0.5, 2.8, 1080).astype(np.float32)
        if kind == 0:
            pass                                              # no gap
        elif kind == 1:
            r[rng.integers(200, 880, 30)] = 9.0               # isolated one-beam gaps
        elif kind == 2:
            r[:] = 9.0                                        # everything far
        elif kind == 3:
            r = np.where(rng.random(1080) < 0.5, 9.0, 1.0).astype(np.float32)   # salt and pepper
        else:
            for _ in range(rng.integers(1, 6)):
                a = rng.integers(0, 1075); w = rng.integers(2, 300)
                r[a:a + w] = rng.uniform(3.01, 10.0)
        scans[s_] = r
    table = np.ascontiguousarray(workloads.traj_table(steer_discrete=19)[:, :, :2])
    xy, _ = workloads.skirk_waypoints()
    sol = pkg.MpcSolver(pkg.default_config(N), pkg.default_settings(warm_start=0, max_iter=25), max_batch=S)
    sol.cycle_host(pkg.default_cycle_config(), poses, scans, None, table, xy)
    l1l2 = sol.cycle_buffers(S)["l1l2"].cpu().numpy()
    n_ok = 0
    for s_ in range(S):
        state = np.array([poses[s_, 0], poses[s_, 1], oracle.car_orientation(poses[s_])])
        ok, l1, l2, lohi = oracle.find_half_spaces(state, amin, amax, inc, scans[s_])
        want = np.concatenate([l1, l2]) if ok else np.zeros(6)
        np.testing.assert_allclose(l1l2[s_], want, rtol=2e-6, atol=2e-5, err_msg="scene %d kind %d lohi %s" % (s_, s_ % 8, lohi))
        n_ok += int(ok)
    assert 100 < n_ok < S
