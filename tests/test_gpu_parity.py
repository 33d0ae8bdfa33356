"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs and against the committed golden vectors.

Tolerances (BASELINE.json north_star): primal and dual within eps_abs = eps_rel = 1e-4 of the oracle,
first applied control within 1e-3 relative (absolute floor 0.05 on the steering, whose set-point is 0);
collision-check outputs bit-exact.  PARITY UNPINNED by the reference (it has no tests)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL_ABS = TOL_REL = 1e-4


def assert_solution_parity(g, o, N):
    np.testing.assert_array_equal(g["status"], o["status"])
    ok = ~np.isin(o["status"], [3, 4, -3, -4, -7])   # OSQP returns its last iterate unless the problem is infeasible / non-convex
    np.testing.assert_allclose(g["x"][ok], o["x"][ok], atol=TOL_ABS, rtol=TOL_REL)
    np.testing.assert_allclose(g["y"][ok], o["y"][ok], atol=TOL_ABS, rtol=TOL_REL)
    assert np.isnan(g["x"][~ok]).all() and np.isnan(g["y"][~ok]).all()     # OSQP NaN-fills infeasible solutions
    u0g, u0o = g["x"][ok][:, 3 * (N + 1):3 * (N + 1) + 2], o["x"][ok][:, 3 * (N + 1):3 * (N + 1) + 2]
    floor = np.array([0.0, 0.05])
    assert (np.abs(u0g - u0o) <= 1e-3 * np.maximum(np.abs(u0o), floor)).all()
    np.testing.assert_array_equal(g["u0"][ok], u0g)


@pytest.mark.parametrize("N", [1, 5, 10, 20, 30, 31])
@pytest.mark.parametrize("eps", [1e-3, 1e-4])
def test_cold_batch_matches_oracle(pkg, oracle, workloads, N, eps):
    B = 192
    recs = workloads.tracking_batch(B, N, seed=100 + N)
    g = pkg.MpcSolver(pkg.default_config(N), pkg.default_settings(eps_abs=eps, eps_rel=eps, warm_start=0), B).solve_host(recs)
    o = oracle.MpcBatch(oracle.default_cfg(N), oracle.default_settings(eps_abs=eps, eps_rel=eps, warm_start=0), B).solve(recs)
    assert_solution_parity(g, o, N)
    np.testing.assert_array_equal(g["iters"], o["iters"])        # same algorithm -> same iteration counts


@pytest.mark.parametrize("N", [32, 50, 63, 64, 100, 127])
def test_long_horizons_match_oracle(pkg, oracle, workloads, N):
    # horizons above 31: the QP spans 2 or 4 warps of one CTA (config 5 of BASELINE.json: N = 50, 100)
    B, eps = 96, 1e-4
    recs = workloads.tracking_batch(B, N, seed=200 + N)
    g = pkg.MpcSolver(pkg.default_config(N), pkg.default_settings(eps_abs=eps, eps_rel=eps, warm_start=0), B).solve_host(recs)
    o = oracle.MpcBatch(oracle.default_cfg(N), oracle.default_settings(eps_abs=eps, eps_rel=eps, warm_start=0), B).solve(recs)
    assert_solution_parity(g, o, N)
    np.testing.assert_array_equal(g["iters"], o["iters"])


def test_long_horizon_gap_mode_and_warm_start(pkg, oracle, workloads):
    N, B, eps = 50, 64, 1e-4
    recs = workloads.tracking_batch(B, N, seed=51, gaps=True)
    sol = pkg.MpcSolver(pkg.default_config(N, 1), pkg.default_settings(eps_abs=eps, eps_rel=eps, warm_start=1), B)
    mb = oracle.MpcBatch(oracle.default_cfg(N, 1), oracle.default_settings(eps_abs=eps, eps_rel=eps, warm_start=1), B)
    for step in range(3):
        g = sol.solve_host(recs)
        o = mb.solve(recs, warm=True)
        assert_solution_parity(g, o, N)
        np.testing.assert_array_equal(g["iters"], o["iters"])
        recs = recs.copy(); recs[:, 0] += 0.01; recs[:, 4] *= 0.9


def test_gap_enabled_mode_including_infeasible(pkg, oracle, workloads):
    N, B, eps = 30, 256, 1e-4
    recs = workloads.tracking_batch(B, N, gaps=True)
    g = pkg.MpcSolver(pkg.default_config(N, 1), pkg.default_settings(eps_abs=eps, eps_rel=eps, warm_start=0), B).solve_host(recs)
    o = oracle.MpcBatch(oracle.default_cfg(N, 1), oracle.default_settings(eps_abs=eps, eps_rel=eps, warm_start=0), B).solve(recs)
    assert (o["status"] == oracle.PRIMAL_INFEASIBLE).any() and (o["status"] == oracle.SOLVED).any()
    assert_solution_parity(g, o, N)
    np.testing.assert_array_equal(g["iters"], o["iters"])


def test_gap_mode_2_half_planes_active(pkg, oracle, workloads):
    # gap rows on for stages k >= 1 only: feasible laser-gap constrained MPC ("approach 2" of the reference README)
    N, B, eps = 30, 256, 1e-4
    recs = workloads.tracking_batch(B, N, seed=8, gaps=True)
    g = pkg.MpcSolver(pkg.default_config(N, 2), pkg.default_settings(eps_abs=eps, eps_rel=eps, warm_start=0), B).solve_host(recs)
    o = oracle.MpcBatch(oracle.default_cfg(N, 2), oracle.default_settings(eps_abs=eps, eps_rel=eps, warm_start=0), B).solve(recs)
    assert (o["status"] == oracle.SOLVED).mean() > 0.9
    assert_solution_parity(g, o, N)
    np.testing.assert_array_equal(g["iters"], o["iters"])
    # the half-plane rows hold on the solution and some are active (non-zero multipliers)
    ok = o["status"] == 1
    xs = g["x"][ok][:, 3:3 * (N + 1)].reshape(-1, N, 3)
    for r in range(2):
        a, b, c = recs[ok][:, 5 + 3 * r], recs[ok][:, 6 + 3 * r], recs[ok][:, 7 + 3 * r]
        assert (a[:, None] * xs[:, :, 0] + b[:, None] * xs[:, :, 1] >= -c[:, None] - 1e-3).all()
    yg = g["y"][ok][:, 3 * (N + 1) + 2:5 * (N + 1)]
    assert (np.abs(yg) > 1e-6).any()


def test_tight_tolerance_solutions_agree(pkg, oracle, workloads):
    # at eps 1e-6 both sides sit on the (unique) solution; iteration counts may differ by one check period
    N, B, eps = 30, 128, 1e-6
    recs = workloads.tracking_batch(B, N, seed=77)
    g = pkg.MpcSolver(pkg.default_config(N), pkg.default_settings(eps_abs=eps, eps_rel=eps, warm_start=0), B).solve_host(recs)
    o = oracle.MpcBatch(oracle.default_cfg(N), oracle.default_settings(eps_abs=eps, eps_rel=eps, warm_start=0), B).solve(recs)
    np.testing.assert_allclose(g["x"], o["x"], atol=1e-6, rtol=1e-6)
    np.testing.assert_allclose(g["y"], o["y"], atol=1e-5, rtol=1e-6)
    assert np.abs(g["iters"] - o["iters"]).max() <= 25


@pytest.mark.parametrize("name", sorted(f for f in os.listdir(GOLD) if f.startswith("qp_")))
def test_golden_vectors(pkg, name):
    gd = np.load(os.path.join(GOLD, name))
    N, gap_mode, eps = int(gd["N"]), int(gd["gap_mode"]), float(gd["eps"])
    B = gd["recs"].shape[0]
    g = pkg.MpcSolver(pkg.default_config(N, gap_mode), pkg.default_settings(eps_abs=eps, eps_rel=eps, warm_start=0), B).solve_host(gd["recs"])
    assert_solution_parity(g, dict(status=gd["status"], x=gd["x"], y=gd["y"]), N)
    np.testing.assert_array_equal(g["iters"], gd["iters"])


@pytest.mark.parametrize("N,B", [(30, 64), (10, 37), (5, 23), (50, 16)])   # 10, 5: several QPs share a warp (ragged last warp)
def test_warm_start_sequence_matches_oracle(pkg, oracle, workloads, N, B):
    # reference steady state: update q / A / bounds, keep iterates and rho (mpc.cpp:83-94, 98)
    recs = workloads.tracking_batch(B, N, seed=9)
    sol = pkg.MpcSolver(pkg.default_config(N), pkg.default_settings(eps_abs=1e-4, eps_rel=1e-4, warm_start=1), B)
    mb = oracle.MpcBatch(oracle.default_cfg(N), oracle.default_settings(eps_abs=1e-4, eps_rel=1e-4, warm_start=1), B)
    for step in range(4):
        g = sol.solve_host(recs)
        o = mb.solve(recs, warm=True)
        assert_solution_parity(g, o, N)
        np.testing.assert_array_equal(g["iters"], o["iters"])
        u0 = o["x"][:, 3 * (N + 1):3 * (N + 1) + 2]
        recs = recs.copy()
        recs[:, 0] += 0.01 * u0[:, 0] * np.cos(recs[:, 2])       # roll the car forward one step
        recs[:, 1] += 0.01 * u0[:, 0] * np.sin(recs[:, 2])
        recs[:, 4] = u0[:, 1]
    sol.reset()
    g = sol.solve_host(recs)
    o = oracle.MpcBatch(oracle.default_cfg(N), oracle.default_settings(eps_abs=1e-4, eps_rel=1e-4, warm_start=0), B).solve(recs)
    assert_solution_parity(g, o, N)


def test_full_size_batch_kkt_properties(pkg, oracle, workloads):
    # BASELINE size (4096 QPs): size-independent properties — KKT residuals of every returned (x, y)
    N, B = 30, 4096
    recs = workloads.tracking_batch(B, N, seed=4096)
    g = pkg.MpcSolver(pkg.default_config(N), pkg.default_settings(eps_abs=1e-5, eps_rel=1e-5, warm_start=0), B).solve_host(recs)
    assert (g["status"] == pkg.SOLVED).all()
    cfg = oracle.default_cfg(N)
    P, q, A, l, u = oracle.mpc_assemble_dense(cfg, recs[0])
    eps = 1e-5
    for b in range(0, B, 37):
        P, q, A, l, u = oracle.mpc_assemble_dense(cfg, recs[b])
        x, y = g["x"][b], g["y"][b]
        Ax = A @ x
        # OSQP's own stopping rule, evaluated independently (z = projection of Ax gives the smallest residual)
        dua = np.abs(P @ x + q + A.T @ y).max()
        pri = np.abs(Ax - np.clip(Ax, l, u)).max()
        assert dua <= eps + eps * max(np.abs(P @ x).max(), np.abs(A.T @ y).max(), np.abs(q).max())
        assert pri <= eps + eps * np.abs(Ax).max()
    # all inputs inside the box, x0 pinned
    u_all = g["x"][:, 3 * (N + 1):].reshape(B, N, 2)
    assert (u_all[..., 0] >= 3 - 1e-4).all() and (u_all[..., 0] <= 4.5 + 1e-4).all() and (np.abs(u_all[..., 1]) <= 0.43 + 1e-4).all()
    np.testing.assert_allclose(g["x"][:, :3], recs[:, :3], atol=1e-4)


def test_device_entry_both_record_paths(pkg, workloads):
    # f110_mpc_solve_device stages records by TMA bulk copy when they are 16-byte aligned (even stride), with plain loads otherwise
    import torch
    N, B = 30, 128
    recs = workloads.tracking_batch(B, N, seed=21)
    dev = torch.device("cuda:0")
    outs = []
    for pad in (0, 1, 3):                       # strides 101 (plain), 102 (TMA), 104 (TMA)
        r = torch.from_numpy(np.pad(recs, ((0, 0), (0, pad)), constant_values=np.nan)).to(dev)
        sol = pkg.MpcSolver(pkg.default_config(N), pkg.default_settings(warm_start=0), B)
        x = torch.empty(B, sol.n, dtype=torch.float64, device=dev); it = torch.empty(B, dtype=torch.int32, device=dev)
        st = torch.empty(B, dtype=torch.int32, device=dev)
        sol.solve_device(r, x, None, None, st, it, None, None)
        torch.cuda.synchronize()
        outs.append((x.cpu().numpy(), it.cpu().numpy(), st.cpu().numpy()))
    for o in outs[1:]:
        np.testing.assert_array_equal(o[0], outs[0][0]); np.testing.assert_array_equal(o[1], outs[0][1])
    assert (outs[0][2] == 1).all()


def test_api_errors(pkg):
    sol = pkg.MpcSolver(max_batch=8)
    with pytest.raises(RuntimeError, match="exceeds max_batch"):
        sol.solve_host(np.zeros((9, 101)))
    with pytest.raises(RuntimeError):
        pkg.MpcSolver(pkg.default_config(200), max_batch=8)
    assert sol.solve_host(np.zeros((0, 101)))["status"].shape == (0,)


# ---- collision check: bit-exact ------------------------------------------------------------------------------
def _scene_inputs(oracle, workloads, S, seed):
    poses, yaws, scans = workloads.scene_batch(S, seed)
    grids = np.zeros((S, 100 * 100), dtype=np.float32)
    offs = np.zeros((S, 2), dtype=np.float32)
    rots = np.zeros((S, 4))
    for s in range(S):
        grids[s], offs[s], _ = oracle.fill_grid(poses[s], workloads.SCAN_ANGLE_MIN, workloads.SCAN_ANGLE_MAX, workloads.SCAN_ANGLE_INC, scans[s])
        rots[s] = oracle.car_to_world_R(poses[s])
    return poses, grids, offs, rots


@pytest.mark.parametrize("table_kind", ["steer19", "steer30", "csv10"])
def test_collision_check_bit_exact(pkg, oracle, workloads, table_kind):
    S = 96
    poses, grids, offs, rots = _scene_inputs(oracle, workloads, S, 20240902)
    if table_kind == "csv10":
        table = workloads.reference_data()["local_traj10_xy"]
    else:
        table = oracle.traj_table(steer_discrete=19 if table_kind == "steer19" else 30)[:, :, :2]
    table = np.ascontiguousarray(table)
    valid, free, endw = pkg.collision_check_host(grids, offs, rots, poses[:, :2], table)
    n_valid = 0
    for s in range(S):
        v, f, e = oracle.collision_check(grids[s], 100, 0.1, offs[s], rots[s], poses[s, :2], table)
        np.testing.assert_array_equal(valid[s], v)
        np.testing.assert_array_equal(free[s], f)
        np.testing.assert_array_equal(endw[s].view(np.uint32), e.view(np.uint32))     # bit pattern
        n_valid += int(v.sum())
    assert 0 < n_valid < S * table.shape[0]          # the scenes exercise both outcomes


def test_collision_check_edges(pkg, oracle):
    # ragged sample counts (not a multiple of 32), out-of-grid samples, fully occupied and empty grids
    rng = np.random.default_rng(5)
    S = 8
    grids = (rng.random((S, 100 * 100)) < 0.02).astype(np.float32)
    grids[0] = 0.0
    grids[1] = 1.0
    offs = rng.uniform(-1, 1, (S, 2)).astype(np.float32)
    yaw = rng.uniform(-3, 3, S)
    rots = np.stack([np.cos(yaw), -np.sin(yaw), np.sin(yaw), np.cos(yaw)], axis=1)
    pose = rng.uniform(-1, 1, (S, 2))
    for samples in (1, 31, 33, 50, 100):
        table = rng.uniform(-6, 6, (7, samples, 2))
        table[0] *= 0.1                                   # one short path that stays inside the grid
        valid, free, endw = pkg.collision_check_host(grids, offs, rots, pose, table)
        for s in range(S):
            v, f, e = oracle.collision_check(grids[s], 100, 0.1, offs[s], rots[s], pose[s], table)
            np.testing.assert_array_equal(valid[s], v)
            np.testing.assert_array_equal(free[s], f)
            np.testing.assert_array_equal(endw[s].view(np.uint32), e.view(np.uint32))
        assert valid[0, 0] == 1 and valid[1].sum() == 0


SETTINGS_VARIANTS = [
    dict(scaling=0), dict(scaling=3), dict(adaptive_rho=0), dict(adaptive_rho_interval=50), dict(adaptive_rho_tolerance=2.0),
    dict(check_termination=10), dict(check_termination=0, max_iter=120), dict(max_iter=30), dict(max_iter=1),
    dict(alpha=1.0), dict(alpha=1.8), dict(rho=1.0), dict(rho=1e-3, sigma=1e-4), dict(eps_abs=1e-5, eps_rel=0.0),
]


@pytest.mark.parametrize("rate", [False, True])
@pytest.mark.parametrize("N", [10, 20])          # 10: two QPs share a warp, 20: one QP per warp
@pytest.mark.parametrize("variant", range(len(SETTINGS_VARIANTS)))
def test_settings_variants_match_oracle(pkg, oracle, workloads, variant, N, rate):
    # every OSQP knob the ABI exports (the reference leaves them at their defaults, mpc.cpp:98-99), on both kernel variants
    B = 63
    kw = dict(eps_abs=1e-4, eps_rel=1e-4, warm_start=0)
    kw.update(SETTINGS_VARIANTS[variant])
    recs = workloads.tracking_batch(B, N, seed=400 + variant, gaps=True)
    rd = 0.02 if rate else None
    g = pkg.MpcSolver(pkg.default_config(N, 2, rate_delta=rd), pkg.default_settings(**kw), B).solve_host(recs)
    o = oracle.MpcBatch(oracle.default_cfg(N, 2, rate_delta=rd), oracle.default_settings(**kw), B).solve(recs)
    assert_solution_parity(g, o, N)
    np.testing.assert_array_equal(g["iters"], o["iters"])


@pytest.mark.parametrize("rate", [False, True])
def test_problem_family_variants_match_oracle(pkg, oracle, workloads, rate):
    # weights, limits, set-points and dt other than params.yaml's
    N, B, eps = 24, 64, 1e-4
    recs = workloads.tracking_batch(B, N, seed=17)
    rd = 0.015 if rate else None
    c = pkg.default_config(N, 0, rate_delta=rd)
    oc = oracle.default_cfg(N, 0, rate_delta=rd)
    c.dt = oc[1] = 0.02
    for j, v in enumerate((3.0, 7.0, 0.5)):
        c.q[j] = oc[2 + j] = v
    for j, v in enumerate((1.0, 0.7)):
        c.r[j] = oc[5 + j] = v
    for j, (d, lo, hi) in enumerate(((3.0, 1.0, 6.0), (0.05, -0.3, 0.35))):
        c.u_des[j] = oc[7 + j] = d
        c.u_min[j] = oc[9 + j] = lo
        c.u_max[j] = oc[11 + j] = hi
    g = pkg.MpcSolver(c, pkg.default_settings(eps_abs=eps, eps_rel=eps, warm_start=0), B).solve_host(recs)
    o = oracle.MpcBatch(oc, oracle.default_settings(eps_abs=eps, eps_rel=eps, warm_start=0), B).solve(recs)
    assert_solution_parity(g, o, N)
    np.testing.assert_array_equal(g["iters"], o["iters"])


@pytest.mark.parametrize("N,B", [(3, 1), (3, 7), (7, 13), (10, 1), (10, 5), (15, 33), (12, 4096)])
def test_packed_short_horizons_ragged_batches_and_empty_slots(pkg, oracle, workloads, N, B):
    # N + 1 <= 16 / 8: 2 / 4 QPs share a warp.  Batch sizes that leave lane groups without a QP, QPs of one warp that stop at
    # different iterations (mixed easy / hard / infeasible problems), and empty slots (NaN linearisation speed) next to live ones.
    eps = 1e-4
    recs = workloads.tracking_batch(B, N, seed=500 + N + B, gaps=True)
    recs[::3, 4] = 0.9                                     # far outside the steering box: longer solves next to short ones
    empty = np.zeros(B, dtype=bool)
    if B > 2:
        empty[1::5] = True
    recs[empty, 3] = np.nan
    gm = 1 if B % 2 else 2                                  # gap_mode 1: mostly infeasible (all-ones stage-0 pair); 2: mostly solved
    g = pkg.MpcSolver(pkg.default_config(N, gm), pkg.default_settings(eps_abs=eps, eps_rel=eps, warm_start=0), B).solve_host(recs)
    live = ~empty
    o = oracle.MpcBatch(oracle.default_cfg(N, gm), oracle.default_settings(eps_abs=eps, eps_rel=eps, warm_start=0), int(live.sum())).solve(recs[live])
    assert (g["status"][empty] == pkg.UNSOLVED).all() and np.isnan(g["u0"][empty]).all() and (g["iters"][empty] == 0).all()
    gl = {k: v[live] for k, v in g.items() if isinstance(v, np.ndarray)}
    assert_solution_parity(gl, o, N)
    np.testing.assert_array_equal(gl["iters"], o["iters"])
    assert len(set(o["iters"].tolist())) > 1 or B < 30     # QPs sharing a warp do stop at different iterations
