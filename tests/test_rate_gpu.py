"""Steering-rate rows: N rows  delta_k - delta_{k-1} in [-D, D]  appended to the reference's
row set.  The reference has no such rows, so parity is against the oracle's own stacking of them
(oracle/f110_ref.hpp qp_build_structure) solved by the same OSQP restatement; same tolerances as test_gpu_parity."""
import numpy as np
import pytest

from test_gpu_parity import assert_solution_parity

pytestmark = pytest.mark.gpu
DELTA = 0.01   # rad per step: binds on most stages of the synthetic tracking problems


def _pair(pkg, oracle, N, B, eps, gap_mode=0, warm=0, delta=DELTA):
    g = pkg.MpcSolver(pkg.default_config(N, gap_mode, rate_delta=delta), pkg.default_settings(eps_abs=eps, eps_rel=eps, warm_start=warm), B)
    o = oracle.MpcBatch(oracle.default_cfg(N, gap_mode, rate_delta=delta), oracle.default_settings(eps_abs=eps, eps_rel=eps, warm_start=warm), B)
    return g, o


@pytest.mark.parametrize("N", [1, 2, 5, 10, 30, 31, 32, 50, 63])
def test_rate_rows_match_oracle(pkg, oracle, workloads, N):
    B, eps = 96, 1e-4
    recs = workloads.tracking_batch(B, N, seed=300 + N)
    gs, ob = _pair(pkg, oracle, N, B, eps)
    assert gs.m == 8 * N + 5
    g, o = gs.solve_host(recs), ob.solve(recs)
    assert_solution_parity(g, o, N)
    np.testing.assert_array_equal(g["iters"], o["iters"])
    # the rows do bind: steering steps sit on the limit, and the limit is respected
    ok = o["status"] == oracle.SOLVED
    steer = g["x"][ok][:, 3 * (N + 1):].reshape(-1, N, 2)[:, :, 1]
    step = np.diff(np.concatenate([recs[ok][:, 4:5], steer], axis=1), axis=1)
    assert np.abs(step).max() <= DELTA + 3e-3          # ADMM at eps 1e-4: feasible to about eps_rel * |z|
    if N >= 5:
        assert (np.abs(step) > DELTA - 1e-4).sum() > B
        assert np.abs(g["y"][ok][:, 7 * N + 5:]).max() > 1e-3     # non-zero multipliers on the rate rows


def test_rate_rows_default_eps_and_loose_limit(pkg, oracle, workloads):
    # OSQP's default tolerances; and a limit so wide the rows never bind -> same solution as without them
    N, B = 30, 128
    recs = workloads.tracking_batch(B, N, seed=77)
    gs, ob = _pair(pkg, oracle, N, B, 1e-3)
    g, o = gs.solve_host(recs), ob.solve(recs)
    assert_solution_parity(g, o, N)
    np.testing.assert_array_equal(g["iters"], o["iters"])
    gs, ob = _pair(pkg, oracle, N, B, 1e-5, delta=10.0)
    g, o = gs.solve_host(recs), ob.solve(recs)
    assert_solution_parity(g, o, N)
    base = pkg.MpcSolver(pkg.default_config(N), pkg.default_settings(eps_abs=1e-5, eps_rel=1e-5, warm_start=0), B).solve_host(recs)
    # (the heading has zero weight, q2 = 0, so it is only weakly determined: compare positions and inputs tighter than headings)
    np.testing.assert_allclose(g["x"], base["x"], atol=1e-2)
    np.testing.assert_allclose(g["u0"], base["u0"], atol=1e-2)
    assert np.abs(g["y"][:, 7 * N + 5:]).max() < 1e-6


def test_rate_rows_with_gaps_and_warm_start(pkg, oracle, workloads):
    N, B, eps = 30, 128, 1e-4
    recs = workloads.tracking_batch(B, N, seed=78, gaps=True)
    gs, ob = _pair(pkg, oracle, N, B, eps, gap_mode=1, warm=1, delta=0.02)
    seen = set()
    for step in range(3):
        g, o = gs.solve_host(recs), ob.solve(recs, warm=True)
        assert_solution_parity(g, o, N)
        np.testing.assert_array_equal(g["iters"], o["iters"])
        seen |= set(o["status"].tolist())
        recs = recs.copy(); recs[:, 0] += 0.01; recs[:, 4] = np.where(o["status"] > 0, o["x"][:, 3 * (N + 1) + 1], recs[:, 4])
    assert oracle.SOLVED in seen


def test_rate_rows_infeasible_box(pkg, oracle, workloads):
    # previous steering far outside the input box: row 0 (|delta_0 - steer_prev| <= D) contradicts the box -> primal infeasible
    N, B, eps = 30, 32, 1e-4
    recs = workloads.tracking_batch(B, N, seed=79)
    recs[::2, 4] = 0.9
    gs, ob = _pair(pkg, oracle, N, B, eps)
    g, o = gs.solve_host(recs), ob.solve(recs)
    assert (o["status"][::2] == oracle.PRIMAL_INFEASIBLE).all() and (o["status"][1::2] == oracle.SOLVED).all()
    assert_solution_parity(g, o, N)


def test_rate_rows_refused_above_63(pkg):
    with pytest.raises(RuntimeError):
        pkg.MpcSolver(pkg.default_config(64, 0, rate_delta=DELTA), pkg.default_settings(), 4)


def test_host_mpc_with_rate_limit(pkg, oracle, workloads):
    # the C++ MPC class with Params::steer_rate_max: same QP as the oracle's with rate_delta = steer_rate_max * dt
    N = 30
    rate = 1.2                                    # rad/s
    dt = float(np.float32(0.01))
    rec = workloads.tracking_batch(1, N, seed=5)[0]
    m = pkg.HostMPC(N, 0, 0, steer_rate_max=rate)
    r = m.update(rec[0:3], rec[3:5], rec[11:].reshape(N, 3))
    o = oracle.MpcBatch(oracle.default_cfg(N, 0, rate_delta=rate * dt), oracle.default_settings(), 1, 1).solve(rec[None, :], warm=True)
    assert r["status"] == o["status"][0] == oracle.SOLVED
    np.testing.assert_allclose(r["x"], o["x"][0], atol=1e-4, rtol=1e-4)
    np.testing.assert_allclose(r["y"], o["y"][0], atol=1e-4, rtol=1e-4)
    assert r["y"].shape[0] == 8 * N + 5
    m.close()


@pytest.mark.parametrize("name", ["qprate_N30_delta0.01.npz", "qprate_N12_delta0.02.npz"])
def test_golden_rate_vectors(pkg, name):
    # committed fixtures (scripts/make_golden.py): N = 12 runs two QPs per warp, N = 30 one
    import os
    gd = np.load(os.path.join(os.path.dirname(__file__), "golden", name))
    N, delta, eps = int(gd["N"]), float(gd["rate_delta"]), float(gd["eps"])
    B = gd["recs"].shape[0]
    g = pkg.MpcSolver(pkg.default_config(N, 0, rate_delta=delta), pkg.default_settings(eps_abs=eps, eps_rel=eps, warm_start=0), B).solve_host(gd["recs"])
    assert_solution_parity(g, dict(status=gd["status"], x=gd["x"], y=gd["y"]), N)
    np.testing.assert_array_equal(g["iters"], gd["iters"])
