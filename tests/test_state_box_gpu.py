"""GPU tests of the state-box rows (f110_mpc_config.state_rows): the CUDA solve against the oracle
(status, iterations, primal, dual) and against the KKT-certified exact optima of tests/golden/exactbox_*.npz."""
import glob
import os

import numpy as np
import pytest

from test_gpu_parity import assert_solution_parity

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "exactbox_*.npz")))


@pytest.mark.parametrize("N,lim", [(30, 1.0), (30, 1.25), (31, 1.0), (20, 0.7), (10, 0.4), (5, 0.25), (1, 0.1), (16, 0.6), (30, 50.0)])
@pytest.mark.parametrize("eps", [1e-3, 1e-4])
def test_state_box_matches_oracle(pkg, oracle, workloads, N, lim, eps):
    B = 96
    recs = workloads.tracking_batch(B, N, seed=600 + N)
    g = pkg.MpcSolver(pkg.default_config(N, 0, state_lim=lim), pkg.default_settings(eps_abs=eps, eps_rel=eps, warm_start=0), B).solve_host(recs)
    o = oracle.MpcBatch(oracle.default_cfg(N, 0, state_lim=lim), oracle.default_settings(eps_abs=eps, eps_rel=eps, warm_start=0), B).solve(recs)
    assert g["y"].shape[1] == 7 * N + 5 + 3 * (N + 1)
    assert_solution_parity(g, o, N)
    np.testing.assert_array_equal(g["iters"], o["iters"])
    if lim < 10 and N >= 10:
        ok = o["status"] == 1
        assert (np.abs(g["y"][ok][:, 7 * N + 5:]) > 1e-6).any()            # the box is active somewhere
        xs = g["x"][ok][:, :3 * (N + 1)].reshape(-1, N + 1, 3)
        assert (np.abs(xs[:, :, :2] - recs[ok][:, None, :2]) <= lim + 20 * eps).all()      # within the ADMM primal tolerance


def test_state_box_with_gap_rows_warm_start_and_infeasible(pkg, oracle, workloads):
    N, B = 30, 64
    recs = workloads.tracking_batch(B, N, seed=8, gaps=True)
    cfg_g, cfg_o = pkg.default_config(N, 2, state_lim=1.1), oracle.default_cfg(N, 2, state_lim=1.1)
    sol = pkg.MpcSolver(cfg_g, pkg.default_settings(warm_start=1), B)
    orc = oracle.MpcBatch(cfg_o, oracle.default_settings(warm_start=1), B)
    for step in range(3):
        g, o = sol.solve_host(recs), orc.solve(recs, warm=True)
        assert_solution_parity(g, o, N)
        np.testing.assert_array_equal(g["iters"], o["iters"])
        recs = recs.copy(); recs[:, 0] += 0.01; recs[:, 4] *= 0.9
    # a box the horizon cannot fit into: primal infeasible on both sides, NaN-filled
    g = pkg.MpcSolver(pkg.default_config(N, 0, state_lim=0.5), pkg.default_settings(warm_start=0), B).solve_host(recs)
    o = oracle.MpcBatch(oracle.default_cfg(N, 0, state_lim=0.5), oracle.default_settings(warm_start=0), B).solve(recs)
    assert (o["status"] == oracle.PRIMAL_INFEASIBLE).all()
    assert_solution_parity(g, o, N)


@pytest.mark.parametrize("name", NAMES)
def test_state_box_gpu_converges_to_the_exact_optimum(pkg, name):
    d = np.load(os.path.join(GOLD, name + ".npz"))
    N, lim = int(d["N"]), float(d["state_lim"])
    st = pkg.default_settings(eps_abs=1e-9, eps_rel=1e-9, warm_start=0, max_iter=40000)
    g = pkg.MpcSolver(pkg.default_config(N, 0, state_lim=lim), st, max_batch=len(d["recs"])).solve_host(d["recs"])
    assert (g["status"] == 1).all()
    np.testing.assert_allclose(g["x"], d["x"], atol=2e-6, rtol=0)
    np.testing.assert_allclose(g["y"], d["y"], atol=1e-4, rtol=1e-5)


def test_state_box_refusals_and_host_class(pkg, oracle, workloads):
    with pytest.raises(RuntimeError, match="state-box"):
        pkg.MpcSolver(pkg.default_config(40, 0, state_lim=1.0), pkg.default_settings(), 4)
    c = pkg.default_config(30, 0, state_lim=1.0)
    c.rate_rows, c.rate_delta = 1, 0.03
    with pytest.raises(RuntimeError, match="state-box"):
        pkg.MpcSolver(c, pkg.default_settings(), 4)
