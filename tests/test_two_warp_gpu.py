"""Two-warp QPs (horizons 32..63) beyond the sizes of test_gpu_parity / test_rate_gpu: the partitioned solve of the base row set
(each warp reduces its own 32 stages, spike correction from one boundary exchange) at horizons that leave warp 1 nearly empty or
full, with odd batches (a CTA with one idle QP slot), gap rows with infeasible QPs, and warm-started sequences; the persistent
steering-rate kernel on batches with more QPs than resident QP slots (2 CTAs x 2 slots x 148 SMs = 592: slots solve several QPs in
turn).  Same tolerances as test_gpu_parity; iteration counts and statuses must equal the oracle's."""
import numpy as np
import pytest

from test_gpu_parity import assert_solution_parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,B", [(33, 77), (40, 5), (47, 131), (62, 77)])
def test_partitioned_solve_odd_batches(pkg, oracle, workloads, N, B):
    eps = 1e-4
    recs = workloads.tracking_batch(B, N, seed=700 + N)
    g = pkg.MpcSolver(pkg.default_config(N), pkg.default_settings(eps_abs=eps, eps_rel=eps, warm_start=0), B).solve_host(recs)
    o = oracle.MpcBatch(oracle.default_cfg(N), oracle.default_settings(eps_abs=eps, eps_rel=eps, warm_start=0), B).solve(recs)
    assert_solution_parity(g, o, N)
    np.testing.assert_array_equal(g["iters"], o["iters"])


@pytest.mark.parametrize("N", [33, 63])
def test_partitioned_solve_gap_rows_and_warm_start(pkg, oracle, workloads, N):
    # gap_mode 1: half-plane rows on every stage, some QPs primal infeasible (certificate path), rho adapts and re-factors
    B, eps = 96, 1e-4
    recs = workloads.tracking_batch(B, N, seed=720 + N, gaps=True)
    sol = pkg.MpcSolver(pkg.default_config(N, 1), pkg.default_settings(eps_abs=eps, eps_rel=eps, warm_start=1), B)
    mb = oracle.MpcBatch(oracle.default_cfg(N, 1), oracle.default_settings(eps_abs=eps, eps_rel=eps, warm_start=1), B)
    for _ in range(3):
        g, o = sol.solve_host(recs), mb.solve(recs, warm=True)
        assert_solution_parity(g, o, N)
        np.testing.assert_array_equal(g["iters"], o["iters"])


@pytest.mark.parametrize("N,B", [(40, 77), (62, 700), (33, 1500)])
def test_persistent_rate_slots(pkg, oracle, workloads, N, B):
    eps, delta = 1e-3, 0.02
    recs = workloads.tracking_batch(B, N, seed=740 + N)
    sol = pkg.MpcSolver(pkg.default_config(N, 0, rate_delta=delta), pkg.default_settings(eps_abs=eps, eps_rel=eps, warm_start=1), B)
    mb = oracle.MpcBatch(oracle.default_cfg(N, 0, rate_delta=delta), oracle.default_settings(eps_abs=eps, eps_rel=eps, warm_start=1), B)
    for _ in range(2):   # cold, then warm-started from the slots' stored iterates
        g, o = sol.solve_host(recs), mb.solve(recs, warm=True)
        assert_solution_parity(g, o, N)
        np.testing.assert_array_equal(g["iters"], o["iters"])
