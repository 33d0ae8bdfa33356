"""CPU tests of the state-box rows: the oracle's stacking of the box the reference stores but never
stacks (constraints.cpp:14-17, Constraints::SetXLims :108-114) against an independent numpy assembly and the KKT-certified
exact optima in tests/golden/exactbox_*.npz."""
import glob
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import qp_exact as E  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")
NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "exactbox_*.npz")))


def test_row_layout(oracle, workloads):
    N, d = 12, 0.8
    rec = workloads.tracking_batch(1, N, seed=3)[0]
    cfg = oracle.default_cfg(N, 0, state_lim=d)
    assert oracle.mpc_rows(cfg) == 7 * N + 5 + 3 * (N + 1)
    P, q, A, l, u = oracle.mpc_assemble_dense(cfg, rec)
    Pe, qe, Ae, le, ue = E.assemble(rec, N, 0, state_lim=d)
    for a, b in ((P, Pe), (q, qe), (A, Ae), (l, le), (u, ue)):
        np.testing.assert_array_equal(a, b)
    r0 = 7 * N + 5
    np.testing.assert_array_equal(A[r0:, :3 * (N + 1)], np.eye(3 * (N + 1)))
    assert (A[r0:, 3 * (N + 1):] == 0).all()
    for k in range(N + 1):
        assert l[r0 + 3 * k] == rec[0] - d and u[r0 + 3 * k] == rec[0] + d and l[r0 + 3 * k + 1] == rec[1] - d and u[r0 + 3 * k + 1] == rec[1] + d
        assert l[r0 + 3 * k + 2] == -1e30 and u[r0 + 3 * k + 2] == 1e30
    # the rows above are untouched
    P0, q0, A0, l0, u0 = oracle.mpc_assemble_dense(oracle.default_cfg(N, 0), rec)
    np.testing.assert_array_equal(A[:r0], A0); np.testing.assert_array_equal(l[:r0], l0); np.testing.assert_array_equal(u[:r0], u0)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_converges_to_the_exact_optimum_with_state_box(oracle, name):
    d = np.load(os.path.join(GOLD, name + ".npz"))
    N, lim = int(d["N"]), float(d["state_lim"])
    assert (np.abs(d["y"][:, 7 * N + 5:]) > 1e-9).any()                 # the box binds somewhere in the sample
    for i in range(0, len(d["recs"]), 8):
        P, q, A, l, u = E.assemble(d["recs"][i], N, 0, state_lim=lim)
        stat, feas, sign = E.kkt_residuals(P, q, A, l, u, d["x"][i], d["y"][i])
        assert stat < 1e-8 and feas < 1e-9 and sign < 1e-9
    st = oracle.default_settings(eps_abs=1e-9, eps_rel=1e-9, warm_start=0, max_iter=40000)
    o = oracle.MpcBatch(oracle.default_cfg(N, 0, state_lim=lim), st, len(d["recs"])).solve(d["recs"])
    assert (o["status"] == 1).all()
    np.testing.assert_allclose(o["x"], d["x"], atol=2e-6, rtol=0)
    np.testing.assert_allclose(o["y"], d["y"], atol=1e-4, rtol=1e-5)


def test_infeasible_box(oracle, workloads):
    # 0.9 m is the least the car can travel in 30 steps at the 3 m/s speed floor: a 0.5 m box cannot hold the horizon
    recs = workloads.tracking_batch(8, 30, seed=4)
    o = oracle.MpcBatch(oracle.default_cfg(30, 0, state_lim=0.5), oracle.default_settings(warm_start=0), 8).solve(recs)
    assert (o["status"] == oracle.PRIMAL_INFEASIBLE).all() and np.isnan(o["x"]).all()
