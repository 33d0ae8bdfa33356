"""Behaviour a drop-in has to keep under load: bit-reproducible results, handles that do not interfere, batch sizes that do
not divide anything, and a host cycle whose pipelining does not change a single bit of the answer."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_results_are_bit_reproducible(pkg, workloads):
    N, B = 30, 1024
    recs = workloads.tracking_batch(B, N, seed=31, gaps=True)
    sol = pkg.MpcSolver(pkg.default_config(N, 2), pkg.default_settings(warm_start=0), B)
    a = sol.solve_host(recs)
    for _ in range(3):
        b = sol.solve_host(recs)
        for k in ("x", "y", "u0", "status", "iters"):
            np.testing.assert_array_equal(a[k], b[k])
    # batch composition does not matter either: every QP is solved by its own warp / lane group
    c = sol.solve_host(recs[::-1].copy())
    np.testing.assert_array_equal(a["x"], c["x"][::-1])
    d = sol.solve_host(recs[100:137])
    np.testing.assert_array_equal(a["x"][100:137], d["x"])


def test_two_handles_two_streams_do_not_interfere(pkg, workloads):
    import torch
    dev = torch.device("cuda:0")
    N1, N2, B = 30, 12, 512
    r1 = workloads.tracking_batch(B, N1, seed=41)
    r2 = workloads.tracking_batch(B, N2, seed=42)
    s1 = pkg.MpcSolver(pkg.default_config(N1), pkg.default_settings(warm_start=0), B)
    s2 = pkg.MpcSolver(pkg.default_config(N2, 0, rate_delta=0.03), pkg.default_settings(warm_start=0), B)
    ref1, ref2 = s1.solve_host(r1), s2.solve_host(r2)
    d1, d2 = torch.from_numpy(r1).to(dev), torch.from_numpy(r2).to(dev)
    x1 = torch.empty(B, s1.n, dtype=torch.float64, device=dev); x2 = torch.empty(B, s2.n, dtype=torch.float64, device=dev)
    st1 = torch.empty(B, dtype=torch.int32, device=dev); st2 = torch.empty(B, dtype=torch.int32, device=dev)
    it1 = torch.empty(B, dtype=torch.int32, device=dev); it2 = torch.empty(B, dtype=torch.int32, device=dev)
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    for _ in range(4):                       # interleaved launches on two streams
        s1.solve_device(d1, x1, None, None, st1, it1, None, None, stream=sa.cuda_stream)
        s2.solve_device(d2, x2, None, None, st2, it2, None, None, stream=sb.cuda_stream)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(x1.cpu().numpy(), ref1["x"]); np.testing.assert_array_equal(it1.cpu().numpy(), ref1["iters"])
    np.testing.assert_array_equal(x2.cpu().numpy(), ref2["x"]); np.testing.assert_array_equal(it2.cpu().numpy(), ref2["iters"])


@pytest.mark.parametrize("N,B", [(30, 20001), (10, 9999), (50, 1500)])
def test_large_ragged_batches(pkg, oracle, workloads, N, B):
    base = workloads.tracking_batch(257, N, seed=51)
    recs = base[np.arange(B) % 257]
    g = pkg.MpcSolver(pkg.default_config(N), pkg.default_settings(warm_start=0), B).solve_host(recs, want_xy=False)
    o = oracle.MpcBatch(oracle.default_cfg(N), oracle.default_settings(warm_start=0), 257).solve(base)
    idx = np.arange(B) % 257
    np.testing.assert_array_equal(g["status"], o["status"][idx])
    np.testing.assert_array_equal(g["iters"], o["iters"][idx])
    np.testing.assert_allclose(g["u0"], o["x"][idx][:, 3 * (N + 1):3 * (N + 1) + 2], atol=1e-4, rtol=1e-3)
    # identical records give identical bits wherever they sit in the batch
    np.testing.assert_array_equal(g["u0"][:257], g["u0"][257 * 3:257 * 4])


def test_cycle_host_pipelining_is_invisible(pkg, workloads):
    # f110_cycle_host splits large batches into two halves on two streams; the answer must not depend on the split.
    # (The chunk count is read from the environment at call time, so the comparison runs in fresh interpreters.)
    code = r'''
import importlib, sys, numpy as np
sys.path.insert(0, %r)
M = importlib.import_module("f110-mpc_b200"); W = importlib.import_module("f110-mpc_b200.workloads")
S = 131
poses, yaws, scans = W.scene_batch(S, seed=88)
table = np.ascontiguousarray(W.traj_table(steer_discrete=19)[:, :, :2]); xy, _ = W.skirk_waypoints()
sol = M.MpcSolver(M.default_config(30), M.default_settings(warm_start=1), max_batch=S * 20)
cc = M.default_cycle_config(qp_mode=1)
acc = []
for step in range(2):
    g = sol.cycle_host(cc, poses, scans, np.linspace(-0.1, 0.1, S), table, xy)
    acc += [g["u0"], g["status"], g["iters"], g["chosen"], g["valid"]]
np.savez(sys.argv[1], *acc)
''' % ROOT
    outs = []
    for chunks in ("1", "2", "5"):
        path = "/tmp/cyc_chunks_%s.npz" % chunks
        env = dict(os.environ, F110_CYCLE_CHUNKS=chunks)
        subprocess.run([sys.executable, "-c", code, path], check=True, env=env, timeout=300)
        outs.append(np.load(path))
    for o in outs[1:]:
        for k in outs[0].files:
            np.testing.assert_array_equal(outs[0][k], o[k])
    assert (outs[0]["arr_1"] == 1).sum() > 100
