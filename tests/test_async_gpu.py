"""GPU tests of the asynchronous cycle entry (f110_cycle_submit / f110_cycle_wait), of handle reuse across different
scene counts, of the NVLink gather ring and of the multi-GPU handle (f110_mpc_create_multi).

The asynchronous entry must return the same bits as f110_cycle_host (which the other cycle tests pin against the oracle
pipeline); the multi-GPU handle must return the same bits as the single-GPU host call on the same records."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

KEYS = ("u0", "status", "iters", "chosen", "valid")


def _scene(workloads, S, seed, sd=19):
    poses, _, scans = workloads.scene_batch(S, seed=seed)
    table = np.ascontiguousarray(workloads.traj_table(steer_discrete=sd)[:, :, :2])
    xy, _ = workloads.skirk_waypoints()
    return (np.ascontiguousarray(poses), np.ascontiguousarray(scans, dtype=np.float32), table,
            np.ascontiguousarray(xy, dtype=np.float32))


def _same(a, b):
    for k in KEYS:
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)


@pytest.mark.parametrize("qp_mode", [0, 2])
def test_two_cycles_in_flight_equal_two_synchronous_ones(pkg, workloads, qp_mode):
    S = 96
    P = 20
    nq = S if qp_mode == 0 else S * P
    cc = pkg.default_cycle_config(qp_mode=qp_mode)
    a, b = _scene(workloads, S, 501), _scene(workloads, S, 502)
    ref = pkg.MpcSolver(pkg.default_config(30), pkg.default_settings(warm_start=0), max_batch=nq)
    want = [ref.cycle_host(cc, s[0], s[1], None, s[2], s[3]) for s in (a, b, a)]
    sol = pkg.MpcSolver(pkg.default_config(30), pkg.default_settings(warm_start=0), max_batch=nq)
    # pageable inputs: staged through the handle's pinned memory, so the caller may overwrite them right after submit
    pa, ra = a[0].copy(), a[1].copy()
    t0 = sol.cycle_submit(cc, pa, ra, None, a[2], a[3])
    pa[:] = 0; ra[:] = 0
    t1 = sol.cycle_submit(cc, b[0], b[1], None, b[2], b[3])
    with pytest.raises(RuntimeError, match="already in flight"):
        sol.cycle_submit(cc, a[0], a[1], None, a[2], a[3])
    got0 = sol.cycle_wait(t0)
    t2 = sol.cycle_submit(cc, a[0], a[1], None, a[2], a[3])
    got1 = sol.cycle_wait(t1)
    got2 = sol.cycle_wait(t2)
    for g, w in zip((got0, got1, got2), want):
        _same(g, w)
    assert (got0["status"] == 1).sum() > 0
    with pytest.raises(RuntimeError, match="no such cycle"):
        sol.cycle_wait(t2)


def test_pinned_inputs_are_read_in_place(pkg, workloads):
    S = 64
    cc = pkg.default_cycle_config(qp_mode=0)
    a = _scene(workloads, S, 503)
    want = pkg.MpcSolver(pkg.default_config(30), pkg.default_settings(warm_start=0), max_batch=S).cycle_host(cc, a[0], a[1], None, a[2], a[3])
    pin = lambda x: torch.from_numpy(x).pin_memory().numpy()
    sol = pkg.MpcSolver(pkg.default_config(30), pkg.default_settings(warm_start=0), max_batch=S)
    prev = pin(np.zeros(S))
    got = sol.cycle_wait(sol.cycle_submit(cc, pin(a[0]), pin(a[1]), None, a[2], a[3]))
    _same(got, want)
    got = sol.cycle_wait(sol.cycle_submit(cc, pin(a[0]), pin(a[1]), prev, a[2], a[3]))   # explicit zero previous steering = the default
    _same(got, want)


def test_warm_started_async_sequence_equals_synchronous_sequence(pkg, workloads):
    # consecutive cycles of one handle share the warm-start slots: the solves must run in submission order
    S = 48
    cc = pkg.default_cycle_config(qp_mode=0)
    scenes = [_scene(workloads, S, 510 + i) for i in range(4)]
    ref = pkg.MpcSolver(pkg.default_config(30), pkg.default_settings(warm_start=1), max_batch=S)
    want = [ref.cycle_host(cc, s[0], s[1], None, s[2], s[3]) for s in scenes]
    sol = pkg.MpcSolver(pkg.default_config(30), pkg.default_settings(warm_start=1), max_batch=S)
    got, pending = [], []
    for s in scenes:
        pending.append(sol.cycle_submit(cc, s[0], s[1], None, s[2], s[3]))
        if len(pending) == 2:
            got.append(sol.cycle_wait(pending.pop(0)))
    got += [sol.cycle_wait(t) for t in pending]
    for g, w in zip(got, want):
        _same(g, w)
    assert any((g["iters"] != want[0]["iters"]).any() for g in got[1:])   # the warm start did change the iteration counts


def test_one_handle_serves_different_scene_counts(pkg, workloads):
    # the constant tables must stay where they were uploaded when a later call has fewer scenes / beams (round-1 advisor finding)
    cc = pkg.default_cycle_config(qp_mode=0)
    big, small = _scene(workloads, 64, 520), _scene(workloads, 8, 521)
    sol = pkg.MpcSolver(pkg.default_config(30), pkg.default_settings(warm_start=0), max_batch=64)
    sol.cycle_host(cc, big[0], big[1], None, big[2], big[3])
    got = sol.cycle_host(cc, small[0], small[1], None, small[2], small[3])
    fresh = pkg.MpcSolver(pkg.default_config(30), pkg.default_settings(warm_start=0), max_batch=64)
    _same(got, fresh.cycle_host(cc, small[0], small[1], None, small[2], small[3]))
    got = sol.cycle_wait(sol.cycle_submit(cc, big[0], big[1], None, big[2], big[3]))
    got_s = sol.cycle_wait(sol.cycle_submit(cc, small[0], small[1], None, small[2], small[3]))
    got_s2 = sol.cycle_wait(sol.cycle_submit(cc, small[0], small[1], None, small[2], small[3]))
    _same(got_s, fresh.cycle_host(cc, small[0], small[1], None, small[2], small[3]))
    _same(got_s2, got_s)
    _same(got, fresh.cycle_host(cc, big[0], big[1], None, big[2], big[3]))


def test_reset_is_ordered_before_the_next_solve(pkg, workloads):
    # warm-started solve, reset, solve again: the second solve must be the cold-start solve, iteration for iteration
    recs = workloads.tracking_batch(512, 30, seed=77)
    cold = pkg.MpcSolver(pkg.default_config(30), pkg.default_settings(warm_start=0), max_batch=512).solve_host(recs, want_xy=False)
    sol = pkg.MpcSolver(pkg.default_config(30), pkg.default_settings(warm_start=1), max_batch=512)
    first = sol.solve_host(recs, want_xy=False)
    np.testing.assert_array_equal(first["iters"], cold["iters"])
    warm = sol.solve_host(recs, want_xy=False)
    assert (warm["iters"] < cold["iters"]).any()
    sol.reset()
    again = sol.solve_host(recs, want_xy=False)
    np.testing.assert_array_equal(again["iters"], cold["iters"])
    np.testing.assert_array_equal(again["u0"], cold["u0"])


def test_gather_ring_two_ranks_on_one_device(pkg, workloads):
    # two "ranks" (two handles, two streams) of one process share a ring: each solve kernel stores its packed rows in the ring,
    # raises its flag, and the root's stream waits for both flags before it copies the slot out
    S, world, slots = 40, 2, 3
    cc = pkg.default_cycle_config(qp_mode=0)
    ring = pkg.GatherRing.create(0, world, S, slots)
    sols = [pkg.MpcSolver(pkg.default_config(30), pkg.default_settings(warm_start=0), max_batch=S) for _ in range(world)]
    for r, s in enumerate(sols):
        s.set_gather(ring.ptr, world, r, S, slots)
    ref = pkg.MpcSolver(pkg.default_config(30), pkg.default_settings(warm_start=0), max_batch=S)
    for cyc in range(5):   # more cycles than slots: the ring wraps
        sc = [_scene(workloads, S, 530 + 2 * cyc + r) for r in range(world)]
        want = [ref.cycle_host(cc, x[0], x[1], None, x[2], x[3]) for x in sc]
        order = (0, 1) if cyc % 2 else (1, 0)   # the root may be queued before or after its peer
        tick = {}
        for r in order:
            tick[r] = sols[r].cycle_submit(cc, sc[r][0], sc[r][1], None, sc[r][2], sc[r][3])
        gathered = np.empty((world, S, 4))
        got1 = sols[1].cycle_wait(tick[1])
        got0 = sols[0].cycle_wait(tick[0], gathered=gathered)
        _same(got0, want[0]); _same(got1, want[1])
        np.testing.assert_array_equal(sols[0].gathered_view(tick[0], world, S), gathered)    # zero-copy view of the same rows
        for r in range(world):
            np.testing.assert_array_equal(gathered[r, :, :2], want[r]["u0"])
            np.testing.assert_array_equal(gathered[r, :, 2].astype(np.int32), want[r]["status"])
            np.testing.assert_array_equal(gathered[r, :, 3].astype(np.int32), want[r]["iters"])
    with pytest.raises(RuntimeError, match="root"):
        sols[1].cycle_wait(sols[1].cycle_submit(cc, sc[1][0], sc[1][1], None, sc[1][2], sc[1][3]), gathered=np.empty((world, S, 4)))
    # detach: the rank that ran ahead must not leave the root waiting
    sols[0].set_gather(None, 0, 0, 0, 0); sols[1].set_gather(None, 0, 0, 0, 0)
    torch.cuda.synchronize()
    ring.close()


@pytest.mark.parametrize("unit", [1, 140])
def test_multi_gpu_handle_matches_single_gpu(pkg, workloads, unit):
    # every visible GPU (one on the default test box, more under gpurun --gpus N); shards are whole units
    n = torch.cuda.device_count()
    B = 140 * 9
    recs = workloads.tracking_batch(B, 30, seed=91)
    want = pkg.MpcSolver(pkg.default_config(30), pkg.default_settings(warm_start=0), max_batch=B).solve_host(recs, want_xy=False)
    m = pkg.MultiGpuSolver(pkg.default_config(30), pkg.default_settings(warm_start=0), max_batch=B, devices=list(range(n)))
    got = m.solve_host(recs, unit=unit)
    for k in ("u0", "status", "iters"):
        np.testing.assert_array_equal(got[k], want[k], err_msg=k)
    shards = m.last_shards()
    assert sum(c for _, c in shards) == B and all(f % unit == 0 and c % unit == 0 for f, c in shards)
    assert shards[0][0] == 0 and all(shards[i][0] + shards[i][1] == shards[i + 1][0] for i in range(n - 1))
    got2 = m.solve_host(recs[:unit * 3], unit=unit)     # fewer units than before, possibly fewer than GPUs
    np.testing.assert_array_equal(got2["u0"], want["u0"][:unit * 3])
    with pytest.raises(RuntimeError, match="multiple of the shard unit"):
        m.solve_host(recs[:unit * 3 + 1], unit=3 if unit == 1 else unit)
    m.close()


def test_four_cycles_in_flight_with_overlapping_solves(pkg, workloads):
    # cold-started solves of the base row set share nothing on the device, so consecutive cycles' solves are not ordered against
    # each other (f110_cycle_set_depth up to 4): same bits as one synchronous call per cycle, in any completion order
    S, P = 80, 20
    cc = pkg.default_cycle_config(qp_mode=2)
    scenes = [_scene(workloads, S, 540 + i) for i in range(7)]
    ref = pkg.MpcSolver(pkg.default_config(30), pkg.default_settings(warm_start=0), max_batch=S * P)
    want = [ref.cycle_host(cc, s[0], s[1], None, s[2], s[3]) for s in scenes]
    sol = pkg.MpcSolver(pkg.default_config(30), pkg.default_settings(warm_start=0), max_batch=S * P)
    sol.set_cycle_depth(4)
    pending, got = [], []
    for s in scenes:
        if len(pending) == 4:
            got.append(sol.cycle_wait(pending.pop(0)))
        pending.append(sol.cycle_submit(cc, s[0], s[1], None, s[2], s[3]))
    with pytest.raises(RuntimeError, match="in flight"):
        sol.set_cycle_depth(2)                      # only between cycles
    got += [sol.cycle_wait(t) for t in pending]
    for g, w in zip(got, want):
        _same(g, w)
    with pytest.raises(RuntimeError, match="depth must be"):
        sol.set_cycle_depth(5)
    sol.set_cycle_depth(1)                          # strictly one at a time
    t0 = sol.cycle_submit(cc, *scenes[0][:2], None, *scenes[0][2:])
    with pytest.raises(RuntimeError, match="already in flight"):
        sol.cycle_submit(cc, *scenes[1][:2], None, *scenes[1][2:])
    _same(sol.cycle_wait(t0), want[0])


def test_depth_four_keeps_warm_started_solves_in_order(pkg, workloads):
    S = 40
    cc = pkg.default_cycle_config(qp_mode=0)
    scenes = [_scene(workloads, S, 560 + i) for i in range(6)]
    ref = pkg.MpcSolver(pkg.default_config(30), pkg.default_settings(warm_start=1), max_batch=S)
    want = [ref.cycle_host(cc, s[0], s[1], None, s[2], s[3]) for s in scenes]
    sol = pkg.MpcSolver(pkg.default_config(30), pkg.default_settings(warm_start=1), max_batch=S)
    sol.set_cycle_depth(4)
    pending, got = [], []
    for s in scenes:
        if len(pending) == 4:
            got.append(sol.cycle_wait(pending.pop(0)))
        pending.append(sol.cycle_submit(cc, s[0], s[1], None, s[2], s[3]))
    got += [sol.cycle_wait(t) for t in pending]
    for g, w in zip(got, want):
        _same(g, w)
