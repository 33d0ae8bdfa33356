#!/usr/bin/env python
"""bench.py — batched per-cycle MPC solve throughput on B200 (metric of BASELINE.json).

One "step" = one pass of the hot path over one batch: the mini-path collision check of every candidate
path of every scene (1 kernel) + the ADMM solve of one tracking QP per candidate path (1 kernel), 4096 QPs
per GPU, N = 30 (params.yaml horizon), OSQP default settings (eps 1e-3 — what the reference runs,
mpc.cpp:98-99), cold start.  Weak scaling: every rank owns its own 4096 QPs; the only exchange is the final
gather of (u0, status, iters) on rank 0's GPU — the solve kernels store their packed rows there themselves
over NVLink (CUDA-IPC-mapped ring, f110_gather_*); no collective launch on the step.

  python bench.py [--gpus N --steps K --warmup W]        product arm (torchrun for N > 1)
  python bench.py --impl reference [...]                 CPU arm: the oracle (OSQP restatement; the OSQP
                                                         binary is not available) on all host threads
Prints ONE JSON line on rank 0.
"""
import argparse
import importlib
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_HORIZON = 30
QPS_PER_GPU = 4096
E2E_MIN_CYCLES = 200   # the e2e leg is timed by the host clock: at least this many cycles (45 ms), so that one scheduling hiccup is not the number
PIPELINE_DEPTH = int(os.environ.get("F110_BENCH_DEPTH", "4"))     # streams the steps of the timed region alternate over / cycles in flight of the e2e leg
PATHS = 20            # README.md:12 count (steer_discrete = 19)
SAMPLES = 50
METRIC = "batched MPC QP solves/sec"
UNIT = "solves/s"


def flops_per_qp(N, iters, rho_updates):
    """Algorithmic flops, reference formulation."""
    f_iter = 358 * N + 158
    f_check = 202 * N + 102
    f_factor = 290 * N
    return iters * f_iter + np.ceil(iters / 25.0) * f_check + (1 + rho_updates) * f_factor


def hbm_bytes_per_qp(N):
    """Algorithmic HBM bytes: parameter record in, (u0, status, iters) out."""
    return 8 * (11 + 3 * N) + 16 + 8


def load_workloads():
    """The synthetic-workload generators (f110-mpc_b200/workloads.py: numpy only), loaded by path so that the reference arm does
    not import the product package."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("f110_workloads", os.path.join(ROOT, "f110-mpc_b200", "workloads.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def build_workload(M, W, rank):
    """4096 QPs = 205 scenes x 20 mini-paths (truncated): ego on a skirk waypoint with a small tracking error,
    reference = mini-path p in the world frame; plus (product arm: M given) the scenes' occupancy grids for the collision check.
    With M = None only numpy is used: that is what the reference arm calls."""
    table = W.traj_table(steer_discrete=PATHS - 1, traj_discrete=SAMPLES)       # (20, 50, 3)
    S = math.ceil(QPS_PER_GPU / PATHS)
    rng = np.random.default_rng(20240901 + 1000 * rank)
    poses, yaws, scans = W.scene_batch(S, seed=20240902 + 1000 * rank)
    recs = np.zeros((S * PATHS, W.record_doubles(N_HORIZON)))
    grids = np.zeros((S, 100 * 100), dtype=np.float32)
    offs = np.zeros((S, 2), dtype=np.float32)
    rots = np.zeros((S, 4))
    for s in range(S):
        x, y, yaw = poses[s, 0], poses[s, 1], yaws[s]
        c, sn = np.cos(yaw), np.sin(yaw)
        if M is not None:
            # occupancy grid and tf2 rotation from the product's C++ host classes (OccGrid::FillOccGrid, Transforms)
            grids[s], offs[s] = M.host_fill_grid(poses[s], W.SCAN_ANGLE_MIN, W.SCAN_ANGLE_MAX, W.SCAN_ANGLE_INC, scans[s])
            rots[s] = M.host_car_to_world_R(poses[s])
        lat, dyaw = rng.uniform(-0.3, 0.3), rng.uniform(-0.2, 0.2)
        x0 = np.array([x - sn * lat, y + c * lat, yaw + dyaw])
        for p in range(PATHS):
            r = recs[s * PATHS + p]
            r[0:3] = x0
            r[3] = 4.5
            r[4] = rng.uniform(-0.4, 0.4)
            r[5:8] = (0.3, -0.8, 1.5)
            r[8:11] = (-0.4, 0.7, 2.0)
            ref = np.zeros((N_HORIZON, 3))
            ref[:, :2] = W.path_to_world(table[p, :N_HORIZON, :2], x, y, yaw)
            r[11:] = ref.reshape(-1)
    return dict(recs=recs[:QPS_PER_GPU], grids=grids, offs=offs, rots=rots, pose_xy=poses[:, :2].copy(),
                table_xy=np.ascontiguousarray(table[:, :, :2]), scenes=S, poses=poses, scans=scans)


def init_nccl(dist, torch, dev):
    """init_process_group + a first collective with file descriptor 1 pointed at stderr: NCCL prints its version banner on
    stdout when the communicator comes up, and stdout carries exactly one JSON line."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        dist.init_process_group("nccl", device_id=dev)
        dist.all_reduce(torch.zeros(1, device=dev))
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


def time_pipelined(torch, streams, n, launch):
    """n independent launches, launch(i, stream) on stream i mod len(streams); device time from a fork event every stream waits
    for to a join event recorded after every stream's last launch (ms)."""
    main = torch.cuda.current_stream()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(main)
    for st in streams:
        st.wait_event(a)
    for i in range(n):
        launch(i, streams[i % len(streams)])
    for st in streams:
        e = torch.cuda.Event()
        e.record(st)
        main.wait_event(e)
    b.record(main)
    torch.cuda.synchronize()
    return float(a.elapsed_time(b))


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region, sampled through NVML from a thread (the region lasts a few
    milliseconds, too short for `nvidia-smi -lms`); falls back to one nvidia-smi query when NVML is unavailable."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and vis.split(",")[self.gpu].isdigit() else self.gpu
            self._h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self._nvml = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        except Exception:
            self._nvml = None

    def _run(self):
        nv = self._nvml
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                r = get_reasons(self._h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                break
            time.sleep(0.0005)

    def stop(self):
        if self._nvml is None:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True).stdout.split(",")
                return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "reasons": [], "samples": 1, "how": "nvidia-smi after the region"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock query unavailable"], "samples": 0}
        self._stop.set()
        self._thread.join(timeout=1.0)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.sm), "how": "NVML polled from a thread during the timed region"}


def cpu_reference_run(recs, steps, warmup, nthreads=0):
    """The reference's CPU path for the same QPs: oracle (OSQP-algorithm restatement) on host threads."""
    from oracle import oracle_py as O
    O.build()
    mb = O.MpcBatch(O.default_cfg(N_HORIZON), O.default_settings(warm_start=0), len(recs), nthreads)
    for _ in range(warmup):
        mb.solve(recs, want_xy=False)
    secs = []
    for _ in range(steps):
        secs.append(mb.solve(recs, want_xy=False)["seconds"])
    return mb.threads, secs


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # nothing of the product is loaded here: records from the numpy generators, solves by oracle/ on the host cores
    wl = build_workload(None, load_workloads(), 0)
    threads, secs = cpu_reference_run(wl["recs"], args.steps, args.warmup)
    tot = float(np.sum(secs))
    value = QPS_PER_GPU * args.steps / tot
    sample = "%d x the %d-QP batch (all QPs of the workload), one QP per thread at a time" % (args.steps, QPS_PER_GPU)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(wl),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                             "threads": "std::thread per core, not pinned (the scheduler spreads them; 15.8x on 16 threads measured)",
                             "note": "OSQP-algorithm restatement (oracle/), not the OSQP binary (unavailable offline)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def workload_config(wl):
    return {"workload": "cfg2x4096: %d skirk scenes x %d mini-paths (first %d), grid collision check of every path + one "
                        "N=%d tracking QP per path, OSQP defaults (eps 1e-3), cold start" % (wl["scenes"], PATHS, QPS_PER_GPU, N_HORIZON),
            "qps_per_gpu": QPS_PER_GPU, "horizon": N_HORIZON, "paths": PATHS, "samples": SAMPLES, "scenes_per_gpu": wl["scenes"],
            "eps_abs": 1e-3, "eps_rel": 1e-3,
            "l2_policy": "inputs larger than L2: 24 copies of the step's records + occupancy grids at distinct addresses (278 MB), step i reads copy i mod 24; no flush",
            "pipelining": "steps are independent batches: step i runs on stream i mod %d, consecutive solves overlap (extra.isolated: one step at a time with an L2 flush)" % PIPELINE_DEPTH,
            "parallelism": "independent QPs sharded by rank; final gather of (u0,status,iters) on rank 0's GPU by the solve kernels' own "
                           "NVLink stores (IPC-mapped ring), NCCL all-gather only as fallback"}


class PeerGather:
    """All ranks' packed rows (u0_v, u0_steer, status, iters) in one ring on rank 0's GPU, written by the solve kernels
    themselves over NVLink: rank 0 allocates the ring (f110_gather_create), the 64-byte CUDA IPC handle travels through
    torch.distributed, the other ranks map it (f110_gather_open).  `ok` is False on every rank if any rank failed to map it;
    the caller then falls back to an NCCL all-gather and says so in its JSON line."""

    def __init__(self, M, dist, torch, dev, world, rank, local, rows, slots):
        self.ring, self.ok, self.why = None, True, ""
        handle = [None]
        if rank == 0:
            try:
                self.ring = M.GatherRing.create(local, world, rows, slots)
                handle[0] = self.ring.handle
            except Exception as e:          # noqa: BLE001 - any failure means "fall back"
                self.ok, self.why = False, str(e)
        dist.broadcast_object_list(handle, src=0)
        if rank != 0:
            if handle[0] is None:
                self.ok = False
            else:
                try:
                    self.ring = M.GatherRing.open(handle[0], local, world, rank, rows, slots)
                except Exception as e:      # noqa: BLE001
                    self.ok, self.why = False, str(e)
        flag = torch.tensor([1 if self.ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        self.ok = bool(flag.item())

    def close(self):
        if self.ring is not None:
            self.ring.close()
            self.ring = None


def ring_slot_to_host(torch, dev, gather, cycle, world, rows):
    """rank 0: the ring slot of `cycle` as a (world, rows, 4) numpy array"""
    ptr, _ = gather.ring.slot(cycle)      # rank 0's block comes first, the other ranks' blocks follow

    class _Dev:   # the raw ring pointer as a torch tensor (CUDA array interface)
        __cuda_array_interface__ = {"shape": (world * rows * 4,), "typestr": "<f8", "data": (ptr, False), "version": 2}
    return torch.as_tensor(_Dev(), device=dev).cpu().numpy().reshape(world, rows, 4)


def main_product(args):
    import torch
    import torch.distributed as dist
    M = importlib.import_module("f110-mpc_b200")
    W = importlib.import_module("f110-mpc_b200.workloads")
    SH = importlib.import_module("f110-mpc_b200.sharding")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py product arm needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        init_nccl(dist, torch, dev)
    M.build()
    wl = build_workload(M, W, rank)
    B, S = QPS_PER_GPU, wl["scenes"]
    NQ = S * PATHS                       # QPs per step of the e2e leg (every candidate path of every scene)
    warmup = max(args.warmup, 3)
    sol = M.MpcSolver(M.default_config(N_HORIZON), M.default_settings(warm_start=0), max_batch=B, device=local)
    # ---- device-resident inputs / outputs
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    # records padded to an even stride: 16-byte aligned rows, which the kernel stages with one TMA bulk copy each
    wl["recs_padded"] = np.pad(wl["recs"], ((0, 0), (0, wl["recs"].shape[1] % 2)))
    d_recs, d_grid, d_off, d_rot, d_pose, d_tab = (t(wl[k]) for k in ("recs_padded", "grids", "offs", "rots", "pose_xy", "table_xy"))
    # L2 policy of the timed region: inputs larger than L2.  COPIES copies of the step's large inputs (records + occupancy grids,
    # 11.6 MB) at distinct addresses, 278 MB in all (L2 is 126 MB); step i reads copy i mod COPIES, so no step finds its inputs in L2.
    COPIES, DEPTH = 24, PIPELINE_DEPTH
    recs_c = [d_recs] + [d_recs.clone() for _ in range(COPIES - 1)]
    grid_c = [d_grid] + [d_grid.clone() for _ in range(COPIES - 1)]
    input_bytes = COPIES * (d_recs.numel() * 8 + d_grid.numel() * 4)
    # one output set per stream: the steps of the timed region alternate over DEPTH streams
    outs = [dict(u0=torch.empty(B, 2, dtype=torch.float64, device=dev), status=torch.empty(B, dtype=torch.int32, device=dev),
                 iters=torch.empty(B, dtype=torch.int32, device=dev), rhoup=torch.empty(B, dtype=torch.int32, device=dev),
                 valid=torch.empty(S, PATHS, dtype=torch.uint8, device=dev), free=torch.empty(S, PATHS, dtype=torch.int32, device=dev),
                 endw=torch.empty(S, PATHS, 2, dtype=torch.float32, device=dev)) for _ in range(DEPTH)]
    d_u0, d_status, d_iters, d_rhoup = (outs[0][k] for k in ("u0", "status", "iters", "rhoup"))
    d_valid, d_free, d_endw = (outs[0][k] for k in ("valid", "free", "endw"))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    main_stream = torch.cuda.current_stream()
    stream = main_stream.cuda_stream
    streams = [torch.cuda.Stream(device=dev) for _ in range(DEPTH)]
    # ---- the gather: a ring on rank 0's GPU that every rank's solve kernel stores into (no collective on the step)
    # ring depth: a rank may run ahead of rank 0's reads by fewer than SLOTS cycles, so the ranks re-align (one host barrier) every
    # SLOTS e2e cycles; deep enough that a default run never needs to (224 slots x 8 ranks x 4100 rows x 32 B = 235 MB on rank 0)
    SLOTS = min(1024, max(64, 8 * ((max(2 * args.steps, E2E_MIN_CYCLES) + max(args.warmup, 3) + 2 * PIPELINE_DEPTH + 15) // 8)))
    gather = PeerGather(M, dist, torch, dev, world, rank, local, NQ, SLOTS) if world > 1 else None
    peer = gather is not None and gather.ok
    d_packed = torch.empty(B, 4, dtype=torch.float64, device=dev) if (world > 1 and not peer) else None   # NCCL fallback
    cyc = [0]

    def solve_step(recs, o, st):
        if peer:
            rows, _ = gather.ring.slot(cyc[0])
            cyc[0] += 1
            M._check(M.lib().f110_mpc_set_packed_output(sol._h, rows), "f110_mpc_set_packed_output")
            sol.solve_device(recs, None, None, o["u0"], o["status"], o["iters"], o["rhoup"], None, stream=st.cuda_stream)
        else:
            sol.solve_device(recs, None, None, o["u0"], o["status"], o["iters"], o["rhoup"], None, stream=st.cuda_stream, packed=d_packed)
            if world > 1:   # fallback: one all-gather of every rank's rows
                with torch.cuda.stream(st):
                    SH.gather_results(d_packed, world, max_rows=B, sizes=[B] * world)

    def step(i, st=None):
        """one step: collision check of every candidate path + the solve, on stream `st` (default: step i's stream of the pipeline)"""
        j = i % DEPTH
        st = st or streams[j]
        o = outs[j]
        M.collision_check_device(grid_c[i % COPIES], d_off, d_rot, d_pose, d_tab, o["valid"], o["free"], o["endw"], stream=st.cuda_stream)
        solve_step(recs_c[i % COPIES], o, st)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- timed region: EXACTLY K steps, bracketed by barrier + synchronize on both sides and by two CUDA events on the device.
    # Steps are independent batches: step i runs on stream i mod DEPTH, so the next step's kernels take the SMs that the tail of
    # the previous solve leaves idle (persistent CTAs retire one by one).  Device time = fork event on the main stream, which every
    # worker stream waits for, to the join event the main stream records after waiting for every worker stream's last launch.
    ev_fork, ev_join = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev_fork.record(main_stream)
    for st in streams:
        st.wait_event(ev_fork)
    for i in range(args.steps):
        step(i)
    for st in streams:
        e = torch.cuda.Event()
        e.record(st)
        main_stream.wait_event(e)
    ev_join.record(main_stream)
    barrier()
    step_ms = float(ev_fork.elapsed_time(ev_join))
    # ---- the same K steps one at a time on one stream, 256 MiB L2 flush before each, CUDA events around each step and around
    # each solve launch: the kernel timed alone (roofline), and the round-1 / round-2a definition of `value` (extra.isolated)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        ev[i][0].record()
        M.collision_check_device(d_grid, d_off, d_rot, d_pose, d_tab, d_valid, d_free, d_endw, stream=stream)
        kev[i][0].record()
        solve_step(d_recs, outs[0], main_stream)
        kev[i][1].record()
        ev[i][1].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    iso_ms = float(sum(a.elapsed_time(b) for a, b in ev))
    admm_ms = float(sum(a.elapsed_time(b) for a, b in kev))
    tt = torch.tensor([step_ms, admm_ms, iso_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    step_ms, admm_ms, iso_ms = tt.tolist()
    value = world * B * args.steps / (step_ms * 1e-3)
    isolated = {"value": world * B * args.steps / (iso_ms * 1e-3), "unit": UNIT, "ms_per_step": iso_ms / args.steps,
                "admm_ms_per_launch": admm_ms / args.steps,
                "what": "the same steps one at a time on one stream, 256 MiB L2 flush before each (how `value` was defined up to round 2a)"}

    iters = d_iters.cpu().numpy()
    rhoup = d_rhoup.cpu().numpy()
    status = d_status.cpu().numpy()
    flops_launch = float(flops_per_qp(N_HORIZON, iters, rhoup).sum())
    admm_s_per_launch = admm_ms * 1e-3 / args.steps
    gathered_check = None
    if peer:
        # what the last timed step left in the ring on rank 0's GPU: every rank's 4096 rows, delivered by the kernels' own stores
        ok_local = torch.tensor([int((status == 1).sum())], dtype=torch.int64, device=dev)
        all_ok = [torch.zeros_like(ok_local) for _ in range(world)]
        dist.all_gather(all_ok, ok_local)
        if rank == 0:
            host = ring_slot_to_host(torch, dev, gather, cyc[0] - 1, world, NQ)
            gathered_check = {"rows_solved_per_rank": [int((host[r, :B, 2] == 1).sum()) for r in range(world)],
                              "solved_per_rank_reported": [int(x.item()) for x in all_ok],
                              "rank0_rows_equal_local": bool(np.array_equal(host[0, :B, :2], d_u0.cpu().numpy()))}
            gathered_check["ok"] = gathered_check["rows_solved_per_rank"] == gathered_check["solved_per_rank_reported"] and gathered_check["rank0_rows_equal_local"]

    if args.skip_extras:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                              "ms_per_step": step_ms / args.steps, "admm_ms_per_launch": admm_ms / args.steps, "isolated": isolated["value"],
                              "note": "--skip-extras run (profiling target), not a bench value"}))
        if gather:
            gather.close()
        if world > 1:
            dist.destroy_process_group()
        return 0
    # ---- e2e: the reference-facing host-buffer call for the whole cycle, in its asynchronous form (f110_cycle_submit /
    # f110_cycle_wait): laser scans + poses in pinned host memory -> grid fill, collision check, gap finder, selection, record
    # build, one QP per candidate path -> controls back on the host.  DEPTH cycles in flight: step k+1's copies and perception
    # kernels run under step k's solve, and (single GPU: cold-started solves share nothing) its solve fills the previous solve's
    # tail; every step's inputs are copied in and every step's results are copied out inside the timed region.  With N > 1 every rank's packed rows go to the ring on rank 0's GPU and rank 0 copies ALL of them to its host
    # buffer each step (it waits for the other ranks' flags first), so the timed region ends with every control on rank 0's host.
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    sol_e = M.MpcSolver(M.default_config(N_HORIZON), M.default_settings(warm_start=0), max_batch=NQ, device=local)
    cc = M.default_cycle_config(qp_mode=2, use_half_spaces=1)
    h_pose, h_scan = pin(wl["poses"]), pin(wl["scans"])
    h_tab, h_wp = np.ascontiguousarray(wl["table_xy"]), np.ascontiguousarray(W.skirk_waypoints()[0], dtype=np.float32)
    out = {"u0": pin(np.empty((NQ, 2))), "status": pin(np.empty(NQ, dtype=np.int32)), "iters": pin(np.empty(NQ, dtype=np.int32)),
           "chosen": pin(np.empty(S, dtype=np.int32)), "valid": pin(np.empty((S, PATHS), dtype=np.uint8))}
    h_gathered = None          # rank 0 reads the gathered rows in place (f110_cycle_gathered_view): no second host copy
    last_view = [None]
    e2e_gather = peer
    if e2e_gather:
        barrier()
        sol_e.set_gather(gather.ring.ptr, world, rank, NQ, SLOTS)

    def finish(tk):
        sol_e.cycle_wait(tk, out=out)
        if e2e_gather and rank == 0:     # every rank's rows of this cycle, in rank 0's (pinned) host memory
            last_view[0] = sol_e.gathered_view(tk, world, NQ)

    sol_e.set_cycle_depth(DEPTH)

    def e2e_loop(n):
        """n cycles, up to DEPTH in flight; ranks re-align every SLOTS cycles so that nobody laps the ring"""
        pending = []
        for i in range(n):
            if e2e_gather and i and i % SLOTS == 0:
                while pending:
                    finish(pending.pop(0))
                dist.barrier()
            if len(pending) == DEPTH:
                finish(pending.pop(0))
            pending.append(sol_e.cycle_submit(cc, h_pose, h_scan, None, h_tab, h_wp))
        while pending:
            finish(pending.pop(0))

    e2e_warm = max(warmup, 2 * DEPTH)   # every lane of the handle allocates its staging on first use: warm all of them up
    e2e_loop(e2e_warm + (SLOTS - e2e_warm % SLOTS) % SLOTS if e2e_gather else e2e_warm)   # (keeps the ring's cycle counter aligned to a slot 0)
    barrier()
    import gc
    gc.collect(); gc.disable()          # a collector pause inside a 0.3 ms host loop would be measured as GPU time
    t0 = time.perf_counter()
    e2e_cycles = max(args.steps, E2E_MIN_CYCLES)
    e2e_loop(e2e_cycles)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    gc.enable()
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = te.item()
    e2e_value = world * NQ * e2e_cycles / e2e_s
    # bytes that really move each step: poses + scans in; the output block (and, on rank 0, every rank's gathered rows) out.
    # The mini-path table and the raceline are start-up constants (project.cpp:34-37): uploaded once, not per step.
    h2d = int(h_pose.nbytes + h_scan.nbytes)
    d2h = int(sum(v.nbytes for v in out.values())) + (int(last_view[0].nbytes) if last_view[0] is not None else 0)
    e2e_launches = sol_e.last_launches
    assert (out["status"] == 1).mean() > 0.95
    e2e_gather_ok = None
    if last_view[0] is not None:
        hg = last_view[0]
        e2e_gather_ok = bool(np.array_equal(hg[0, :, :2], out["u0"]) and all((hg[r, :, 2] == 1).mean() > 0.95 for r in range(world)))
    if e2e_gather:
        barrier()
        sol_e.set_gather(None, 0, 0, 0, 0)

    if rank != 0:
        barrier()
        if gather:
            gather.close()
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- extras (rank 0): sustained rate, the same batch at the parity tolerance
    def timed_solves(solver, n, secs=None):
        """back-to-back solves of the 4096-QP batch (no L2 flush, no host sync in between); n launches or, with secs, until the
        device has been busy for that long"""
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        done, total_ms = 0, 0.0
        while True:
            a.record()
            for _ in range(n):
                solver.solve_device(d_recs, None, None, d_u0, d_status, d_iters, d_rhoup, None, stream=stream)
            b.record(); torch.cuda.synchronize()
            total_ms += a.elapsed_time(b); done += n
            if secs is None or total_ms >= secs * 1e3:
                return done, total_ms
    sus_sampler = ClockSampler(local)
    sus_sampler.start()
    n_sus, ms_sus = timed_solves(sol, 500, secs=2.0)
    sus_clocks = sus_sampler.stop()
    sustained = {"value": B * n_sus / (ms_sus * 1e-3), "unit": UNIT, "seconds": ms_sus * 1e-3, "launches": n_sus,
                 "ms_per_launch": ms_sus / n_sus, "sm_mhz_median": sus_clocks.get("sm_mhz"), "reasons": sus_clocks.get("reasons"),
                 "what": "the 4096-QP solve launched back to back for >= 2 s (no flush, no collision check): the thermal / power steady state"}
    sol4 = M.MpcSolver(M.default_config(N_HORIZON), M.default_settings(warm_start=0, eps_abs=1e-4, eps_rel=1e-4), max_batch=B, device=local)
    timed_solves(sol4, 3)
    n4, ms4 = timed_solves(sol4, 20)
    eps4 = {"value": B * n4 / (ms4 * 1e-3), "unit": UNIT, "ms_per_launch": ms4 / n4, "mean_iters": float(d_iters.float().mean().item()),
            "solved_fraction": float((d_status == 1).float().mean().item()),
            "what": "same batch at eps_abs = eps_rel = 1e-4 (the tolerance north_star states for parity), back to back"}
    timed_solves(sol, 1)   # leave the default-tolerance results in the output buffers

    # ---- single-QP latency (config 1 style): B = 1 host calls, warm started, sequential
    lat_sol = M.MpcSolver(M.default_config(N_HORIZON), M.default_settings(warm_start=1), max_batch=1, device=local)
    lat = []
    for i in range(300):
        r = wl["recs"][i * 13 % B: i * 13 % B + 1]
        t0 = time.perf_counter()
        lat_sol.solve_host(r, want_xy=False)
        lat.append((time.perf_counter() - t0) * 1e6)
    lat = np.array(lat[50:])

    # ---- roofline: FP64 issue + HBM for completeness
    peak_fp64 = M.fp64_fma_peak_tflops(local)
    achieved_tf = flops_launch / admm_s_per_launch / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_ach = B * hbm_bytes_per_qp(N_HORIZON) / admm_s_per_launch / 1e9
    traffic = None
    try:   # DRAM bytes per launch of this kernel from the committed ncu --set full capture (same workload size)
        tj = json.load(open(os.path.join(ROOT, "profiles", "admm_traffic.json")))
        if tj.get("qps_per_launch") == B:
            traffic = tj["traffic_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "fp64_issue", "kernel": "admm_kernel_tm", "achieved": achieved_tf, "peak": peak_fp64, "unit": "TFLOP/s",
                "frac": achieved_tf / peak_fp64, "traffic": traffic, "traffic_unit": "bytes/launch (ncu dram read+write)",
                "algorithmic_bytes_per_launch": B * hbm_bytes_per_qp(N_HORIZON),
                "peak_source": "DFMA micro-benchmark in this run (f110_bench_fp64_fma); MEASURED_PEAKS.json has no FP64 figure",
                "algorithmic_flops_per_launch": flops_launch, "launch_ms": admm_s_per_launch * 1e3,
                "launch_timing": "the solve launched alone after a 256 MiB L2 flush, CUDA events around the launch (the isolated leg of this run)",
                "frac_pipelined": flops_launch / (step_ms * 1e-3 / args.steps) / 1e12 / peak_fp64,
                "mean_iters": float(iters.mean()), "rho_updates_mean": float(rhoup.mean()),
                "hbm": {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}}

    # ---- CPU baseline beside it: bounded sample, all host threads
    threads, secs = cpu_reference_run(wl["recs"], steps=3, warmup=1)
    cpu_val = B * len(secs) / float(np.sum(secs))
    thr1, secs1 = cpu_reference_run(wl["recs"][:512], steps=1, warmup=0, nthreads=1)
    cpu_baseline = {"value": cpu_val, "unit": UNIT, "cores": threads, "kind": "port",
                    "sample": "3 x the 4096-QP batch on all host threads; single-thread rate from 512 QPs",
                    "per_core_value": 512 / float(np.sum(secs1)),
                    "threads": "std::thread per core, not pinned",
                    "note": "OSQP-algorithm restatement (oracle/), not the OSQP binary (unavailable offline)"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": step_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(wl),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "cycles": e2e_cycles,
                    "call": "f110_cycle_submit / f110_cycle_wait (qp_mode 2), %d cycles in flight: scans + poses in, %d kernels, controls out; "
                            "%d QPs per step per GPU%s" % (DEPTH, e2e_launches, NQ, "; every rank's rows gathered to rank 0's host buffer each step" if peer else ""),
                    "gathered_to_rank0_host": e2e_gather_ok},
            "gather": ({"how": "solve kernels store their packed rows into a CUDA-IPC-mapped ring on rank 0's GPU over NVLink; no collective on the step",
                        "check": gathered_check} if peer else
                       ({"how": "NCCL all-gather per step (IPC mapping unavailable: %s)" % (gather.why or "a rank failed to open the handle")} if world > 1 else None)),
            "gpu_launches": 2 * args.steps, "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks,
            "extra": {"isolated": isolated, "sustained": sustained, "eps1e-4": eps4},
            "latency_us": {"what": "B=1 f110_mpc_solve_host, warm start, sequential", "p50": float(np.percentile(lat, 50)),
                           "p90": float(np.percentile(lat, 90)), "p99": float(np.percentile(lat, 99))},
            "solved_fraction": float((status == 1).mean())}
    print(json.dumps(line))
    barrier()
    if gather:
        gather.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


# ---------------------------------------------------------------------------------------------------------------
# --configs: the five BASELINE.json configurations, each measured on the GPU with the CPU oracle beside it.
# Not the default bench line: writes a JSON report (profiles/configs_rNN.json).
def config4_records(W, n_sc=64):
    """BASELINE config 4 (7 lanes x 20 mini-paths x 64 scenarios = 8960 QPs): f110-mpc_b200/workloads.py::config4_records."""
    return W.config4_records(n_sc, N_HORIZON)


def main_config4(args):
    """Strong scaling of BASELINE config 4 over the ranks of a torchrun launch: 64 scenarios sharded by scenario,
    every rank solves its 140-QP scenarios, one all-gather of (u0, status, iters) per step.  Prints one JSON line (rank 0)."""
    import torch
    import torch.distributed as dist
    M = importlib.import_module("f110-mpc_b200")
    W = importlib.import_module("f110-mpc_b200.workloads")
    SH = importlib.import_module("f110-mpc_b200.sharding")
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        init_nccl(dist, torch, dev)
    M.build()
    recs = config4_records(W)
    total, per_sc = recs.shape[0], 7 * 20
    (s_lo, s_hi), (q_lo, q_hi) = SH.shard_by_scenario(64, per_sc, world, rank)
    mine = np.pad(recs[q_lo:q_hi], ((0, 0), (0, recs.shape[1] % 2)))
    b = mine.shape[0]
    sizes = [SH.shard_by_scenario(64, per_sc, world, r)[1] for r in range(world)]
    sizes = [hi - lo for lo, hi in sizes]
    sol = M.MpcSolver(M.default_config(N_HORIZON), M.default_settings(warm_start=0), max_batch=b, device=local)
    d = torch.from_numpy(np.ascontiguousarray(mine)).to(dev)
    streams = [torch.cuda.Stream(device=dev) for _ in range(PIPELINE_DEPTH)]
    outs = [(torch.empty(b, 2, dtype=torch.float64, device=dev), torch.empty(b, dtype=torch.int32, device=dev),
             torch.empty(b, dtype=torch.int32, device=dev)) for _ in range(PIPELINE_DEPTH)]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    main_stream = torch.cuda.current_stream()
    gathered = None
    # the gather: every rank's solve kernel stores its packed rows into a ring on rank 0's GPU (NVLink, CUDA IPC); NCCL only as fallback
    gather = PeerGather(M, dist, torch, dev, world, rank, local, max(sizes), 8) if world > 1 else None
    peer = gather is not None and gather.ok
    packed = torch.empty(b, 4, dtype=torch.float64, device=dev) if not peer else None
    cyc = [0]

    def step(i=0, ts=None):
        ts = ts or main_stream
        u0, st, it = outs[i % PIPELINE_DEPTH]
        if peer:
            rows, _ = gather.ring.slot(cyc[0])
            cyc[0] += 1
            M._check(M.lib().f110_mpc_set_packed_output(sol._h, rows), "f110_mpc_set_packed_output")
            sol.solve_device(d, None, None, u0, st, it, None, None, stream=ts.cuda_stream)
            return None
        sol.solve_device(d, None, None, u0, st, it, None, None, stream=ts.cuda_stream, packed=packed)
        with torch.cuda.stream(ts):
            return SH.gather_results(packed, world, max_rows=max(sizes), sizes=sizes)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for i in range(max(args.warmup, 3)):
        gathered = step(i, streams[i % PIPELINE_DEPTH])
    barrier()
    # `value`: K independent steps over PIPELINE_DEPTH streams (a shard of 1120 QPs at 8 GPUs is under one wave of warps, so one
    # step alone is bound by a single QP's latency; steps in flight fill the machine).  The NCCL fallback stays on one stream.
    pipe_ms = time_pipelined(torch, streams if (peer or world == 1) else [main_stream], args.steps, step)
    barrier()
    # isolated: one step at a time, 256 MiB L2 flush before each
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        ev[i][0].record(); gathered = step(0); ev[i][1].record()
    barrier()
    ms = torch.tensor([float(sum(a.elapsed_time(c) for a, c in ev)), pipe_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms, pipe_ms = ms.tolist()
    if rank == 0:
        if peer:   # the last step's slot: rank blocks in batch order, trimmed to each rank's shard
            blocks = ring_slot_to_host(torch, dev, gather, cyc[0] - 1, world, max(sizes))
            g = np.concatenate([blocks[r, :sizes[r]] for r in range(world)], axis=0)
        else:
            g = gathered.cpu().numpy()
        from oracle import oracle_py as O
        O.build()
        idx = np.arange(0, total, 35)           # a sample across every rank's shard, checked against the CPU oracle
        o = O.MpcBatch(O.default_cfg(N_HORIZON), O.default_settings(warm_start=0), len(idx)).solve(recs[idx])
        u0o = o["x"][:, 3 * (N_HORIZON + 1):3 * (N_HORIZON + 1) + 2]
        print(json.dumps({"metric": METRIC, "config": {"workload": "cfg4: 7 lanes x 20 mini-paths x 64 scenarios = 8960 N=30 QPs, sharded by scenario",
                                                        "qps_total": int(total), "qps_per_rank": sizes,
                                                        "l2_policy": "value: steps back to back over %d streams (3.3 MB of records per GPU, L2-resident); isolated: 256 MiB flush before each step" % PIPELINE_DEPTH},
                          "value": total * args.steps / (pipe_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                          "ms_per_step": pipe_ms / args.steps, "scaling": "strong", "higher_is_better": True, "dtype": "f64", "data": "synthetic",
                          "isolated": {"value": total * args.steps / (ms * 1e-3), "ms_per_step": ms / args.steps,
                                       "what": "one step at a time, 256 MiB L2 flush before each (the definition of `value` up to round 2a)"},
                          "gather": "solve kernels' NVLink stores into an IPC-mapped ring on rank 0" if peer else ("NCCL all-gather" if world > 1 else None),
                          "gathered_rows": int(g.shape[0]), "solved": int((g[:, 2] == 1).sum()),
                          "parity_sample": {"n": int(len(idx)), "status_equal": bool((g[idx, 2] == o["status"]).all()),
                                            "iters_equal": bool((g[idx, 3] == o["iters"]).all()),
                                            "max_abs_du0": float(np.abs(g[idx, :2] - u0o).max())}}))
    barrier()
    if gather:
        gather.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main_sweep(args):
    """BASELINE config 5 across the ranks of a torchrun launch: horizon sweep, 4096 QPs per GPU (weak scaling), one all-gather of
    (u0, status, iters) per step.  Prints one JSON line (rank 0) with a row per horizon."""
    import torch
    import torch.distributed as dist
    M = importlib.import_module("f110-mpc_b200")
    W = importlib.import_module("f110-mpc_b200.workloads")
    SH = importlib.import_module("f110-mpc_b200.sharding")
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        init_nccl(dist, torch, dev)
    M.build()
    B = 4096
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    main_stream = torch.cuda.current_stream()
    streams = [torch.cuda.Stream(device=dev) for _ in range(PIPELINE_DEPTH)]
    gather = PeerGather(M, dist, torch, dev, world, rank, local, B, 8) if world > 1 else None
    peer = gather is not None and gather.ok
    cyc = [0]
    rows = {}
    for N in (10, 20, 30, 50, 100):
        recs = W.tracking_batch(B, N, seed=20240905 + rank)
        recs = np.pad(recs, ((0, 0), (0, recs.shape[1] % 2)))
        sol = M.MpcSolver(M.default_config(N), M.default_settings(warm_start=0), max_batch=B, device=local)
        d = torch.from_numpy(np.ascontiguousarray(recs)).to(dev)
        packed = torch.empty(B, 4, dtype=torch.float64, device=dev)
        outs = [(torch.empty(B, 2, dtype=torch.float64, device=dev), torch.empty(B, dtype=torch.int32, device=dev),
                 torch.empty(B, dtype=torch.int32, device=dev)) for _ in range(PIPELINE_DEPTH)]

        def step(i=0, ts=None):
            ts = ts or main_stream
            u0, st, it = outs[i % PIPELINE_DEPTH]
            if peer:   # packed rows straight into the ring on rank 0's GPU
                rp, _ = gather.ring.slot(cyc[0])
                cyc[0] += 1
                M._check(M.lib().f110_mpc_set_packed_output(sol._h, rp), "f110_mpc_set_packed_output")
                sol.solve_device(d, None, None, u0, st, it, None, None, stream=ts.cuda_stream)
                return None
            sol.solve_device(d, None, None, u0, st, it, None, None, stream=ts.cuda_stream, packed=packed)
            with torch.cuda.stream(ts):
                return SH.gather_results(packed, world, max_rows=B, sizes=[B] * world)
        for _ in range(max(args.warmup, 3)):
            g = step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        # pipelined: batches back to back over PIPELINE_DEPTH streams — only where the kernel keeps its working state on chip
        # (horizons 16..127: launches of one handle then share nothing) and the gather needs no collective
        pipe_ms = None
        if 16 <= N <= 127 and (peer or world == 1):
            pipe_ms = time_pipelined(torch, streams, args.steps, step)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for i in range(args.steps):
            flush.fill_(i & 0xFF)
            ev[i][0].record(); g = step(); ev[i][1].record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([float(sum(a.elapsed_time(c) for a, c in ev)), pipe_ms or 0.0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        pipe_ms = ms[1].item() if pipe_ms is not None else None
        ms = ms[:1]
        if peer:
            gg = ring_slot_to_host(torch, dev, gather, cyc[0] - 1, world, B).reshape(world * B, 4) if rank == 0 else np.zeros((1, 4))
        else:
            gg = g.cpu().numpy()
        rows["N=%d" % N] = {"qps_total": world * B, "ms_per_step": ms.item() / args.steps, "solves_per_s": world * B * args.steps / (ms.item() * 1e-3),
                            "ms_per_step_pipelined": pipe_ms / args.steps if pipe_ms else None,
                            "solves_per_s_pipelined": world * B * args.steps / (pipe_ms * 1e-3) if pipe_ms else None,
                            "solved": int((gg[:, 2] == 1).sum()), "mean_iters": float(gg[:, 3].mean())}
    if rank == 0:
        print(json.dumps({"metric": METRIC, "unit": UNIT, "n_gpus": world, "steps": args.steps, "scaling": "weak", "dtype": "f64", "data": "synthetic",
                          "config": {"workload": "cfg5: horizon sweep, 4096 QPs per GPU, OSQP defaults, cold start",
                                     "l2_policy": "ms_per_step: 256 MiB flush before each step; *_pipelined: the batch back to back over %d streams" % PIPELINE_DEPTH},
                          "gather": "solve kernels' NVLink stores into an IPC-mapped ring on rank 0" if peer else ("NCCL all-gather" if world > 1 else None),
                          "horizons": rows}))
    if world > 1:
        dist.barrier()
    if gather:
        gather.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_configs(args):
    import torch
    M = importlib.import_module("f110-mpc_b200")
    W = importlib.import_module("f110-mpc_b200.workloads")
    from oracle import oracle_py as O
    M.build(); O.build()
    dev = torch.device("cuda:0")
    stream = torch.cuda.current_stream().cuda_stream
    amin, amax, inc = W.SCAN_ANGLE_MIN, W.SCAN_ANGLE_MAX, W.SCAN_ANGLE_INC
    report = {"gpu": torch.cuda.get_device_name(0), "host_threads": os.cpu_count(),
              "note": "CPU numbers: oracle/ (OSQP-algorithm restatement, not the OSQP binary). Parity UNPINNED by the reference."}

    pipe_streams = [torch.cuda.Stream(device=dev) for _ in range(PIPELINE_DEPTH)]

    def gpu_batch_time(N, recs, gap_mode=0, reps=10, rate_delta=None, **st):
        """One batch per launch: median of `reps` launches timed alone (ms, solves_per_s), and the same batch launched 4 x reps times
        over PIPELINE_DEPTH streams (pipelined_*: independent batches in flight, a batch's stragglers do not hold up the next)."""
        B = recs.shape[0]
        sol = M.MpcSolver(M.default_config(N, gap_mode, rate_delta), M.default_settings(warm_start=0, **st), max_batch=B)
        r = torch.from_numpy(np.ascontiguousarray(recs)).to(dev)
        u0 = torch.empty(B, 2, dtype=torch.float64, device=dev); stt = torch.empty(B, dtype=torch.int32, device=dev)
        it = torch.empty(B, dtype=torch.int32, device=dev); ru = torch.empty(B, dtype=torch.int32, device=dev)
        for _ in range(3):
            sol.solve_device(r, None, None, u0, stt, it, ru, None, stream=stream)
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); sol.solve_device(r, None, None, u0, stt, it, ru, None, stream=stream); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = float(np.median(ts))
        res = dict(ms=ms, solves_per_s=B / (ms * 1e-3), status=stt.cpu().numpy(), iters=it.cpu().numpy(), rho_updates=ru.cpu().numpy(),
                   u0=u0.cpu().numpy())
        # pipelined: only for kernels that keep their working state on chip (launches of one handle then share nothing)
        if rate_delta is None and 16 <= N <= 127:
            po = [(torch.empty(B, 2, dtype=torch.float64, device=dev), torch.empty(B, dtype=torch.int32, device=dev),
                   torch.empty(B, dtype=torch.int32, device=dev)) for _ in pipe_streams]
            n = 4 * reps
            main = torch.cuda.current_stream()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(main)
            for ps in pipe_streams:
                ps.wait_event(a)
            for i in range(n):
                j = i % len(pipe_streams)
                sol.solve_device(r, None, None, po[j][0], po[j][1], po[j][2], None, None, stream=pipe_streams[j].cuda_stream)
            for ps in pipe_streams:
                e = torch.cuda.Event(); e.record(ps); main.wait_event(e)
            b.record(main); torch.cuda.synchronize()
            res["pipelined_ms"] = a.elapsed_time(b) / n
            res["pipelined_solves_per_s"] = B / (res["pipelined_ms"] * 1e-3)
            res["pipelined_equal"] = bool((po[0][1].cpu().numpy() == res["status"]).all() and (po[0][2].cpu().numpy() == res["iters"]).all()
                                          and np.array_equal(po[0][0].cpu().numpy(), res["u0"], equal_nan=True))
        return res

    def pipe(g):
        return ({"gpu_ms_pipelined": g["pipelined_ms"], "gpu_solves_per_s_pipelined": g["pipelined_solves_per_s"],
                 "pipelined_results_equal": g["pipelined_equal"]} if "pipelined_ms" in g else {})

    def iter_hist(it):
        edges = [0, 25, 50, 75, 100, 200, 400, 1000, 2000, 3999, 4000]
        h = {}
        for lo, hi in zip(edges[:-1], edges[1:]):
            h["%d..%d" % (lo + 1, hi)] = int(((it > lo) & (it <= hi)).sum())
        return h

    def cpu_batch(N, recs, gap_mode=0, rate_delta=None, **st):
        mb = O.MpcBatch(O.default_cfg(N, gap_mode, rate_delta), O.default_settings(warm_start=0, **st), recs.shape[0])
        mb.solve(recs[: min(256, len(recs))], want_xy=False)
        r = mb.solve(recs, want_xy=True)
        return r, mb.threads

    def parity(g, o, N):
        ok = o["status"] > 0
        u0o = o["x"][:, 3 * (N + 1):3 * (N + 1) + 2]
        return {"status_equal": bool((g["status"] == o["status"]).all()), "iters_equal": bool((g["iters"] == o["iters"]).all()),
                "max_abs_du0": float(np.abs(g["u0"][ok] - u0o[ok]).max()) if ok.any() else None, "n": int(len(ok)), "n_solved": int(ok.sum())}

    # ---- config 1: single QP following skirk, full pipeline per cycle, sequential, warm start ------------------
    xy, ori = W.skirk_waypoints()
    n1 = 500 if not args.quick else 60
    free_scan = np.full(W.SCAN_BEAMS, 10.0, dtype=np.float32)
    mpc = M.HostMPC(N_HORIZON)
    mpc.update_scan(amin, amax, inc, free_scan)
    orc = O.MpcBatch(O.default_cfg(N_HORIZON), O.default_settings(warm_start=1), 1, 1)
    lat_g, lat_c, lat_plan, du0, steer_g, steer_c = [], [], [], [], 0.0, 0.0
    for i in range(n1):
        pose = W.yaw_pose(float(xy[i, 0]), float(xy[i, 1]), float(ori[i]))
        t0 = time.perf_counter()
        idx, path, valid, bg = M.host_plan(pose, amin, amax, inc, free_scan, xy)
        lat_plan.append((time.perf_counter() - t0) * 1e6)
        if idx < 0:
            continue
        state = np.array([pose[0], pose[1], float(ori[i])])
        t0 = time.perf_counter()
        g = mpc.update(state, [4.5, steer_g], path)
        lat_g.append((time.perf_counter() - t0) * 1e6)
        rec = np.concatenate([state, [4.5, steer_c], g["l1"], g["l2"], path[:N_HORIZON].reshape(-1)])[None, :]
        t0 = time.perf_counter()
        o = orc.solve(rec, warm=True)
        lat_c.append((time.perf_counter() - t0) * 1e6)
        uo = o["x"][0][3 * (N_HORIZON + 1):3 * (N_HORIZON + 1) + 2]
        if g["status"] == 1 and o["status"][0] == 1:
            du0.append(np.abs(g["inputs"][0] - uo).max())
            steer_g, steer_c = float(g["inputs"][0, 1]), float(uo[1])
    pct = lambda a: {"p50": float(np.percentile(a, 50)), "p90": float(np.percentile(a, 90)), "p99": float(np.percentile(a, 99))}
    report["config1_single_qp_skirk"] = {"cycles": len(lat_g), "gpu_MPC_Update_us": pct(lat_g[5:]), "cpu_oracle_solve_us": pct(lat_c[5:]),
                                         "plan_cycle_us (grid fill host + collision check GPU + selection)": pct(lat_plan[5:]),
                                         "max_abs_du0_vs_oracle": float(np.max(du0)) if du0 else None,
                                         "note": "GPU = C++ MPC::Update through f110_mpc_solve_host (B=1, H2D+kernel+D2H+sync); CPU = oracle update+solve only"}

    # ---- config 2: 20 mini-paths + grid check + QP per surviving path (and the 10 CSV paths) --------------------
    S2 = 256 if not args.quick else 32
    poses, yaws, scans = W.scene_batch(S2, seed=20240902)
    grids = np.zeros((S2, 10000), dtype=np.float32); offs = np.zeros((S2, 2), dtype=np.float32); rots = np.zeros((S2, 4))
    for s_ in range(S2):
        grids[s_], offs[s_] = M.host_fill_grid(poses[s_], amin, amax, inc, scans[s_])
        rots[s_] = M.host_car_to_world_R(poses[s_])
    c2 = {}
    for name, table in (("steer19_P20", np.ascontiguousarray(W.traj_table(steer_discrete=19)[:, :, :2])),
                        ("csv10 (local_traj_50.csv, axes swapped to x-forward)", np.ascontiguousarray(W.reference_data()["local_traj10_xy"]))):
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        dg, do, dr, dp, dt_ = t(grids), t(offs), t(rots), t(poses[:, :2].copy()), t(table)
        P = table.shape[0]
        dv = torch.empty(S2, P, dtype=torch.uint8, device=dev); df = torch.empty(S2, P, dtype=torch.int32, device=dev)
        de = torch.empty(S2, P, 2, dtype=torch.float32, device=dev)
        for _ in range(3):
            M.collision_check_device(dg, do, dr, dp, dt_, dv, df, de, stream=stream)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            M.collision_check_device(dg, do, dr, dp, dt_, dv, df, de, stream=stream)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        valid = dv.cpu().numpy(); endw = de.cpu().numpy()
        mism = 0
        t0 = time.perf_counter()
        for s_ in range(S2):
            v, f, e = O.collision_check(grids[s_], 100, 0.1, offs[s_], rots[s_], poses[s_, :2], table)
            mism += int((v != valid[s_]).sum()) + int((e.view(np.uint32) != endw[s_].view(np.uint32)).sum())
        cpu_ms = (time.perf_counter() - t0) * 1e3
        c2[name] = {"scenes": S2, "paths": P, "valid_paths": int(valid.sum()), "bit_mismatches_vs_oracle": mism, "gpu_check_ms": ms,
                    "cpu_oracle_check_ms_1thread": cpu_ms, "algorithmic_bytes": int(S2 * P * 50 * 20 + S2 * P * 13),
                    "achieved_GBps": (S2 * P * 50 * 20 + S2 * P * 13) / (ms * 1e-3) / 1e9}
        if name.startswith("steer19"):
            recs = []
            for s_ in range(S2):
                for pidx in np.nonzero(valid[s_])[0]:
                    ref = np.zeros((N_HORIZON, 3)); ref[:, :2] = W.path_to_world(table[pidx, :N_HORIZON], poses[s_, 0], poses[s_, 1], yaws[s_])
                    recs.append(np.concatenate([[poses[s_, 0], poses[s_, 1], yaws[s_]], [4.5, 0.0], [0.3, -0.8, 1.5], [-0.4, 0.7, 2.0], ref.reshape(-1)]))
            recs = np.array(recs)
            g = gpu_batch_time(N_HORIZON, recs)
            o, thr = cpu_batch(N_HORIZON, recs)
            c2[name]["qp_per_surviving_path"] = {"qps": len(recs), "gpu_ms": g["ms"], "gpu_solves_per_s": g["solves_per_s"], **pipe(g),
                                                 "cpu_solves_per_s": len(recs) / o["seconds"], "cpu_threads": thr, "parity": parity(g, o, N_HORIZON)}
    report["config2_minipaths_grid_check"] = c2

    # ---- config 3: laser-gap half-plane constrained, B = 1024 ------------------------------------------------------
    B3 = 1024 if not args.quick else 128
    rng = np.random.default_rng(20240903)
    recs3 = W.tracking_batch(B3, N_HORIZON, seed=20240903)
    n_gap = 0
    for b_ in range(B3):
        r = rng.uniform(0.5, 2.8, W.SCAN_BEAMS).astype(np.float32)
        for _ in range(rng.integers(1, 4)):
            a = rng.integers(150, 880); w = rng.integers(8, 201)
            r[a:a + w] = rng.uniform(3.5, 10.0)
        ok, l1, l2, _ = M.host_find_half_spaces(recs3[b_, :3], amin, amax, inc, r)
        if ok:
            recs3[b_, 5:8] = l1; recs3[b_, 8:11] = l2; n_gap += 1
    c3 = {"B": B3, "scans_with_gap": n_gap}
    for mode, label in ((0, "as_shipped (gap bounds +-1e30)"), (1, "gap_enabled (lower = -l(2), every stage incl. the all-ones stage-0 pair)"),
                        (2, "gap_enabled_k>=1 (stage-0 pair loose)")):
        g = gpu_batch_time(N_HORIZON, recs3, gap_mode=mode)
        o, thr = cpu_batch(N_HORIZON, recs3, gap_mode=mode)
        c3[label] = {"gpu_ms": g["ms"], "gpu_solves_per_s": g["solves_per_s"], **pipe(g), "cpu_solves_per_s": B3 / o["seconds"], "cpu_threads": thr,
                     "parity": parity(g, o, N_HORIZON), "iters_hist": iter_hist(g["iters"]), "max_iters": int(g["iters"].max()), "status_hist": {str(k): int(v) for k, v in zip(*np.unique(g["status"], return_counts=True))}}
    report["config3_gap_constrained"] = c3

    # ---- config 4: 7 lanes x 20 paths x 64 scenarios = 8960 QPs (1 GPU here; bench.py --gpus N shards) ------------
    recs4 = config4_records(W, 64 if not args.quick else 8)
    g = gpu_batch_time(N_HORIZON, recs4)
    o, thr = cpu_batch(N_HORIZON, recs4)
    report["config4_7lanes_20paths_64scenarios"] = {"qps": len(recs4), "gpu_ms": g["ms"], "gpu_solves_per_s": g["solves_per_s"], **pipe(g),
                                                    "cpu_solves_per_s": len(recs4) / o["seconds"], "cpu_threads": thr, "parity": parity(g, o, N_HORIZON)}

    # ---- config 5: horizon sweep, 4096 QPs per GPU ----------------------------------------------------------------------
    peak = M.fp64_fma_peak_tflops(0)
    c5 = {"fp64_fma_peak_tflops_measured": peak}
    B5 = 4096 if not args.quick else 512
    for N in (10, 20, 30, 50, 100):
        recs5 = W.tracking_batch(B5, N, seed=20240905)
        g = gpu_batch_time(N, recs5)
        o, thr = cpu_batch(N, recs5)
        fl = float(flops_per_qp(N, g["iters"], g["rho_updates"]).sum())
        c5["N=%d" % N] = {"qps": B5, "gpu_ms": g["ms"], "gpu_solves_per_s": g["solves_per_s"], **pipe(g), "cpu_solves_per_s": B5 / o["seconds"],
                          "cpu_threads": thr, "speedup_vs_host": g["solves_per_s"] / (B5 / o["seconds"]), "mean_iters": float(g["iters"].mean()),
                          "algorithmic_tflops": fl / (g["ms"] * 1e-3) / 1e12, "fp64_roofline_frac": fl / (g["ms"] * 1e-3) / 1e12 / peak,
                          "parity": parity(g, o, N)}
    report["config5_horizon_sweep"] = c5

    # ---- steering-rate rows (not in the reference): 4x4-block variant of the kernel --------------------------
    rd = 3.2 * float(np.float32(0.01))      # 3.2 rad/s servo limit (f1tenth simulator's max_steering_vel) x dt
    rr = {"rate_delta_rad_per_step": rd}
    for N in (30, 50):
        recs6 = W.tracking_batch(B5, N, seed=20240907)
        g = gpu_batch_time(N, recs6, rate_delta=rd)
        o, thr = cpu_batch(N, recs6, rate_delta=rd)
        rr["N=%d" % N] = {"qps": B5, "gpu_ms": g["ms"], "gpu_solves_per_s": g["solves_per_s"], "cpu_solves_per_s": B5 / o["seconds"],
                          "cpu_threads": thr, "speedup_vs_host": g["solves_per_s"] / (B5 / o["seconds"]), "mean_iters": float(g["iters"].mean()),
                          "parity": parity(g, o, N)}
    report["steering_rate_rows"] = rr

    # ---- device-resident cycle: scan + pose in, control out, one QP per car ----------------------------
    Sc = 4096 if not args.quick else 256
    poses, yaws, scans = W.scene_batch(Sc, seed=20240906)
    table = np.ascontiguousarray(W.traj_table(steer_discrete=19)[:, :, :2])
    sol = M.MpcSolver(M.default_config(N_HORIZON), M.default_settings(warm_start=0), max_batch=Sc)
    cc = M.default_cycle_config()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    h_pose, h_scan = torch.from_numpy(poses).pin_memory(), torch.from_numpy(scans).pin_memory()
    d_pose, d_scan, d_tab, d_wp = t(poses), t(scans), t(table), t(xy)
    u0 = torch.empty(Sc, 2, dtype=torch.float64, device=dev); stt = torch.empty(Sc, dtype=torch.int32, device=dev)
    it = torch.empty(Sc, dtype=torch.int32, device=dev); ch = torch.empty(Sc, dtype=torch.int32, device=dev)
    h_u0 = torch.empty(Sc, 2, dtype=torch.float64).pin_memory()
    for _ in range(3):
        sol.cycle_device(cc, d_pose, d_scan, None, d_tab, d_wp, u0, stt, it, ch, stream=stream)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        sol.cycle_device(cc, d_pose, d_scan, None, d_tab, d_wp, u0, stt, it, ch, stream=stream)
    b.record(); torch.cuda.synchronize()
    ms_dev = a.elapsed_time(b) / 10
    t0 = time.perf_counter()
    for _ in range(10):
        d_pose.copy_(h_pose, non_blocking=True); d_scan.copy_(h_scan, non_blocking=True)
        sol.cycle_device(cc, d_pose, d_scan, None, d_tab, d_wp, u0, stt, it, ch, stream=stream)
        h_u0.copy_(u0, non_blocking=True)
        torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / 10
    n_solved = int((stt == 1).sum().item())
    report["device_cycle"] = {"scenes": Sc, "kernels_per_cycle": sol.last_launches, "solved": n_solved, "device_ms": ms_dev,
                              "cars_per_s_device": Sc / (ms_dev * 1e-3), "e2e_ms (scan+pose H2D, u0 D2H)": ms_e2e,
                              "cars_per_s_e2e": Sc / (ms_e2e * 1e-3), "h2d_bytes": int(poses.nbytes + scans.nbytes)}
    # ---- batched closed loop: cars x ticks of the reference's plan / control / drive state machine on the device,
    # warm-started MPC cycle per car and tick, no host round trip inside the run ------------------------------------------------
    fl = {}
    for cars in ((12, 4096) if not args.quick else (12, 256)):
        fsol = M.MpcSolver(M.default_config(N_HORIZON), M.default_settings(warm_start=1), max_batch=cars)
        fleet = M.Fleet(fsol, M.default_cycle_config(qp_mode=0), cars, table, xy, drive_every=2, scan_every=4, dt_tick=0.01)
        fp, fy, fs = W.scene_batch(cars, seed=20240908)
        pose3 = np.stack([fp[:, 0], fp[:, 1], fy], axis=1)
        ticks = 200 if not args.quick else 40
        fleet.reset(pose3, fs)
        fleet.run(20, log=False)
        fleet.reset(pose3, fs)
        t0 = time.perf_counter()
        li, ld = fleet.run(ticks, log=True)
        dt_run = time.perf_counter() - t0
        ctrl = li[:, :, 0] == M.Fleet.CONTROL
        fl["cars=%d" % cars] = {"ticks": ticks, "seconds": dt_run, "us_per_tick": dt_run / ticks * 1e6, "car_ticks_per_s": cars * ticks / dt_run,
                                "mpc_cycles": int(ctrl.sum()), "mpc_cycles_per_s": float(ctrl.sum() / dt_run),
                                "solved_fraction_of_mpc_cycles": float((li[:, :, 2][ctrl] == 1).mean()) if ctrl.any() else None,
                                "mean_iters_of_mpc_cycles": float(li[:, :, 3][ctrl].mean()) if ctrl.any() else None,
                                "launches_per_tick": fsol.last_launches / ticks if fsol.last_launches else None,
                                "note": "host clock around f110_fleet_run incl. the device-to-host copy of the per-tick log"}
        fleet.close()
    report["fleet_closed_loop"] = fl
    out = args.configs_out or os.path.join(ROOT, "gpurun_out", "configs.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    json.dump(report, open(out, "w"), indent=1)
    print(json.dumps({"configs_report": out}))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--skip-extras", action="store_true",
                    help="only the device-timed region (no e2e / latency / CPU-baseline legs): the command ncu wraps")
    ap.add_argument("--configs", action="store_true", help="measure the five BASELINE.json configs (report file, not the bench line)")
    ap.add_argument("--configs-out", default=None)
    ap.add_argument("--quick", action="store_true", help="smaller --configs run")
    ap.add_argument("--sweep", action="store_true", help="BASELINE config 5 (horizon sweep, 4096 QPs per GPU) over the ranks of a torchrun launch")
    ap.add_argument("--config4", action="store_true", help="strong scaling of BASELINE config 4 (8960 QPs) over the ranks of a torchrun launch")
    args = ap.parse_args()
    if args.configs:
        return run_configs(args)
    if args.config4:
        return main_config4(args)
    if args.sweep:
        return main_sweep(args)
    if args.impl == "reference":
        return main_reference(args)
    return main_product(args)


if __name__ == "__main__":
    sys.exit(main())
