#!/usr/bin/env python
"""bench.py — batched per-cycle MPC solve throughput on B200 (metric of BASELINE.json).

One "step" = one pass of the hot path over one batch: the mini-path collision check of every candidate
path of every scene (1 kernel) + the ADMM solve of one tracking QP per candidate path (1 kernel), 4096 QPs
per GPU, N = 30 (params.yaml horizon), OSQP default settings (eps 1e-3 — what the reference runs,
mpc.cpp:98-99), cold start.  Weak scaling: every rank owns its own 4096 QPs; the only collective is the
final gather of (u0, status, iters) to every rank.

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


def build_workload(M, W, rank):
    """4096 QPs = 205 scenes x 20 mini-paths (truncated): ego on a skirk waypoint with a small tracking error,
    reference = mini-path p in the world frame; plus the scenes' occupancy grids for the collision check."""
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
                table_xy=np.ascontiguousarray(table[:, :, :2]), scenes=S)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


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
    M = importlib.import_module("f110-mpc_b200")
    W = importlib.import_module("f110-mpc_b200.workloads")
    M.build()
    wl = build_workload(M, W, 0)
    threads, secs = cpu_reference_run(wl["recs"], args.steps, args.warmup)
    tot = float(np.sum(secs))
    value = QPS_PER_GPU * args.steps / tot
    sample = "%d x the %d-QP batch (all QPs of the workload), one QP per thread at a time" % (args.steps, QPS_PER_GPU)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(wl),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                             "note": "OSQP-algorithm restatement (oracle/), not the OSQP binary (unavailable offline)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def workload_config(wl):
    return {"workload": "cfg2x4096: %d skirk scenes x %d mini-paths (first %d), grid collision check of every path + one "
                        "N=%d tracking QP per path, OSQP defaults (eps 1e-3), cold start" % (wl["scenes"], PATHS, QPS_PER_GPU, N_HORIZON),
            "qps_per_gpu": QPS_PER_GPU, "horizon": N_HORIZON, "paths": PATHS, "samples": SAMPLES, "scenes_per_gpu": wl["scenes"],
            "eps_abs": 1e-3, "eps_rel": 1e-3, "l2_policy": "256 MiB buffer written between timed steps (inputs are 3.3 MB)",
            "parallelism": "independent QPs sharded by rank, final all-gather of (u0,status,iters)"}


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
        dist.init_process_group("nccl", device_id=dev)
    M.build()
    wl = build_workload(M, W, rank)
    B, S = QPS_PER_GPU, wl["scenes"]
    sol = M.MpcSolver(M.default_config(N_HORIZON), M.default_settings(warm_start=0), max_batch=B, device=local)
    # ---- device-resident inputs / outputs
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_recs, d_grid, d_off, d_rot, d_pose, d_tab = (t(wl[k]) for k in ("recs", "grids", "offs", "rots", "pose_xy", "table_xy"))
    d_u0 = torch.empty(B, 2, dtype=torch.float64, device=dev)
    d_status = torch.empty(B, dtype=torch.int32, device=dev)
    d_iters = torch.empty(B, dtype=torch.int32, device=dev)
    d_rhoup = torch.empty(B, dtype=torch.int32, device=dev)
    d_valid = torch.empty(S, PATHS, dtype=torch.uint8, device=dev)
    d_free = torch.empty(S, PATHS, dtype=torch.int32, device=dev)
    d_endw = torch.empty(S, PATHS, 2, dtype=torch.float32, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        M.collision_check_device(d_grid, d_off, d_rot, d_pose, d_tab, d_valid, d_free, d_endw, stream=stream)
        sol.solve_device(d_recs, None, None, d_u0, d_status, d_iters, d_rhoup, None, stream=stream)
        if world > 1:   # the one collective: chosen controls of every rank, in batch order
            SH.gather_results(SH.pack_result(d_u0, d_status, d_iters), world, max_rows=B, sizes=[B] * world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- timed region: K steps, CUDA events on the launching stream around each step, L2 flushed in between
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        ev[i][0].record()
        M.collision_check_device(d_grid, d_off, d_rot, d_pose, d_tab, d_valid, d_free, d_endw, stream=stream)
        kev[i][0].record()
        sol.solve_device(d_recs, None, None, d_u0, d_status, d_iters, d_rhoup, None, stream=stream)
        kev[i][1].record()
        if world > 1:   # the one collective: chosen controls of every rank, in batch order
            SH.gather_results(SH.pack_result(d_u0, d_status, d_iters), world, max_rows=B, sizes=[B] * world)
        ev[i][1].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    step_ms = float(sum(a.elapsed_time(b) for a, b in ev))
    admm_ms = float(sum(a.elapsed_time(b) for a, b in kev))
    tt = torch.tensor([step_ms, admm_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    step_ms, admm_ms = tt.tolist()
    value = world * B * args.steps / (step_ms * 1e-3)

    iters = d_iters.cpu().numpy()
    rhoup = d_rhoup.cpu().numpy()
    status = d_status.cpu().numpy()
    flops_launch = float(flops_per_qp(N_HORIZON, iters, rhoup).sum())
    admm_s_per_launch = admm_ms * 1e-3 / args.steps

    if args.skip_extras:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                              "ms_per_step": step_ms / args.steps, "admm_ms_per_launch": admm_ms / args.steps,
                              "note": "--skip-extras run (profiling target), not a bench value"}))
        if world > 1:
            dist.destroy_process_group()
        return 0
    # ---- e2e: the reference-facing host-buffer calls, pinned host memory, copies inside the timed region
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    h_recs = pin(wl["recs"])
    h_grid, h_off, h_rot, h_pose, h_tab = (pin(wl[k]) for k in ("grids", "offs", "rots", "pose_xy", "table_xy"))
    out = {"u0": pin(np.empty((B, 2))), "status": pin(np.empty(B, dtype=np.int32)), "iters": pin(np.empty(B, dtype=np.int32))}
    for _ in range(3):
        M.collision_check_host(h_grid, h_off, h_rot, h_pose, h_tab, device=local)
        sol.solve_host(h_recs, want_xy=False, out=out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        M.collision_check_host(h_grid, h_off, h_rot, h_pose, h_tab, device=local)
        sol.solve_host(h_recs, want_xy=False, out=out)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = te.item()
    h2d = int(h_recs.nbytes + h_grid.nbytes + h_off.nbytes + h_rot.nbytes + h_pose.nbytes + h_tab.nbytes)
    d2h = int(out["u0"].nbytes + out["status"].nbytes + out["iters"].nbytes + S * PATHS * (1 + 4 + 8))
    assert np.array_equal(out["iters"], iters) and np.array_equal(out["status"], status)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- single-QP latency (config 1 style): B = 1 host calls, warm started, sequential
    lat_sol = M.MpcSolver(M.default_config(N_HORIZON), M.default_settings(warm_start=1), max_batch=1, device=local)
    lat = []
    for i in range(300):
        r = h_recs[i * 13 % B: i * 13 % B + 1]
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
    roofline = {"bound": "fp64_issue", "kernel": "admm_kernel", "achieved": achieved_tf, "peak": peak_fp64, "unit": "TFLOP/s",
                "frac": achieved_tf / peak_fp64, "traffic": None,
                "peak_source": "DFMA micro-benchmark in this run (f110_bench_fp64_fma); MEASURED_PEAKS.json has no FP64 figure",
                "algorithmic_flops_per_launch": flops_launch, "launch_ms": admm_s_per_launch * 1e3,
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
                    "note": "OSQP-algorithm restatement (oracle/), not the OSQP binary (unavailable offline)"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": step_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(wl),
            "e2e": {"value": world * B * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": 2 * args.steps, "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks,
            "latency_us": {"what": "B=1 f110_mpc_solve_host, warm start, sequential", "p50": float(np.percentile(lat, 50)),
                           "p90": float(np.percentile(lat, 90)), "p99": float(np.percentile(lat, 99))},
            "solved_fraction": float((status == 1).mean())}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--skip-extras", action="store_true",
                    help="only the device-timed region (no e2e / latency / CPU-baseline legs): the command ncu wraps")
    args = ap.parse_args()
    if args.impl == "reference":
        return main_reference(args)
    return main_product(args)


if __name__ == "__main__":
    sys.exit(main())
