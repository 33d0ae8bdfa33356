"""ORACLE — test infrastructure only. ctypes loader for oracle/_ref/liboracle.so (PARITY UNPINNED).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "liboracle.so")

SOLVED, SOLVED_INACCURATE, MAX_ITER_REACHED = 1, 2, -2
PRIMAL_INFEASIBLE, DUAL_INFEASIBLE = -3, -4


def build(force=False):
    """Compile the restatement (gcc only; never reads /root/reference)."""
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "f110_ref.hpp", "osqp_restated.hpp", "Makefile")]
    stale = force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        dp, ip, fp, u8p = (C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_uint8))
        L.orc_qp_record_doubles.restype = C.c_int
        L.orc_mpc_create.restype = C.c_void_p
        L.orc_mpc_create.argtypes = [dp, dp, C.c_int, C.c_int]
        L.orc_mpc_destroy.argtypes = [C.c_void_p]
        L.orc_mpc_threads.argtypes = [C.c_void_p]
        L.orc_mpc_solve.restype = C.c_double
        L.orc_mpc_solve.argtypes = [C.c_void_p, dp, C.c_int, C.c_int, C.c_int, dp, dp, ip, ip, ip, dp]
        L.orc_mpc_get_scaling.argtypes = [C.c_void_p, C.c_int, dp, dp, dp]
        L.orc_osqp_dense.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, dp, dp, dp]
        L.orc_mpc_assemble_dense.argtypes = [dp] * 7
        L.orc_mpc_nnz.argtypes = [dp, ip, ip, ip]
        L.orc_linearize.argtypes = [C.c_double] * 4 + [dp] * 3
        L.orc_simulate_dynamics.argtypes = [dp, dp, C.c_double, dp]
        L.orc_traj_table.argtypes = [C.c_double, C.c_int, C.c_int, C.c_double, C.c_double, dp]
        L.orc_car_to_world_R.argtypes = [dp, dp]
        L.orc_car_orientation.restype = C.c_float
        L.orc_car_orientation.argtypes = [dp]
        L.orc_fill_grid.argtypes = [C.c_int, C.c_float, C.c_float, dp, C.c_float, C.c_float, C.c_float, fp, C.c_int, fp, fp]
        L.orc_grid_blocks.argtypes = [C.c_int, C.c_float]
        L.orc_collision_check.argtypes = [fp, C.c_int, C.c_float, fp, dp, dp, dp, C.c_int, C.c_int, u8p, ip, fp]
        L.orc_select_best.argtypes = [u8p, fp, C.c_int, C.c_double, C.c_double]
        L.orc_waypoint_headings.argtypes = [fp, C.c_int, dp]
        L.orc_best_global_idx.argtypes = [fp, C.c_int, dp, C.c_float]
        L.orc_find_half_spaces.argtypes = [C.c_float] * 3 + [dp] + [C.c_float] * 3 + [fp, C.c_int, dp, dp, ip]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int)) if a is not None else None


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


def default_cfg(N=30, gap_mode=0, rate_delta=None, state_lim=None):
    """rate_delta: max steering change per step (rad) -> N steering-rate rows appended; state_lim: d of Constraints::SetXLims ->
    3(N+1) state-box rows appended; None = the reference's row set."""
    c = np.zeros(18)
    lib().orc_default_cfg(_dp(c))
    c[0] = N
    c[13] = gap_mode
    if rate_delta is not None:
        c[14], c[15] = 1, rate_delta
    if state_lim is not None:
        c[16], c[17] = 1, state_lim
    return c


def mpc_rows(cfg):
    N = int(cfg[0])
    return 7 * N + 5 + (N if len(cfg) > 14 and cfg[14] else 0) + (3 * (N + 1) if len(cfg) > 16 and cfg[16] else 0)


def default_settings(**kw):
    s = np.zeros(15)
    lib().orc_default_settings(_dp(s))
    names = ["rho", "sigma", "alpha", "eps_abs", "eps_rel", "eps_prim_inf", "eps_dual_inf", "max_iter",
             "check_termination", "scaling", "adaptive_rho", "adaptive_rho_interval", "adaptive_rho_tolerance",
             "warm_start", "scaled_termination"]
    for k, v in kw.items():
        s[names.index(k)] = v
    return s


def record_doubles(N):
    return 11 + 3 * N


class MpcBatch:
    """B QP slots solved on host threads (one QP per thread at a time)."""

    def __init__(self, cfg, settings, B, nthreads=0):
        self.cfg = np.ascontiguousarray(cfg, dtype=np.float64)
        self.settings = np.ascontiguousarray(settings, dtype=np.float64)
        self.N = int(cfg[0])
        self.n, self.m = 5 * self.N + 3, mpc_rows(self.cfg)
        self.B = B
        self.h = lib().orc_mpc_create(_dp(self.cfg), _dp(self.settings), B, nthreads)
        self.threads = lib().orc_mpc_threads(self.h)

    def solve(self, recs, warm=False, want_xy=True):
        recs = np.ascontiguousarray(recs, dtype=np.float64)
        B = recs.shape[0]
        x = np.zeros((B, self.n)) if want_xy else None
        y = np.zeros((B, self.m)) if want_xy else None
        status = np.zeros(B, dtype=np.int32)
        iters = np.zeros(B, dtype=np.int32)
        rho_updates = np.zeros(B, dtype=np.int32)
        extra = np.zeros((B, 4))
        secs = lib().orc_mpc_solve(self.h, _dp(recs), recs.shape[1], B, int(warm), _dp(x), _dp(y), _ip(status),
                                   _ip(iters), _ip(rho_updates), _dp(extra))
        if secs < 0:
            raise RuntimeError("oracle solve failed: %r" % secs)
        return dict(x=x, y=y, status=status, iters=iters, rho_updates=rho_updates, obj=extra[:, 0],
                    pri_res=extra[:, 1], dua_res=extra[:, 2], rho=extra[:, 3], seconds=secs)

    def scaling(self, slot):
        D, E, c = np.zeros(self.n), np.zeros(self.m), np.zeros(1)
        if lib().orc_mpc_get_scaling(self.h, slot, _dp(D), _dp(E), _dp(c)):
            raise RuntimeError("slot not set up (warm mode only)")
        return D, E, float(c[0])

    def close(self):
        if self.h:
            lib().orc_mpc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def osqp_dense(P, q, A, l, u, settings=None):
    P = np.ascontiguousarray(P, dtype=np.float64)
    A = np.ascontiguousarray(A, dtype=np.float64)
    q, l, u = (np.ascontiguousarray(v, dtype=np.float64) for v in (q, l, u))
    n, m = P.shape[0], A.shape[0]
    s = default_settings() if settings is None else np.ascontiguousarray(settings, dtype=np.float64)
    x, y, info = np.zeros(n), np.zeros(m), np.zeros(8)
    rc = lib().orc_osqp_dense(n, m, _dp(P), _dp(q), _dp(A), _dp(l), _dp(u), _dp(s), _dp(x), _dp(y), _dp(info))
    if rc:
        raise RuntimeError("orc_osqp_dense rc=%d" % rc)
    return dict(x=x, y=y, iter=int(info[0]), status=int(info[1]), obj=info[2], pri_res=info[3], dua_res=info[4],
                rho_estimate=info[5], rho_updates=int(info[6]), n_factor=int(info[7]))


def mpc_assemble_dense(cfg, rec):
    N = int(cfg[0])
    n, m = 5 * N + 3, mpc_rows(cfg)
    P, q, A, l, u = np.zeros((n, n)), np.zeros(n), np.zeros((m, n)), np.zeros(m), np.zeros(m)
    cfg = np.ascontiguousarray(cfg, dtype=np.float64)
    rec = np.ascontiguousarray(rec, dtype=np.float64)
    lib().orc_mpc_assemble_dense(_dp(cfg), _dp(rec), _dp(P), _dp(q), _dp(A), _dp(l), _dp(u))
    return P, q, A, l, u


def mpc_nnz(cfg):
    a, b, c = (np.zeros(1, dtype=np.int32) for _ in range(3))
    cfg = np.ascontiguousarray(cfg, dtype=np.float64)
    lib().orc_mpc_nnz(_dp(cfg), _ip(a), _ip(b), _ip(c))
    return int(a[0]), int(b[0]), int(c[0])


def linearize(ori, v, steer, dt):
    A, B, Cc = np.zeros(9), np.zeros(6), np.zeros(3)
    lib().orc_linearize(ori, v, steer, dt, _dp(A), _dp(B), _dp(Cc))
    return A.reshape(3, 3), B.reshape(3, 2), Cc


def traj_table(steer_max=0.4, steer_discrete=30, traj_discrete=50, speed_max=4.5, dt=0.01):
    out = np.zeros((steer_discrete + 1, traj_discrete, 3))
    P = lib().orc_traj_table(steer_max, steer_discrete, traj_discrete, speed_max, dt, _dp(out))
    assert P == steer_discrete + 1
    return out


def car_to_world_R(pose7):
    pose7 = np.ascontiguousarray(pose7, dtype=np.float64)
    R = np.zeros(4)
    lib().orc_car_to_world_R(_dp(pose7), _dp(R))
    return R


def car_orientation(pose7):
    pose7 = np.ascontiguousarray(pose7, dtype=np.float64)
    return float(lib().orc_car_orientation(_dp(pose7)))


def fill_grid(pose7, angle_min, angle_max, angle_inc, ranges, occ_size=10, discrete=0.1, dilation=0.15):
    pose7 = np.ascontiguousarray(pose7, dtype=np.float64)
    ranges = np.ascontiguousarray(ranges, dtype=np.float32)
    blocks = lib().orc_grid_blocks(occ_size, discrete)
    grid = np.zeros(blocks * blocks, dtype=np.float32)
    off = np.zeros(2, dtype=np.float32)
    lib().orc_fill_grid(occ_size, discrete, dilation, _dp(pose7), angle_min, angle_max, angle_inc, _fp(ranges),
                        len(ranges), _fp(grid), _fp(off))
    return grid, off, blocks


def collision_check(grid, blocks, discrete, offset, R4, pose_xy, table_xy):
    grid = np.ascontiguousarray(grid, dtype=np.float32)
    offset = np.ascontiguousarray(offset, dtype=np.float32)
    R4 = np.ascontiguousarray(R4, dtype=np.float64)
    pose_xy = np.ascontiguousarray(pose_xy, dtype=np.float64)
    table_xy = np.ascontiguousarray(table_xy, dtype=np.float64)
    P, S = table_xy.shape[0], table_xy.shape[1]
    valid = np.zeros(P, dtype=np.uint8)
    free = np.zeros(P, dtype=np.int32)
    endw = np.zeros((P, 2), dtype=np.float32)
    lib().orc_collision_check(_fp(grid), blocks, discrete, _fp(offset), _dp(R4), _dp(pose_xy), _dp(table_xy), P, S,
                              valid.ctypes.data_as(C.POINTER(C.c_uint8)), _ip(free), _fp(endw))
    return valid, free, endw


def select_best(valid, end_world, gx, gy):
    valid = np.ascontiguousarray(valid, dtype=np.uint8)
    end_world = np.ascontiguousarray(end_world, dtype=np.float32)
    return lib().orc_select_best(valid.ctypes.data_as(C.POINTER(C.c_uint8)), _fp(end_world), len(valid), gx, gy)


def waypoint_headings(wp_xy):
    wp_xy = np.ascontiguousarray(wp_xy, dtype=np.float32)
    out = np.zeros(len(wp_xy))
    lib().orc_waypoint_headings(_fp(wp_xy), len(wp_xy), _dp(out))
    return out


def best_global_idx(wp_xy, pose7, lookahead=2.5):
    wp_xy = np.ascontiguousarray(wp_xy, dtype=np.float32)
    pose7 = np.ascontiguousarray(pose7, dtype=np.float64)
    return lib().orc_best_global_idx(_fp(wp_xy), len(wp_xy), _dp(pose7), lookahead)


def find_half_spaces(state3, angle_min, angle_max, angle_inc, ranges, thresh=3.0, divider=1.5, buffer=3.0):
    state3 = np.ascontiguousarray(state3, dtype=np.float64)
    ranges = np.ascontiguousarray(ranges, dtype=np.float32)
    l1, l2, lohi = np.zeros(3), np.zeros(3), np.zeros(2, dtype=np.int32)
    ok = lib().orc_find_half_spaces(thresh, divider, buffer, _dp(state3), angle_min, angle_max, angle_inc,
                                    _fp(ranges), len(ranges), _dp(l1), _dp(l2), _ip(lohi))
    return bool(ok), l1, l2, lohi
