"""f110-mpc on B200 — Python harness over the C ABI (include/f110_mpc_b200.h).

The product is the CUDA library `libf110mpc_b200.so` (hand-written sm_100a kernels behind a C ABI) and the
C++ host classes in `host/`.  This module is only the ctypes binding used by tests/ and bench.py; torch is
used for device buffers, streams and torch.distributed — plumbing, not compute.  There is no CPU fallback:
loading fails loudly if the CUDA library is missing, and every compute call fails without a CUDA device.

Import with `importlib.import_module("f110-mpc_b200")` (the directory name is not a Python identifier).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("F110_LIB", os.path.join(_HERE, "libf110mpc_b200.so"))  # F110_LIB: tuning builds only

HOST_LIB_PATH = os.path.join(_HERE, "libf110mpc_host.so")

SOLVED, SOLVED_INACCURATE, MAX_ITER_REACHED, UNSOLVED = 1, 2, -2, -10
PRIMAL_INFEASIBLE, DUAL_INFEASIBLE = -3, -4


class MpcConfig(C.Structure):
    """f110_mpc_config"""
    _fields_ = [("horizon", C.c_int32), ("gap_mode", C.c_int32), ("dt", C.c_double), ("wheelbase", C.c_double),
                ("q", C.c_double * 3), ("r", C.c_double * 2), ("u_des", C.c_double * 2), ("u_min", C.c_double * 2),
                ("u_max", C.c_double * 2), ("rate_rows", C.c_int32), ("state_rows", C.c_int32), ("rate_delta", C.c_double),
                ("state_lim", C.c_double)]


class SolverSettings(C.Structure):
    """f110_solver_settings"""
    _fields_ = [("rho", C.c_double), ("sigma", C.c_double), ("alpha", C.c_double), ("eps_abs", C.c_double),
                ("eps_rel", C.c_double), ("eps_prim_inf", C.c_double), ("eps_dual_inf", C.c_double),
                ("adaptive_rho_tolerance", C.c_double), ("max_iter", C.c_int32), ("check_termination", C.c_int32),
                ("scaling", C.c_int32), ("adaptive_rho", C.c_int32), ("adaptive_rho_interval", C.c_int32),
                ("warm_start", C.c_int32), ("scaled_termination", C.c_int32), ("reserved", C.c_int32)]


class CycleConfig(C.Structure):
    """f110_cycle_config"""
    _fields_ = [("n_beams", C.c_int32), ("angle_min", C.c_float), ("angle_max", C.c_float), ("angle_increment", C.c_float),
                ("occ_size", C.c_int32), ("occ_discrete", C.c_float), ("occ_dilation", C.c_float),
                ("follow_gap_thresh", C.c_float), ("fov_divider", C.c_float), ("buffer", C.c_float), ("lookahead", C.c_float),
                ("use_half_spaces", C.c_int32), ("qp_mode", C.c_int32), ("reserved", C.c_int32), ("v_lin", C.c_double)]


EXPORTS = ["f110_mpc_default_config", "f110_solver_default_settings", "f110_mpc_record_doubles",
           "f110_mpc_num_variables", "f110_mpc_num_constraints", "f110_mpc_num_rows", "f110_last_error", "f110_device_count",
           "f110_mpc_create", "f110_mpc_destroy", "f110_mpc_solve_host", "f110_mpc_solve_device", "f110_mpc_reset",
           "f110_mpc_last_launches", "f110_mpc_set_packed_output", "f110_collision_check_device", "f110_collision_check_host", "f110_bench_fp64_fma", "f110_cycle_default_config",
           "f110_cycle_device", "f110_cycle_host", "f110_cycle_buffers", "f110_cycle_submit", "f110_cycle_wait",
           "f110_mpc_create_multi", "f110_mpc_destroy_multi", "f110_mpc_solve_multi_host", "f110_mpc_multi_devices",
           "f110_mpc_multi_uses_peer_stores", "f110_mpc_multi_last_shard", "f110_gather_bytes", "f110_gather_create",
           "f110_gather_open", "f110_gather_close", "f110_gather_slot", "f110_stream_signal", "f110_stream_wait_flags",
           "f110_cycle_set_gather", "f110_cycle_gathered_view", "f110_cycle_set_depth", "f110_fleet_create", "f110_fleet_destroy", "f110_fleet_reset", "f110_fleet_run",
           "f110_fleet_get_pose"]


def build(force=False, verbose=False):
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(_HERE, d, f) for d in ("csrc", "host") for f in os.listdir(os.path.join(_HERE, d))]
    srcs += [os.path.join(_HERE, "Makefile"), os.path.join(_HERE, "..", "include", "f110_mpc_b200.h")]
    outs = [LIB_PATH, HOST_LIB_PATH]
    stale = force or any(not os.path.exists(o) for o in outs) or any(
        os.path.getmtime(s) > min(os.path.getmtime(o) for o in outs) for s in srcs)
    if stale:
        out = subprocess.run(["make", "-j8", "-C", _HERE], capture_output=True, text=True)
        if verbose or out.returncode:
            print(out.stdout + out.stderr)
        if out.returncode:
            raise RuntimeError("nvcc build of libf110mpc_b200.so failed")
    return LIB_PATH


_lib = None


def lib():
    """Load the CUDA library (raises if it is not built — no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libf110mpc_b200.so is not built (run __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int32)
        L.f110_last_error.restype = C.c_char_p
        L.f110_mpc_default_config.argtypes = [C.POINTER(MpcConfig)]
        L.f110_solver_default_settings.argtypes = [C.POINTER(SolverSettings)]
        L.f110_mpc_create.argtypes = [C.POINTER(MpcConfig), C.POINTER(SolverSettings), C.c_int, C.c_int, C.POINTER(vp)]
        L.f110_mpc_destroy.argtypes = [vp]
        L.f110_mpc_reset.argtypes = [vp]
        L.f110_mpc_last_launches.argtypes = [vp]
        L.f110_mpc_set_packed_output.argtypes = [vp, vp]
        L.f110_mpc_solve_host.argtypes = [vp, C.c_int, dp, C.c_int, dp, dp, dp, ip, ip]
        L.f110_mpc_solve_device.argtypes = [vp, C.c_int, vp, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp]
        L.f110_collision_check_device.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_float] + [vp] * 9
        L.f110_collision_check_host.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_float] + [vp] * 8 + [C.c_int]
        L.f110_bench_fp64_fma.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.f110_cycle_default_config.argtypes = [C.POINTER(CycleConfig)]
        L.f110_cycle_device.argtypes = [vp, C.POINTER(CycleConfig), C.c_int, vp, vp, vp, vp, C.c_int, C.c_int, vp, C.c_int,
                                        vp, vp, vp, vp, vp, vp]
        L.f110_cycle_host.argtypes = [vp, C.POINTER(CycleConfig), C.c_int, vp, vp, vp, vp, C.c_int, C.c_int, vp, C.c_int, vp, vp, vp, vp, vp]
        L.f110_cycle_buffers.argtypes = [vp] + [C.POINTER(vp)] * 5
        L.f110_cycle_submit.argtypes = [vp, C.POINTER(CycleConfig), C.c_int, vp, vp, vp, vp, C.c_int, C.c_int, vp, C.c_int, C.POINTER(C.c_int)]
        L.f110_cycle_wait.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp]
        L.f110_mpc_create_multi.argtypes = [C.POINTER(MpcConfig), C.POINTER(SolverSettings), C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
        L.f110_mpc_destroy_multi.argtypes = [vp]
        L.f110_mpc_solve_multi_host.argtypes = [vp, C.c_int, C.c_int, dp, C.c_int, dp, ip, ip]
        L.f110_mpc_multi_devices.argtypes = [vp]
        L.f110_mpc_multi_uses_peer_stores.argtypes = [vp, C.c_int]
        L.f110_mpc_multi_last_shard.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.f110_gather_bytes.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]
        L.f110_gather_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp), C.c_char_p]
        L.f110_gather_open.argtypes = [C.c_int, C.c_char_p, C.POINTER(vp)]
        L.f110_gather_close.argtypes = [C.c_int, vp, C.c_int]
        L.f110_gather_slot.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_longlong, C.POINTER(vp), C.POINTER(vp)]
        L.f110_stream_signal.argtypes = [vp, vp, C.c_int32]
        L.f110_stream_wait_flags.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int32]
        L.f110_cycle_set_gather.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int]
        L.f110_cycle_gathered_view.argtypes = [vp, C.c_int, C.POINTER(dp), C.POINTER(C.c_size_t)]
        L.f110_cycle_set_depth.argtypes = [vp, C.c_int]
        L.f110_fleet_create.argtypes = [vp, C.POINTER(CycleConfig), C.c_int, vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_double,
                                        C.POINTER(vp)]
        L.f110_fleet_destroy.argtypes = [vp]
        L.f110_fleet_reset.argtypes = [vp, vp, vp]
        L.f110_fleet_run.argtypes = [vp, C.c_int, vp, vp]
        L.f110_fleet_get_pose.argtypes = [vp, vp]
        _lib = L
    return _lib


def _check(rc, what):
    if rc:
        raise RuntimeError("%s failed (rc=%d): %s" % (what, rc, lib().f110_last_error().decode()))


def default_config(horizon=30, gap_mode=0, rate_delta=None, state_lim=None):
    """rate_delta: max steering change per step (rad) -> N steering-rate rows appended; state_lim: d of Constraints::SetXLims ->
    3(N+1) state-box rows appended; None = the reference's row set."""
    c = MpcConfig()
    lib().f110_mpc_default_config(C.byref(c))
    c.horizon = horizon
    c.gap_mode = gap_mode
    if rate_delta is not None:
        c.rate_rows, c.rate_delta = 1, rate_delta
    if state_lim is not None:
        c.state_rows, c.state_lim = 1, state_lim
    return c


def default_cycle_config(**kw):
    c = CycleConfig()
    lib().f110_cycle_default_config(C.byref(c))
    for k, v in kw.items():
        if not hasattr(c, k):
            raise KeyError(k)
        setattr(c, k, v)
    return c


def default_settings(**kw):
    s = SolverSettings()
    lib().f110_solver_default_settings(C.byref(s))
    for k, v in kw.items():
        if not hasattr(s, k):
            raise KeyError(k)
        setattr(s, k, v)
    return s


def record_doubles(N):
    return 11 + 3 * N


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None else None


def _tp(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class MpcSolver:
    """One handle per GPU (f110_mpc_create).  `solve_host` is the reference-facing call (host buffers in,
    host buffers out); `solve_device` takes torch CUDA tensors and is stream-ordered."""

    def __init__(self, config=None, settings=None, max_batch=4096, device=0):
        self.config = config or default_config()
        self.settings = settings or default_settings()
        self.N = self.config.horizon
        self.n = 5 * self.N + 3
        self.m = 7 * self.N + 5 + (self.N if self.config.rate_rows else 0) + (3 * (self.N + 1) if self.config.state_rows else 0)   # f110_mpc_num_rows
        self.max_batch = max_batch
        self.device = device
        self._h = C.c_void_p()
        _check(lib().f110_mpc_create(C.byref(self.config), C.byref(self.settings), max_batch, device, C.byref(self._h)),
               "f110_mpc_create")

    def solve_host(self, recs, want_xy=True, out=None):
        """`out` may carry preallocated (e.g. pinned) u0/status/iters/x/y arrays."""
        recs = np.ascontiguousarray(recs, dtype=np.float64)
        B = recs.shape[0]
        out = out or {}
        x = out.get("x", np.empty((B, self.n)) if want_xy else None)
        y = out.get("y", np.empty((B, self.m)) if want_xy else None)
        u0 = out.get("u0", np.empty((B, 2)))
        status = out.get("status", np.empty(B, dtype=np.int32))
        iters = out.get("iters", np.empty(B, dtype=np.int32))
        _check(lib().f110_mpc_solve_host(self._h, B, _dp(recs), recs.shape[1], _dp(x), _dp(y), _dp(u0), _ip(status),
                                         _ip(iters)), "f110_mpc_solve_host")
        return dict(x=x, y=y, u0=u0, status=status, iters=iters)

    def solve_device(self, recs, x=None, y=None, u0=None, status=None, iters=None, rho_updates=None, info=None,
                     stream=None, count=None, packed=None):
        """recs etc. are torch CUDA tensors (float64 / int32, contiguous).  packed: optional (B,4) f64 rows
        (u0_v, u0_steer, status, iters) written by the solve kernel for the multi-GPU gather."""
        B = recs.shape[0] if count is None else count
        if packed is not None:
            _check(lib().f110_mpc_set_packed_output(self._h, _tp(packed)), "f110_mpc_set_packed_output")
        sp = C.c_void_p(stream) if stream else None
        _check(lib().f110_mpc_solve_device(self._h, B, _tp(recs), recs.stride(0), _tp(x), _tp(y), _tp(u0), _tp(status),
                                           _tp(iters), _tp(rho_updates), _tp(info), sp), "f110_mpc_solve_device")

    def cycle_device(self, cc, pose7, ranges, prev_steer, table_xy, wp_xy, u0, status, iters, chosen, valid=None, stream=None):
        """Whole planning + control cycle on the device (f110_cycle_device); all arguments are torch CUDA tensors."""
        S = pose7.shape[0]
        sp = C.c_void_p(stream) if stream else None
        _check(lib().f110_cycle_device(self._h, C.byref(cc), S, _tp(pose7), _tp(ranges), _tp(prev_steer), _tp(table_xy),
                                       table_xy.shape[0], table_xy.shape[1], _tp(wp_xy), wp_xy.shape[0], _tp(u0), _tp(status),
                                       _tp(iters), _tp(chosen), _tp(valid), sp), "f110_cycle_device")

    def cycle_host(self, cc, pose7, ranges, prev_steer, table_xy, wp_xy, out=None):
        """f110_cycle_host: numpy in / numpy out.  Returns dict(u0, status, iters, chosen, valid)."""
        pose7 = np.ascontiguousarray(pose7, dtype=np.float64); ranges = np.ascontiguousarray(ranges, dtype=np.float32)
        table_xy = np.ascontiguousarray(table_xy, dtype=np.float64); wp_xy = np.ascontiguousarray(wp_xy, dtype=np.float32)
        S, P = pose7.shape[0], table_xy.shape[0]
        nqp = S if cc.qp_mode == 0 else S * P
        out = out or {}
        u0 = out.get("u0", np.empty((nqp, 2))); status = out.get("status", np.empty(nqp, dtype=np.int32))
        iters = out.get("iters", np.empty(nqp, dtype=np.int32)); chosen = out.get("chosen", np.empty(S, dtype=np.int32))
        valid = out.get("valid", np.empty((S, P), dtype=np.uint8))
        vp = lambda a: C.c_void_p(a.ctypes.data) if a is not None else None
        if prev_steer is not None:
            prev_steer = np.ascontiguousarray(prev_steer, dtype=np.float64)
        _check(lib().f110_cycle_host(self._h, C.byref(cc), S, vp(pose7), vp(ranges), vp(prev_steer), vp(table_xy), P, table_xy.shape[1],
                                     vp(wp_xy), wp_xy.shape[0], vp(u0), vp(status), vp(iters), vp(chosen), vp(valid)), "f110_cycle_host")
        return dict(u0=u0, status=status, iters=iters, chosen=chosen, valid=valid)

    def cycle_submit(self, cc, pose7, ranges, prev_steer, table_xy, wp_xy):
        """f110_cycle_submit: queue one cycle (inputs must be C-contiguous numpy arrays of the ABI's dtypes; pinned arrays are read
        in place and must stay untouched until cycle_wait).  Returns the ticket."""
        S, P = pose7.shape[0], table_xy.shape[0]
        vp = lambda a: C.c_void_p(a.ctypes.data) if a is not None else None
        for a, dt in ((pose7, np.float64), (ranges, np.float32), (table_xy, np.float64), (wp_xy, np.float32)):
            assert a.dtype == dt and a.flags["C_CONTIGUOUS"]
        t = C.c_int(-1)
        _check(lib().f110_cycle_submit(self._h, C.byref(cc), S, vp(pose7), vp(ranges), vp(prev_steer), vp(table_xy), P, table_xy.shape[1],
                                       vp(wp_xy), wp_xy.shape[0], C.byref(t)), "f110_cycle_submit")
        self._inflight = getattr(self, "_inflight", {})
        self._inflight[t.value] = (S, P, S if cc.qp_mode == 0 else S * P)
        return t.value

    def cycle_wait(self, ticket, out=None, gathered=None):
        """f110_cycle_wait: block until the cycle's results are on the host.  Returns dict(u0, status, iters, chosen, valid)."""
        info = getattr(self, "_inflight", {}).pop(ticket, None)
        if info is None:   # not a ticket of this handle: let the library say so
            _check(lib().f110_cycle_wait(self._h, ticket, None, None, None, None, None, None), "f110_cycle_wait")
            raise RuntimeError("f110_cycle_wait: unknown ticket %d" % ticket)
        S, P, nqp = info
        out = out or {}
        u0 = out.get("u0", np.empty((nqp, 2))); status = out.get("status", np.empty(nqp, dtype=np.int32))
        iters = out.get("iters", np.empty(nqp, dtype=np.int32)); chosen = out.get("chosen", np.empty(S, dtype=np.int32))
        valid = out.get("valid", np.empty((S, P), dtype=np.uint8))
        vp = lambda a: C.c_void_p(a.ctypes.data) if a is not None else None
        _check(lib().f110_cycle_wait(self._h, ticket, vp(u0), vp(status), vp(iters), vp(chosen), vp(valid), vp(gathered)), "f110_cycle_wait")
        return dict(u0=u0, status=status, iters=iters, chosen=chosen, valid=valid)

    def gathered_view(self, ticket, world, rows):
        """Gather root, after cycle_wait(ticket): the cycle's gathered rows as a (world, rows, 4) numpy VIEW of the handle's pinned
        host buffer (no copy; valid until the second-next cycle_submit)."""
        ptr, n = C.POINTER(C.c_double)(), C.c_size_t()
        _check(lib().f110_cycle_gathered_view(self._h, ticket, C.byref(ptr), C.byref(n)), "f110_cycle_gathered_view")
        assert n.value == world * rows * 4
        return np.ctypeslib.as_array(ptr, shape=(world, rows, 4))

    def set_cycle_depth(self, depth):
        """f110_cycle_set_depth: cycles that may be in flight on this handle (1..4, default 2); only while none is in flight."""
        _check(lib().f110_cycle_set_depth(self._h, depth), "f110_cycle_set_depth")

    def set_gather(self, ring_ptr, world, rank, rows_per_rank, slots):
        """Attach (or, with ring_ptr None, detach) a gather ring to the asynchronous cycle entry."""
        _check(lib().f110_cycle_set_gather(self._h, C.c_void_p(ring_ptr) if ring_ptr else None, world, rank, rows_per_rank, slots),
               "f110_cycle_set_gather")

    def cycle_buffers(self, scenes, nrec=None):
        """Torch views (no copy) of the device buffers the last cycle filled:
        dict(grid (S, blocks^2) f32, offset (S,2) f32, l1l2 (S,6) f64, recs (S, 11+3N) f64, best_global (S,) i32)."""
        import torch
        ptrs = [C.c_void_p() for _ in range(5)]
        _check(lib().f110_cycle_buffers(self._h, *[C.byref(p) for p in ptrs]), "f110_cycle_buffers")

        class _Dev:
            def __init__(self, ptr, shape, typestr):
                self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}

        def view(ptr, shape, typestr):
            return torch.as_tensor(_Dev(ptr.value, shape, typestr), device="cuda:%d" % self.device)
        blocks2 = 100 * 100
        return dict(grid=view(ptrs[0], (scenes, blocks2), "<f4"), offset=view(ptrs[1], (scenes, 2), "<f4"),
                    l1l2=view(ptrs[2], (scenes, 6), "<f8"),
                    recs=view(ptrs[3], (nrec or scenes, (record_doubles(self.N) + 1) & ~1), "<f8")[:, :record_doubles(self.N)],
                    best_global=view(ptrs[4], (scenes,), "<i4"))

    def reset(self):
        _check(lib().f110_mpc_reset(self._h), "f110_mpc_reset")

    @property
    def last_launches(self):
        return lib().f110_mpc_last_launches(self._h)

    def close(self):
        if self._h:
            lib().f110_mpc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Fleet:
    """f110_fleet_*: `cars` simulated cars in closed loop on the device (the reference's plan / control / drive state machine against
    the kinematic plant).  `solver` must have warm_start = 1 and max_batch >= cars."""
    PLAN, IDLE, DROP, CONTROL = 0, 1, 2, 3

    def __init__(self, solver, cc, cars, table_xy, wp_xy, drive_every=2, scan_every=4, dt_tick=0.01):
        self.solver, self.cars = solver, cars
        table_xy = np.ascontiguousarray(table_xy, dtype=np.float64); wp_xy = np.ascontiguousarray(wp_xy, dtype=np.float32)
        self._h = C.c_void_p()
        _check(lib().f110_fleet_create(solver._h, C.byref(cc), cars, C.c_void_p(table_xy.ctypes.data), table_xy.shape[0], table_xy.shape[1],
                                       C.c_void_p(wp_xy.ctypes.data), wp_xy.shape[0], drive_every, scan_every, dt_tick, C.byref(self._h)),
               "f110_fleet_create")

    def reset(self, pose3, ranges):
        pose3 = np.ascontiguousarray(pose3, dtype=np.float64); ranges = np.ascontiguousarray(ranges, dtype=np.float32)
        assert pose3.shape == (self.cars, 3) and ranges.shape[0] == self.cars
        _check(lib().f110_fleet_reset(self._h, C.c_void_p(pose3.ctypes.data), C.c_void_p(ranges.ctypes.data)), "f110_fleet_reset")

    def run(self, ticks, log=True):
        """Returns (log_i (ticks, cars, 4) int32, log_d (ticks, cars, 13) float64) or None without logging."""
        if not log:
            _check(lib().f110_fleet_run(self._h, ticks, None, None), "f110_fleet_run")
            return None
        li = np.empty((ticks, self.cars, 4), dtype=np.int32); ld = np.empty((ticks, self.cars, 13))
        _check(lib().f110_fleet_run(self._h, ticks, C.c_void_p(li.ctypes.data), C.c_void_p(ld.ctypes.data)), "f110_fleet_run")
        return li, ld

    def poses(self):
        p = np.empty((self.cars, 3))
        _check(lib().f110_fleet_get_pose(self._h, C.c_void_p(p.ctypes.data)), "f110_fleet_get_pose")
        return p

    def close(self):
        if self._h:
            lib().f110_fleet_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiGpuSolver:
    """f110_mpc_create_multi: one process, several GPUs; shards of whole units, gather by peer stores into devices[0]."""

    def __init__(self, config=None, settings=None, max_batch=4096, devices=(0,)):
        self.config = config or default_config()
        self.settings = settings or default_settings()
        self.devices = list(devices)
        arr = (C.c_int * len(self.devices))(*self.devices)
        self._h = C.c_void_p()
        _check(lib().f110_mpc_create_multi(C.byref(self.config), C.byref(self.settings), max_batch, arr, len(self.devices), C.byref(self._h)),
               "f110_mpc_create_multi")

    def solve_host(self, recs, unit=1):
        recs = np.ascontiguousarray(recs, dtype=np.float64)
        B = recs.shape[0]
        u0, status, iters = np.empty((B, 2)), np.empty(B, dtype=np.int32), np.empty(B, dtype=np.int32)
        _check(lib().f110_mpc_solve_multi_host(self._h, B, unit, _dp(recs), recs.shape[1], _dp(u0), _ip(status), _ip(iters)),
               "f110_mpc_solve_multi_host")
        return dict(u0=u0, status=status, iters=iters)

    def last_shards(self):
        out = []
        for i in range(len(self.devices)):
            a, b = C.c_int(), C.c_int()
            _check(lib().f110_mpc_multi_last_shard(self._h, i, C.byref(a), C.byref(b)), "f110_mpc_multi_last_shard")
            out.append((a.value, b.value))
        return out

    def uses_peer_stores(self, i):
        return bool(lib().f110_mpc_multi_uses_peer_stores(self._h, i))

    def close(self):
        if self._h:
            lib().f110_mpc_destroy_multi(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GatherRing:
    """The NVLink gather ring (f110_gather_*).  Root: GatherRing.create(...) -> .handle (64 bytes) goes to the peers; peers:
    GatherRing.open(handle, ...).  `slot(cycle)` -> (device pointer of this rank's rows, device pointer of its flag)."""

    def __init__(self, ptr, device, world, rank, rows, slots, opened, handle=None):
        self.ptr, self.device, self.world, self.rank, self.rows, self.slots, self.opened, self.handle = ptr, device, world, rank, rows, slots, opened, handle

    @classmethod
    def create(cls, device, world, rows, slots):
        p = C.c_void_p()
        h = C.create_string_buffer(64)
        _check(lib().f110_gather_create(device, world, rows, slots, C.byref(p), h), "f110_gather_create")
        return cls(p.value, device, world, 0, rows, slots, False, h.raw)

    @classmethod
    def open(cls, handle, device, world, rank, rows, slots):
        p = C.c_void_p()
        _check(lib().f110_gather_open(device, handle, C.byref(p)), "f110_gather_open")
        return cls(p.value, device, world, rank, rows, slots, True, handle)

    def slot(self, cycle):
        rows, flag = C.c_void_p(), C.c_void_p()
        _check(lib().f110_gather_slot(C.c_void_p(self.ptr), self.world, self.rank, self.rows, self.slots, cycle, C.byref(rows), C.byref(flag)),
               "f110_gather_slot")
        return rows.value, flag.value

    def nbytes(self):
        n = C.c_size_t()
        _check(lib().f110_gather_bytes(self.world, self.rows, self.slots, C.byref(n)), "f110_gather_bytes")
        return n.value

    def close(self):
        if self.ptr:
            lib().f110_gather_close(self.device, C.c_void_p(self.ptr), 1 if self.opened else 0)
            self.ptr = None


def fp64_fma_peak_tflops(device=0, iters=20000):
    t = C.c_double()
    _check(lib().f110_bench_fp64_fma(device, iters, C.byref(t)), "f110_bench_fp64_fma")
    return t.value


def collision_check_host(grid, offset, rot, pose_xy, table_xy, blocks=100, discrete=0.1, device=0):
    """grid (S, blocks*blocks) f32 col-major cells; offset (S,2) f32; rot (S,4) f64; pose_xy (S,2) f64;
    table_xy (P, samples, 2) f64.  Returns valid (S,P) u8, free_count (S,P) i32, end_world (S,P,2) f32."""
    grid = np.ascontiguousarray(grid, dtype=np.float32)
    offset = np.ascontiguousarray(offset, dtype=np.float32)
    rot = np.ascontiguousarray(rot, dtype=np.float64)
    pose_xy = np.ascontiguousarray(pose_xy, dtype=np.float64)
    table_xy = np.ascontiguousarray(table_xy, dtype=np.float64)
    S = grid.shape[0]
    P, samples = table_xy.shape[0], table_xy.shape[1]
    valid = np.empty((S, P), dtype=np.uint8)
    free = np.empty((S, P), dtype=np.int32)
    endw = np.empty((S, P, 2), dtype=np.float32)
    vp = lambda a: C.c_void_p(a.ctypes.data)
    _check(lib().f110_collision_check_host(S, P, samples, blocks, discrete, vp(grid), vp(offset), vp(rot), vp(pose_xy),
                                           vp(table_xy), vp(valid), vp(free), vp(endw), device),
           "f110_collision_check_host")
    return valid, free, endw


def collision_check_device(grid, offset, rot, pose_xy, table_xy, valid, free_count, end_world, blocks=100, discrete=0.1,
                           stream=None):
    """torch CUDA tensors, stream-ordered."""
    S = grid.shape[0]
    P, samples = table_xy.shape[0], table_xy.shape[1]
    sp = C.c_void_p(stream) if stream else None
    _check(lib().f110_collision_check_device(S, P, samples, blocks, discrete, _tp(grid), _tp(offset), _tp(rot),
                                             _tp(pose_xy), _tp(table_xy), _tp(valid), _tp(free_count), _tp(end_world),
                                             sp), "f110_collision_check_device")


# ---- C++ host classes (f110-mpc_b200/host) through their C shims ---------------------------------------------------
_host = None


def host():
    """libf110mpc_host.so: the ROS-free C++ mirror of the reference classes (MPC, Constraints, OccGrid, ...)."""
    global _host
    if _host is None:
        if not os.path.exists(HOST_LIB_PATH):
            raise RuntimeError("libf110mpc_host.so is not built (run __graft_entry__.build())")
        lib()
        H = C.CDLL(HOST_LIB_PATH)
        dp, fp, ip = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int)
        H.f110h_linearize.argtypes = [C.c_double] * 4 + [dp] * 3
        H.f110h_traj_table.argtypes = [C.c_int, C.c_int, dp]
        H.f110h_car_to_world_R.argtypes = [dp, dp]
        H.f110h_fill_grid.argtypes = [dp, C.c_float, C.c_float, C.c_float, fp, C.c_int, fp, fp]
        H.f110h_find_half_spaces.argtypes = [dp, C.c_float, C.c_float, C.c_float, fp, C.c_int, dp, dp, ip]
        H.f110h_best_global_idx.argtypes = [fp, C.c_int, dp, dp]
        H.f110h_mpc_create.restype = C.c_void_p
        H.f110h_mpc_create.argtypes = [C.c_int, C.c_int, C.c_int]
        H.f110h_mpc_create_rate.restype = C.c_void_p
        H.f110h_mpc_create_rate.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int]
        H.f110h_mpc_destroy.argtypes = [C.c_void_p]
        H.f110h_mpc_update_scan.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, fp, C.c_int]
        H.f110h_mpc_update.argtypes = [C.c_void_p, dp, dp, dp, C.c_int, dp, dp, dp, ip, ip, dp]
        H.f110h_plan.argtypes = [C.c_int, C.c_int, dp, C.c_float, C.c_float, C.c_float, fp, C.c_int, fp, C.c_int, C.c_int, dp,
                                 C.POINTER(C.c_uint8), ip]
        H.f110h_closed_loop.argtypes = [C.c_int, C.c_double, C.c_int, C.c_int, fp, C.c_int, dp, C.c_float, C.c_float, C.c_float,
                                        fp, C.c_int, C.c_int, dp, ip]
        _host = H
    return _host


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def host_linearize(ori, v, steer, dt):
    A, B, Cc = np.zeros(9), np.zeros(6), np.zeros(3)
    host().f110h_linearize(ori, v, steer, dt, _dp(A), _dp(B), _dp(Cc))
    return A.reshape(3, 3), B.reshape(3, 2), Cc


def host_traj_table(steer_discrete=30, traj_discrete=50):
    out = np.zeros((steer_discrete + 1, traj_discrete, 3))
    host().f110h_traj_table(steer_discrete, traj_discrete, _dp(out))
    return out


def host_car_to_world_R(pose7):
    pose7 = np.ascontiguousarray(pose7, dtype=np.float64)
    R = np.zeros(4)
    host().f110h_car_to_world_R(_dp(pose7), _dp(R))
    return R


def host_fill_grid(pose7, angle_min, angle_max, angle_inc, ranges):
    pose7 = np.ascontiguousarray(pose7, dtype=np.float64)
    ranges = np.ascontiguousarray(ranges, dtype=np.float32)
    grid = np.zeros(100 * 100, dtype=np.float32)
    off = np.zeros(2, dtype=np.float32)
    host().f110h_fill_grid(_dp(pose7), angle_min, angle_max, angle_inc, _fp(ranges), len(ranges), _fp(grid), _fp(off))
    return grid, off


def host_find_half_spaces(state3, angle_min, angle_max, angle_inc, ranges):
    state3 = np.ascontiguousarray(state3, dtype=np.float64)
    ranges = np.ascontiguousarray(ranges, dtype=np.float32)
    l1, l2, lohi = np.zeros(3), np.zeros(3), np.zeros(2, dtype=np.int32)
    ok = host().f110h_find_half_spaces(_dp(state3), angle_min, angle_max, angle_inc, _fp(ranges), len(ranges), _dp(l1), _dp(l2),
                                       lohi.ctypes.data_as(C.POINTER(C.c_int)))
    return bool(ok), l1, l2, lohi


def host_best_global_idx(wp_xy, pose7):
    wp_xy = np.ascontiguousarray(wp_xy, dtype=np.float32)
    pose7 = np.ascontiguousarray(pose7, dtype=np.float64)
    head = np.zeros(len(wp_xy))
    idx = host().f110h_best_global_idx(_fp(wp_xy), len(wp_xy), _dp(pose7), _dp(head))
    return idx, head


class HostMPC:
    """The C++ `MPC` class (host/mpc.h): Update(State, Input, vector<State>&) -> solved_trajectory()."""

    def __init__(self, horizon=30, gap_mode=0, device=0, steer_rate_max=0.0):
        self.N = horizon
        self.m = 7 * horizon + 5 + (horizon if steer_rate_max > 0 else 0)
        self._h = host().f110h_mpc_create_rate(horizon, gap_mode, steer_rate_max, device)
        if not self._h:
            raise RuntimeError("MPC construction failed: " + lib().f110_last_error().decode())

    def update_scan(self, angle_min, angle_max, angle_inc, ranges):
        ranges = np.ascontiguousarray(ranges, dtype=np.float32)
        host().f110h_mpc_update_scan(self._h, angle_min, angle_max, angle_inc, _fp(ranges), len(ranges))

    def update(self, state3, input2, desired):
        state3 = np.ascontiguousarray(state3, dtype=np.float64)
        input2 = np.ascontiguousarray(input2, dtype=np.float64)
        desired = np.ascontiguousarray(desired, dtype=np.float64)
        N = self.N
        inputs = np.zeros((N, 2)); x = np.zeros(5 * N + 3); y = np.zeros(self.m); l1l2 = np.zeros(6)
        st, it = C.c_int(), C.c_int()
        n = host().f110h_mpc_update(self._h, _dp(state3), _dp(input2), _dp(desired), desired.shape[0], _dp(inputs), _dp(x), _dp(y),
                                    C.byref(st), C.byref(it), _dp(l1l2))
        return dict(inputs=inputs[:n], x=x, y=y, status=st.value, iters=it.value, l1=l1l2[:3], l2=l1l2[3:])

    def close(self):
        if self._h:
            host().f110h_mpc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def host_plan(pose7, angle_min, angle_max, angle_inc, ranges, wp_xy, steer_discrete=30, traj_discrete=50, device=0):
    """One planning cycle (project.cpp:73-157): returns (chosen index or -1, mini_path (S,3), valid flags, best_global)."""
    pose7 = np.ascontiguousarray(pose7, dtype=np.float64)
    ranges = np.ascontiguousarray(ranges, dtype=np.float32)
    wp_xy = np.ascontiguousarray(wp_xy, dtype=np.float32)
    path = np.zeros((traj_discrete, 3))
    valid = np.zeros(steer_discrete + 1, dtype=np.uint8)
    bg = C.c_int(-1)
    idx = host().f110h_plan(steer_discrete, traj_discrete, _dp(pose7), angle_min, angle_max, angle_inc, _fp(ranges), len(ranges),
                            _fp(wp_xy), len(wp_xy), device, _dp(path), valid.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(bg))
    return idx, path, valid, bg.value


def host_closed_loop(ticks, wp_xy, start_xyyaw, angle_min, angle_max, angle_inc, ranges, dt_tick=0.01, drive_every=2,
                     scan_every=4, device=0):
    """The C++ `project` orchestrator (host/project.h) driving a simulated car.  Returns (traj (ticks,5) = x, y, yaw,
    v, steer per tick; MPC cycles solved; planning cycles)."""
    wp_xy = np.ascontiguousarray(wp_xy, dtype=np.float32)
    start = np.ascontiguousarray(start_xyyaw, dtype=np.float64)
    ranges = np.ascontiguousarray(ranges, dtype=np.float32)
    traj = np.zeros((ticks, 5))
    plans = C.c_int(0)
    solved = host().f110h_closed_loop(ticks, dt_tick, drive_every, scan_every, _fp(wp_xy), len(wp_xy), _dp(start), angle_min,
                                      angle_max, angle_inc, _fp(ranges), len(ranges), device, _dp(traj), C.byref(plans))
    if solved < 0:
        raise RuntimeError("closed loop failed: " + lib().f110_last_error().decode())
    return traj, solved, plans.value
