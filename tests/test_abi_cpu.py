"""CPU tests of the boundary: the C-ABI library loads and exports every symbol include/f110_mpc_b200.h
declares; without a CUDA device the compute entries fail loudly (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "f110_mpc_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(f110_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_entry_points(pkg):
    syms = declared_symbols()
    assert set(syms) == set(pkg.EXPORTS)


def test_library_exports_every_declared_symbol(pkg):
    pkg.build()
    out = subprocess.run(["nm", "-D", "--defined-only", pkg.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(l.split()[-1] for l in out.splitlines() if l.strip())
    for s in declared_symbols():
        assert s in exported, s
    L = ctypes.CDLL(pkg.LIB_PATH)        # loads without a GPU
    for s in declared_symbols():
        assert hasattr(L, s)


def test_helpers_and_defaults(pkg):
    L = pkg.lib()
    assert L.f110_mpc_record_doubles(30) == 101 and L.f110_mpc_num_variables(30) == 153 and L.f110_mpc_num_constraints(30) == 215
    c = pkg.default_config()
    assert c.horizon == 30 and c.dt == pytest.approx(0.009999999776482582, abs=0) and c.wheelbase == pytest.approx(0.3301999866962433, abs=0)
    assert list(c.q) == [10.0, 10.0, 0.0] and list(c.r) == [0.1, 5.0] and c.u_max[1] == pytest.approx(0.4300000071525574, abs=0)
    s = pkg.default_settings()
    assert (s.rho, s.sigma, s.alpha, s.eps_abs, s.max_iter, s.check_termination, s.scaling) == (0.1, 1e-6, 1.6, 1e-3, 4000, 25, 10)


def test_no_cpu_fallback(pkg):
    if pkg.lib().f110_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        pkg.MpcSolver(max_batch=4)


def test_product_does_not_touch_oracle():
    # the oracle is test infrastructure: nothing under the package may import, link or load it
    pk = os.path.join(ROOT, "f110-mpc_b200")
    for dp, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "liboracle" not in txt and "oracle_py" not in txt and "osqp_restated" not in txt, f


def test_header_is_plain_c_and_struct_layouts_match_bindings(pkg, tmp_path):
    # the boundary is a C ABI: the header must compile as C99 on its own, and the ctypes mirrors in the Python harness must have
    # the compiler's field offsets (a silent mismatch would shift every config value)
    import ctypes as C
    structs = {"f110_mpc_config": pkg.MpcConfig, "f110_solver_settings": pkg.SolverSettings, "f110_cycle_config": pkg.CycleConfig}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "f110_mpc_b200.h"', 'int main(void) {']
    for name, cls in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (name, name))
        for fname, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (name, fname, name, fname))
    lines += ['return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for name, cls in structs.items():
        assert int(out[name]) == C.sizeof(cls), name
        for fname, _ in cls._fields_:
            assert int(out["%s.%s" % (name, fname)]) == getattr(cls, fname).offset, (name, fname)


def test_new_entry_points_check_their_arguments_without_a_gpu(pkg):
    # argument errors are reported before any CUDA call: they must come back as F110_ERR_ARG (1) on a box without a GPU too
    import ctypes as C
    L = pkg.lib()
    n = C.c_size_t()
    assert L.f110_gather_bytes(8, 4100, 64, C.byref(n)) == 0 and n.value == 64 * 8 * 4 + 64 * 8 * 4100 * 32   # slots x world flags (a multiple of 256 bytes here), then the rows
    assert L.f110_gather_bytes(0, 4100, 64, C.byref(n)) == 1 and L.f110_gather_bytes(8, 0, 64, C.byref(n)) == 1
    assert L.f110_gather_slot(None, 2, 0, 8, 4, 0, None, None) == 1
    assert L.f110_stream_signal(None, None, 1) == 1 and L.f110_stream_wait_flags(None, None, 2, 0, 1) == 1
    cfg, st = pkg.default_config(), pkg.default_settings()
    h = C.c_void_p()
    devs = (C.c_int * 2)(0, 0)
    assert L.f110_mpc_create_multi(C.byref(cfg), C.byref(st), 64, devs, 2, C.byref(h)) == 1          # a device listed twice
    assert b"twice" in L.f110_last_error()
    assert L.f110_mpc_create_multi(C.byref(cfg), C.byref(st), 0, devs, 1, C.byref(h)) == 1
    t = C.c_int()
    assert L.f110_cycle_submit(None, None, 1, None, None, None, None, 1, 1, None, 1, C.byref(t)) == 1
    assert L.f110_cycle_wait(None, 0, None, None, None, None, None, None) == 1
    assert L.f110_cycle_gathered_view(None, 0, None, None) == 1
    assert L.f110_cycle_set_depth(None, 2) == 1
    assert L.f110_fleet_create(None, None, 1, None, 1, 1, None, 1, 2, 4, 0.01, C.byref(h)) == 1
    assert L.f110_fleet_run(None, 1, None, None) == 1 and L.f110_fleet_reset(None, None, None) == 1
    assert L.f110_mpc_multi_devices(None) == 0
    # the state-box option is part of the config struct and of the row count
    c = pkg.default_config(30, 0, state_lim=1.0)
    assert L.f110_mpc_num_rows(C.byref(c)) == 7 * 30 + 5 + 3 * 31
    c = pkg.default_config(30, 0, rate_delta=0.03)
    assert L.f110_mpc_num_rows(C.byref(c)) == 8 * 30 + 5
