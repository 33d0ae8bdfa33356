"""CPU check of the bench.py contract on the arm that runs without a GPU (--impl reference): one JSON line with the
keys the driver reads; torchrun ranks other than 0 print nothing."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "solves/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 100


def test_reference_arm_other_ranks_are_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, cwd=ROOT, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_committed_bench_lines_follow_the_contract():
    # profiles/r1_bench_product.json / _reference.json are verbatim bench.py lines from a B200; every key the driver reads is there
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    prod = json.load(open(os.path.join(root, "profiles", "r1_bench_product.json")))
    ref = json.load(open(os.path.join(root, "profiles", "r1_bench_reference.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in prod, k
    assert prod["metric"] == ref["metric"] and prod["unit"] == ref["unit"] and prod["config"]["workload"] == ref["config"]["workload"]
    assert prod["dtype"] == "f64" and prod["data"] == "synthetic" and prod["vs_baseline"] is None and prod["higher_is_better"] is True
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(prod["e2e"])
    assert prod["e2e"]["h2d_bytes_per_step"] > 0 and prod["e2e"]["d2h_bytes_per_step"] > 0 and prod["e2e"]["value"] < prod["value"]
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(prod["roofline"])
    assert abs(prod["roofline"]["frac"] - prod["roofline"]["achieved"] / prod["roofline"]["peak"]) < 1e-9
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(prod["cpu_baseline"]) and prod["cpu_baseline"]["kind"] == "port"
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(prod["clocks"])
    assert prod["gpu_launches"] >= prod["steps"] and prod["warmup"] >= 3
    assert ref["impl"] == "reference" and ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["e2e"]["value"] == ref["value"]


def test_config4_batch_definition_and_sharding():
    # BASELINE config 4: 7 lanes x 20 mini-paths x 64 scenarios = 8960 QPs, scenario-major so whole scenarios shard per rank
    import importlib
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    W = importlib.import_module("f110-mpc_b200.workloads")
    SH = importlib.import_module("f110-mpc_b200.sharding")
    recs = bench.config4_records(W)
    assert recs.shape == (8960, 11 + 3 * 30)
    assert (recs[:, 3] == 4.5).all() and (recs[:, 4] == 0.0).all()
    per_sc = 7 * 20
    # within a scenario: 7 distinct start states (lanes), each repeated for its 20 paths; references differ per path
    sc0 = recs[:per_sc]
    assert len({tuple(r[:3]) for r in sc0}) == 7
    assert (sc0[:20, :3] == sc0[0, :3]).all() and len({tuple(r[11:14 + 3 * 28]) for r in sc0[:20]}) == 20
    lane_step = sc0[20, :2] - sc0[0, :2]
    assert abs(float((lane_step ** 2).sum()) ** 0.5 - 0.25) < 1e-9          # lanes 0.25 m apart along the left normal
    covered = []
    for world in (1, 2, 4, 8):
        for rank in range(world):
            (s_lo, s_hi), (q_lo, q_hi) = SH.shard_by_scenario(64, per_sc, world, rank)
            assert (q_lo, q_hi) == (s_lo * per_sc, s_hi * per_sc) and s_hi - s_lo == 64 // world
            if world == 8:
                covered += list(range(q_lo, q_hi))
    assert covered == list(range(8960))
