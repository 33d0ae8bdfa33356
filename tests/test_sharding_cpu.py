"""World-size-2 gloo test of the multi-GPU host logic (sharding + the final gather); no GPU involved, the
per-rank "solve" is a stand-in that tags every QP with its global index."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, total, out_dir):
    sys.path.insert(0, ROOT)
    S = importlib.import_module("f110-mpc_b200.sharding")
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = S.shard_range(total, world, rank)
    idx = torch.arange(lo, hi, dtype=torch.float64)
    local = S.pack_result(torch.stack([idx, -idx], dim=1), torch.ones(hi - lo, dtype=torch.int32), (25 * (1 + idx % 3)).to(torch.int32))
    full = S.gather_results(local, world)
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), full.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [8960, 4097, 3])
def test_shard_and_gather_world2(tmp_path, total):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), total, str(tmp_path)), nprocs=world, join=True)
    idx = np.arange(total, dtype=np.float64)
    want = np.stack([idx, -idx, np.ones(total), 25 * (1 + idx % 3)], axis=1)
    for r in range(world):
        np.testing.assert_array_equal(np.load(tmp_path / ("rank%d.npy" % r)), want)


def test_shard_ranges_cover_batch():
    S = importlib.import_module("f110-mpc_b200.sharding")
    for total in (0, 1, 7, 4096, 8960):
        for world in (1, 2, 4, 8):
            spans = [S.shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[r][1] == spans[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    # config 4: 64 scenarios x 140 QPs over 8 ranks -> 8 scenarios = 1120 QPs each
    (slo, shi), (qlo, qhi) = S.shard_by_scenario(64, 140, 8, 3)
    assert (slo, shi, qlo, qhi) == (24, 32, 3360, 4480)
