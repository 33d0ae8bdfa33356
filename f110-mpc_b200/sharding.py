"""Multi-GPU plumbing: QPs are independent, so the batch is cut into contiguous per-rank shards and the only
collective is one gather of the chosen controls.  Works on any torch.distributed backend
(nccl on the GPU box, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(total, world, rank):
    """Contiguous, balanced [lo, hi) of a batch of `total` items for `rank` of `world`."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_by_scenario(n_scenarios, per_scenario, world, rank):
    """Config-4 style: whole scenarios per rank so a scenario's grid, pose and its lane x path QPs stay together.
    Returns (scenario lo, hi), (QP lo, hi)."""
    lo, hi = shard_range(n_scenarios, world, rank)
    return (lo, hi), (lo * per_scenario, hi * per_scenario)


def pack_result(u0, status, iters):
    """(B,2) f64, (B,) i32, (B,) i32 -> (B,4) f64 rows (u0_v, u0_steer, status, iters): what is gathered."""
    out = torch.empty(u0.shape[0], 4, dtype=torch.float64, device=u0.device)
    out[:, :2] = u0
    out[:, 2] = status
    out[:, 3] = iters
    return out


def gather_results(local, world, max_rows=None, sizes=None):
    """All-gather per-rank result rows (possibly ragged) into batch order.  `local`: (b_r, 4) f64.
    Pass `sizes` (rows per rank) when they are known up front: it saves the size exchange and its host sync."""
    if world == 1:
        return local
    if sizes is None:
        counts = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
        all_counts = [torch.zeros_like(counts) for _ in range(world)]
        dist.all_gather(all_counts, counts)
        sizes = [int(c.item()) for c in all_counts]
    width = max_rows or max(sizes)
    if local.shape[0] == width and local.is_contiguous():
        padded = local                      # equal shards: gather straight from the kernel's output rows
    else:
        padded = torch.zeros(width, local.shape[1], dtype=local.dtype, device=local.device)
        padded[: local.shape[0]] = local
    buf = torch.empty(world * width, local.shape[1], dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, padded)
    if all(sz == width for sz in sizes):
        return buf
    return torch.cat([buf[r * width: r * width + sizes[r]] for r in range(world)], dim=0)
