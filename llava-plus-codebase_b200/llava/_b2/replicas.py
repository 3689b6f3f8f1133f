"""Data-parallel replicas: the path shards by sample, one process per GPU, full weight replica per
GPU and NO collective on the hot path. The reference scales eval the same way (one process per GPU +
get_chunk: llava/eval/model_vqa_loader.py:19-27, scripts/v1_5/eval/vqav2.sh:11-21) and merges results with
`cat`. Here the only communication is the eval-harness gather of token ids / last-position logits and the
max-over-ranks timing reduction, over NCCL (NVLink) on GPUs or gloo in the CPU tests.
"""
import math

import torch
import torch.distributed as dist


def shard_range(n_items, rank, world_size):
    """Contiguous chunk [lo, hi) of rank `rank`, ceil-sized chunks like the reference's split_list/get_chunk."""
    if world_size <= 1:
        return 0, n_items
    chunk = math.ceil(n_items / world_size)
    lo = min(rank * chunk, n_items)
    return lo, min(lo + chunk, n_items)


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def gather_rows(local, total_rows):
    """All-gather a [rows_local, ...] tensor sharded with shard_range back into [total_rows, ...] on every rank
    (token ids [B_local, N] int32 or last-position logits [B_local, V]); ragged shards are padded."""
    world = _world()
    if world == 1:
        return local
    chunk = math.ceil(total_rows / world)
    pad = torch.zeros((chunk,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    rows = []
    for r in range(world):
        lo, hi = shard_range(total_rows, r, world)
        rows.append(out[r][: hi - lo])
    return torch.cat(rows, dim=0)


def max_over_ranks(value, device="cpu"):
    """Timing reduction: the step time of a replica job is the slowest rank's."""
    if _world() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device="cpu"):
    if _world() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
