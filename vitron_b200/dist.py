"""Request-level data parallelism for the vision-LLM path.

The path shards over independent requests: every rank holds a full weight replica and a local KV
cache, processes its contiguous slice of the request batch, and the ONLY collective is one
all_gather of the per-request results (generated ids, or last-token logits) at the end. The reference
has no inference-time collective at all (its only multi-GPU inference code, modules/i2vgen-xl/tools/
inferences/inference_i2vgen_entrance.py:59-85, runs whole replicas per rank).
Backend: NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous, balanced slice [lo, hi) of n_items for `rank` (first n % world ranks get one more)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_requests(items, rank=None, world=None):
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_range(len(items), rank, world)
    return items[lo:hi]


def gather_results(local, n_total, pad_value=0):
    """local [n_local, T] (n_local may differ by one across ranks) -> [n_total, T] on every rank, in
    request order. One all_gather over equal-sized padded blocks."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    per = (n_total + world - 1) // world
    block = torch.full((per, *local.shape[1:]), pad_value, dtype=local.dtype, device=local.device)
    block[:local.shape[0]] = local
    out = torch.empty((world * per, *local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, block.contiguous())
    parts = []
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        parts.append(out[r * per:r * per + (hi - lo)])
    return torch.cat(parts, 0)
