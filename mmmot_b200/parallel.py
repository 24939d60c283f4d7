"""Frame-pair sharding across GPUs.

Frame-pairs are independent units (all GroupNorm coupling is intra-pair), so a batch is
split into contiguous blocks, one per rank, with NO data-path collective; the only collective is
the final gather of the per-pair assignment results.  The reference has no distributed code at all
 — this module is the one parallel strategy the new build adds.  Works on any
``torch.distributed`` backend (NCCL over NVLink on the B200 box; gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(total_pairs, rank, world):
    """Contiguous block [lo, hi) of frame-pairs owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(total_pairs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_pairs(local, total_pairs, group=None):
    """All-gather a per-pair result tensor (first dim = this rank's pairs) into the full batch order.

    Shards may differ by one pair, so shards are padded to the largest size for the collective and
    trimmed afterwards.  Returns a tensor of shape (total_pairs, ...) on every rank."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(total_pairs, r, world) for r in range(world)]
    lo, hi = sizes[rank]
    assert local.shape[0] == hi - lo, (local.shape, lo, hi)
    mx = max(h - l for l, h in sizes)
    pad = local.new_zeros((mx,) + tuple(local.shape[1:]))
    pad[: hi - lo] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[: h - l] for o, (l, h) in zip(out, sizes)], dim=0)
