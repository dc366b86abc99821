"""Batch-sharded inference over one process per GPU.

Every sample is independent (thresholds, masks and active lists are per sample; weights are
replicated), so the batch is partitioned contiguously over ranks, each rank runs encoder features ->
decoder locally, and exactly ONE collective moves data: an all-gather of the full-resolution
``("disp", 0)`` tensor.  The reference has no distributed code at all (single process, single device:
KITTI/trainer.py:45, evaluate_depth.py:116); this is the one parallelism the hot path admits.
"""
import torch
import torch.distributed as dist


def shard_bounds(n, world, rank):
    """Contiguous [lo, hi) slice of a batch of n for `rank`; sizes differ by at most one."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_features(feats, world, rank):
    """Slice every feature map of a global batch to this rank's shard (views, no copy)."""
    lo, hi = shard_bounds(feats[0].shape[0], world, rank)
    return [f[lo:hi] for f in feats]


def all_gather_batch(local, n_global, group=None):
    """All-gather a batch-sharded tensor (shards from shard_bounds) into the global (n_global, ...) tensor.

    One collective.  Ragged shards (n_global % world != 0) are padded to the largest shard and trimmed.
    Uses all_gather_into_tensor on NCCL (a single NVLink/NVSwitch all-gather into the output buffer) and
    the list form elsewhere (gloo in the CPU tests).
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    rank = dist.get_rank(group)
    per = -(-n_global // world)
    lo, hi = shard_bounds(n_global, world, rank)
    assert local.shape[0] == hi - lo, (local.shape, lo, hi)
    if local.shape[0] != per:
        pad = torch.zeros((per - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], 0)
    local = local.contiguous()
    gathered = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(gathered, local, group=group)
    else:
        dist.all_gather(list(gathered.chunk(world, 0)), local, group=group)
    if n_global % world == 0:
        return gathered
    parts = []
    for r in range(world):
        a, b = shard_bounds(n_global, world, r)
        parts.append(gathered[r * per:r * per + (b - a)])
    return torch.cat(parts, 0)


def sharded_decode(decoder, local_feats, n_global, *args, gather_key=("disp", 0), group=None, **kwargs):
    """Run `decoder` on this rank's shard and all-gather its full-resolution output.

    Returns (local_outputs, global_disp)."""
    out = decoder(local_feats, *args, **kwargs)
    return out, all_gather_batch(out[gather_key], n_global, group=group)
