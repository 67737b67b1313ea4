"""Batch sharding for multi-GPU runs: clips are independent (no cross-clip state in run_classifier; the CMVN window is
intra-clip), so GPU g of G gets the contiguous range [g*B/G, (g+1)*B/G) and there is no collective on the data path
(SURVEY.md §8e).  torch.distributed is only used by callers to gather the 4*L bytes/clip of results."""


def shard_range(n_total: int, world_size: int, rank: int):
    """contiguous [lo, hi) of the batch owned by `rank`; sizes differ by at most one clip"""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(n_total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_results(local, n_total: int, world_size: int, rank: int, group=None):
    """all-gather per-shard result tensors [n_local, L] into the full [n_total, L] tensor (off the hot path)"""
    import torch
    import torch.distributed as dist
    if world_size == 1:
        return local
    sizes = [shard_range(n_total, world_size, r) for r in range(world_size)]
    pad = max(hi - lo for lo, hi in sizes)
    buf = torch.zeros((pad, local.shape[1]), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world_size)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], dim=0)
