"""Multi-GPU plumbing: the ensemble shards over ranks with no exchange during the solve; the only
collective is one all-gather of member-major final concentrations / per-species maxima / status at
the end.  One process per GPU, torch.distributed (NCCL on GPUs, gloo in CPU tests).
"""
from __future__ import annotations


def member_slice(B_total: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of the (temperature-sorted) member axis owned by `rank`;
    the first B_total % world ranks take one extra member."""
    base, rem = divmod(B_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allgather_members(local, B_total: int, group=None):
    """All-gather member-major tensors `local[B_loc, ...]` into `[B_total, ...]` in rank order.
    Ragged shards (B_total % world != 0) are padded to the largest shard for the collective."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = [member_slice(B_total, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = local
    if local.shape[0] < mx:
        pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[:local.shape[0]] = local
    out = torch.empty((world * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    if all(hi - lo == mx for lo, hi in sizes):
        return out
    return torch.cat([out[r * mx: r * mx + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], dim=0)
