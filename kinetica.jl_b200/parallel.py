"""Multi-GPU front door: the ensemble shards over ranks with no exchange during the solve; the one
collective is an all-gather of member-major final concentrations and per-species maxima at the end
(SURVEY.md §8e), done inside libkinetica_b200.so over NCCL (kb2_comm_* / kb2_allgather_results).
One process per GPU; torch.distributed (any backend) is only used to hand rank 0's NCCL unique id
to the other ranks.
"""
from __future__ import annotations

import copy
from typing import Sequence

import numpy as np


def member_indices(B_total: int, rank: int, world: int) -> np.ndarray:
    """Members owned by `rank`: b = rank, rank + world, rank + 2*world, ...  Strided, not contiguous:
    the step count of a member depends on its condition (C3: 2204-2804 accepted steps across the
    temperature sweep), and every rank waits for the slowest one at the gather, so each rank takes an
    even sample of the (sorted) member axis."""
    return np.arange(rank, B_total, world)


def shard_members(items: Sequence, rank: int, world: int):
    """This rank's members of a per-member list, padded with copies of its last entry to the largest
    shard (the all-gather needs the same member count on every rank) -> (padded list, valid count)."""
    B = len(items)
    idx = member_indices(B, rank, world)
    per = -(-B // world)
    loc = [items[int(b)] for b in idx]
    if not loc:
        raise ValueError("more ranks than ensemble members")
    return loc + [loc[-1]] * (per - len(loc)), len(idx)


def unpad_gathered(arr: np.ndarray, B_total: int, world: int):
    """[world * per, ...] rank-major gathered array -> [B_total, ...] in member order, padding dropped."""
    per = -(-B_total // world)
    b = np.arange(B_total)
    return arr[(b % world) * per + b // world]


def exchange_unique_id(make_id, rank: int, group=None) -> bytes:
    """Rank 0 creates the NCCL unique id (`make_id()`), every rank returns it."""
    import torch.distributed as dist
    box = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    return box[0]


def init_comm(handle, rank: int, world: int, group=None):
    """Attach an NCCL communicator to a kb2 handle (one handle = one GPU = one rank)."""
    uid = exchange_unique_id(handle.comm_unique_id, rank, group)
    handle.comm_init_rank(world, rank, uid)


def solve_network_sharded(method, sd, rd, rank: int, world: int, device: int = 0, group=None, solver=None):
    """`solve_network(B200EnsembleODESolve(...))` over `world` GPUs: every rank integrates its
    strided share of `method.conditions` and all ranks receive the gathered summaries.

    Returns (outs, final_all, umax_all): this rank's `ODESolveOutput`s, and `[B_total, S]` arrays of
    final concentrations and per-species maxima of the whole ensemble (rank-major = member order).
    """
    from .solve import B200EnsembleODESolve, EnsembleSolver, solve_network
    if not isinstance(method, B200EnsembleODESolve):
        raise TypeError("solve_network_sharded needs a B200EnsembleODESolve")
    B_total = len(method.conditions)
    conds, nvalid = shard_members(method.conditions, rank, world)
    local = B200EnsembleODESolve(method.pars, conds, method.calculator, method.filter)
    keep = {}
    outs = solve_network(local, sd, rd, device=device, solver=solver, _keep_solver=keep)
    es = keep["solver"]
    try:
        if world > 1:
            init_comm(es.h, rank, world, group)
            fin, mx = es.h.allgather_results(to_host=True)
        else:
            fin = np.array([o.sol.u[-1] for o in outs])
            mx = np.array([o.umax for o in outs])
    finally:
        if solver is None:
            es.close()
    return outs[:nvalid], unpad_gathered(fin, B_total, world), unpad_gathered(mx, B_total, world)
