"""N > 1 path on CPU: world_size-2 gloo.  The ensemble shards over ranks with no exchange during
the solve; each rank integrates its member slice (here with the plain-C oracle standing in for the
GPU) and one all-gather assembles the member-major results."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, B_total, q):
    import torch
    import torch.distributed as dist
    from kinetica_b200.parallel import allgather_members, member_slice
    from oracle import c_oracle as co, kinetica_oracle as ko
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    net = ko.Network(3, [[0], [1], [1, 2]], [[1], [1, 2], [0, 2]], [[1], [2], [1, 1]], [[1], [1, 1], [1, 1]])
    lo, hi = member_slice(B_total, rank, world)
    Ts = [300.0 + 10.0 * b for b in range(lo, hi)]
    A = np.array([0.04, 3e7, 1e4]) / ko.N_A
    Ea = np.array([0.0, 2e3, 1e3])
    out, st, _, _ = co.solve_rodas4(net, A, Ea, None, 1.0, Ts, None, None, [1.0, 0, 0], (0.0, 1.0), np.array([0.0, 1.0]),
                                    nthreads=1)
    fin = torch.from_numpy(out[:, -1, :].copy())
    allfin = allgather_members(fin, B_total)
    status = allgather_members(torch.from_numpy(st.astype(np.int64)), B_total)
    if rank == 0:
        q.put((allfin.numpy(), status.numpy()))
    dist.destroy_process_group()


@pytest.mark.parametrize("B_total", [6, 7])
def test_sharded_ensemble_allgather(built, B_total):
    import torch.multiprocessing as mp
    from kinetica_b200.parallel import member_slice
    from oracle import c_oracle as co, kinetica_oracle as ko
    assert member_slice(7, 0, 2) == (0, 4) and member_slice(7, 1, 2) == (4, 7)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    allfin, status = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    net = ko.Network(3, [[0], [1], [1, 2]], [[1], [1, 2], [0, 2]], [[1], [2], [1, 1]], [[1], [1, 1], [1, 1]])
    Ts = [300.0 + 10.0 * b for b in range(B_total)]
    ref, st, _, _ = co.solve_rodas4(net, np.array([0.04, 3e7, 1e4]) / ko.N_A, np.array([0.0, 2e3, 1e3]), None, 1.0, Ts,
                                    None, None, [1.0, 0, 0], (0.0, 1.0), np.array([0.0, 1.0]), nthreads=1)
    assert allfin.shape == (B_total, 3) and np.array_equal(allfin, ref[:, -1, :]) and np.all(status == 0)
