"""N > 1 path on CPU: world_size-2 gloo.  The ensemble shards over ranks with no exchange during
the solve (kinetica_b200.parallel); the NCCL unique id travels from rank 0 through
torch.distributed, every rank integrates its (padded) member share (strided over the member axis) — here with the plain-C oracle
standing in for the GPU — and the rank-major gathered array is un-padded into member order."""
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


class _FakeHandle:
    """Records what parallel.init_comm hands to the C ABI (no GPU here)."""
    def __init__(self, rank):
        self.rank, self.calls = rank, []

    def comm_unique_id(self):
        return bytes([7 + self.rank]) * 128           # only rank 0's id may travel

    def comm_init_rank(self, nranks, rank, uid):
        self.calls.append((nranks, rank, uid))


def _worker(rank, world, port, B_total, q):
    import torch
    import torch.distributed as dist
    from kinetica_b200.parallel import init_comm, member_indices, shard_members
    from oracle import c_oracle as co, kinetica_oracle as ko
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    h = _FakeHandle(rank)
    init_comm(h, rank, world)
    assert h.calls == [(world, rank, bytes([7]) * 128)]
    net = ko.Network(3, [[0], [1], [1, 2]], [[1], [1, 2], [0, 2]], [[1], [2], [1, 1]], [[1], [1, 1], [1, 1]])
    Ts_all = [300.0 + 10.0 * b for b in range(B_total)]
    Ts, nvalid = shard_members(Ts_all, rank, world)
    idx = member_indices(B_total, rank, world)
    assert nvalid == len(idx) and len(Ts) == -(-B_total // world) and Ts[:nvalid] == [Ts_all[b] for b in idx]
    A = np.array([0.04, 3e7, 1e4]) / ko.N_A
    Ea = np.array([0.0, 2e3, 1e3])
    out, st, _, _ = co.solve_rodas4(net, A, Ea, None, 1.0, Ts, None, None, [1.0, 0, 0], (0.0, 1.0), np.array([0.0, 1.0]),
                                    nthreads=1)
    # what kb2_allgather_results does over NCCL: equal-sized member-major blocks, rank-major
    fin = torch.from_numpy(out[:, -1, :].copy())
    gathered = torch.empty((world * fin.shape[0], fin.shape[1]), dtype=fin.dtype)
    dist.all_gather_into_tensor(gathered, fin)
    if rank == 0:
        q.put(gathered.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("B_total", [6, 7])
def test_sharded_ensemble_allgather(built, B_total):
    import torch.multiprocessing as mp
    from kinetica_b200.parallel import member_indices, unpad_gathered
    from oracle import c_oracle as co, kinetica_oracle as ko
    assert member_indices(7, 0, 2).tolist() == [0, 2, 4, 6] and member_indices(7, 1, 2).tolist() == [1, 3, 5]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    allfin = unpad_gathered(gathered, B_total, 2)
    net = ko.Network(3, [[0], [1], [1, 2]], [[1], [1, 2], [0, 2]], [[1], [2], [1, 1]], [[1], [1, 1], [1, 1]])
    Ts = [300.0 + 10.0 * b for b in range(B_total)]
    ref, st, _, _ = co.solve_rodas4(net, np.array([0.04, 3e7, 1e4]) / ko.N_A, np.array([0.0, 2e3, 1e3]), None, 1.0, Ts,
                                    None, None, [1.0, 0, 0], (0.0, 1.0), np.array([0.0, 1.0]), nthreads=1)
    assert allfin.shape == (B_total, 3) and np.array_equal(allfin, ref[:, -1, :])


def test_shard_helpers():
    from kinetica_b200.parallel import member_indices, shard_members, unpad_gathered
    for B, world in ((10, 4), (8, 8), (65536, 8), (5, 2)):
        per = -(-B // world)
        cover = []
        blocks = []
        for r in range(world):
            loc, nv = shard_members(list(range(B)), r, world)
            assert len(loc) == per and loc[:nv] == member_indices(B, r, world).tolist()
            cover += loc[:nv]
            blocks.append(np.array(loc)[:, None])
        assert sorted(cover) == list(range(B))
        assert np.array_equal(unpad_gathered(np.concatenate(blocks), B, world)[:, 0], np.arange(B))
    with pytest.raises(ValueError):
        shard_members([1], 1, 2)
