"""Multi-GPU check (run under torchrun): every rank solves its slice of a small ensemble, the results are
all-gathered by libkinetica_b200.so over NCCL (kb2_allgather_results) and compared with a
torch.distributed all_gather of the same data; prints the device time of repeated gathers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import kinetica_b200 as kb
from kinetica_b200 import parallel
from kinetica_b200.synthetic import synthetic_crn, synthetic_u0, SEED_BASE

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
S, R, Btot = 200, 1000, int(sys.argv[1]) if len(sys.argv) > 1 else 37
sd, rd, Ea, A = synthetic_crn(S, R, SEED_BASE + 77)
calc = kb.PrecalculatedArrheniusCalculator(Ea, A, k_max=1e12)
pars = kb.ODESimulationParams(tspan=(0.0, 0.2), u0=synthetic_u0(S), save_interval=0.1, low_k_cutoff="none", solve_chunks=False)
conds = [kb.ConditionSet({"T": kb.LinearDirectProfile(rate=100.0, X_start=700.0 + 5.0 * b, X_end=800.0 + 5.0 * b)}, ts_update=1e-2)
         for b in range(Btot)]
outs, fin, mx = parallel.solve_network_sharded(kb.B200EnsembleODESolve(pars, conds, calc), sd, rd, rank, world, device=lr)
# reference gather through torch.distributed of each rank's own finals (padded like the shards)
loc, nv = parallel.shard_members(list(range(Btot)), rank, world)
mine = torch.tensor(np.array([o.sol.u[-1] for o in outs] + [outs[-1].sol.u[-1]] * (len(loc) - nv)), device=f"cuda:{lr}")
allf = torch.empty((world * len(loc), S), dtype=torch.float64, device=f"cuda:{lr}")
dist.all_gather_into_tensor(allf, mine)
ref = parallel.unpad_gathered(allf.cpu().numpy(), Btot, world)
ok = np.array_equal(ref, fin) and fin.shape == (Btot, S)
mxl = np.array([o.umax for o in outs])
idx = parallel.member_indices(Btot, rank, world)
ok = ok and np.array_equal(mx[idx], mxl)
print(f"rank {rank}: gathered finals identical to torch all_gather: {ok}; members {idx[:4].tolist()}...", flush=True)
# repeated gathers on a persistent solver: device time
es = kb.EnsembleSolver(sd, rd, calc, device=lr)
parallel.init_comm(es.h, rank, world)
es.solve(loc and [conds[b] for b in loc], pars, pars.u0)
for it in range(4):
    es.h.allgather_results(to_host=False)
    print(f"rank {rank} gather {it}: {es.h.gathered_device()['gather_ms']:.3f} ms", flush=True)
es.close()
dist.destroy_process_group()
assert ok
