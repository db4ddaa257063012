"""Stand-alone timing of the window LU and the triangular sweeps on a C3-shaped ensemble for the
candidate orderings of `auto` (kb2_symbolic).  Usage: python scripts/time_orderings.py [S] [B] [config id]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kinetica_b200 import _lib
from kinetica_b200.synthetic import synthetic_crn, SEED_BASE
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
cid = int(sys.argv[3]) if len(sys.argv) > 3 else 3
R = 5 * S
sd, rd, Ea, A = synthetic_crn(S, R, SEED_BASE + cid)
rng = np.random.default_rng(1)
u = rng.uniform(0, 1e-2, (S, B)); k = 10 ** rng.uniform(-3, 3, (R, B)); hg = np.full(B, 1e4)
names = {3: "natural, hubs last", 5: "RCM", 6: "Sloan 1:2", 7: "Sloan 2:1", 4: "auto"}
for ordering in (3, 5, 6, 7, 4):
    h = _lib.Handle(0)
    h.set_network(S, *rd.flatten())
    nnzJ, nnzLU, nfma = h.symbolic(ordering)
    ps, fp = h.get_plan_stats(), h.get_front_plan()
    h.factor(u, k, hg, want_lu=False)
    t_lu = h.time_kernel(7, B, 5)
    t_tri = h.time_kernel(4, B, 5)
    print("%-20s | nnzLU %6d padded %6d fma_pad %8d window %3dx%3d fronts %4d | window LU %.3f ms  trisolve %.3f ms  (LU + 6 sweeps %.3f ms)"
          % (names[ordering], nnzLU, ps["padded"], ps["fma_padded"], fp["Wr"], fp["Wc"], fp["NF"], t_lu, t_tri, t_lu + 6 * t_tri), flush=True)
    h.close()
