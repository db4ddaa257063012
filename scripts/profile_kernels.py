"""One launch of every stand-alone kernel at a workload size, for ncu:
ncu --set full --clock-control none --import-source on -k regex:'k_(rhs|factor|trisolve)' python scripts/profile_kernels.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kinetica_b200 import _lib
from kinetica_b200.synthetic import synthetic_crn, SEED_BASE

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
mb = int(sys.argv[3]) if len(sys.argv) > 3 else 0
which = [int(x) for x in (sys.argv[4] if len(sys.argv) > 4 else "1,5,6,4").split(",")]
R = 5 * S
sd, rd, Ea, A = synthetic_crn(S, R, SEED_BASE + 3)
h = _lib.Handle(0)
h.set_network(S, *rd.flatten())
h.symbolic(4)
h.set_arrhenius(A, Ea, None, 1e12, 1.0)
h.set_tiling(mb, 0)
rng = np.random.default_rng(0)
u = rng.uniform(0, 0.1, (S, B)); k = 10 ** rng.uniform(-3, 5, (R, B))
h.eval_rhs(u, k)
h.factor(u, k, np.full(B, 1e6), want_lu=False)
names = ["arrhenius", "rhs", "jac", "factor", "trisolve", "assemble", "lu_only"]
for w in which:
    print(names[w], h.time_kernel(w, B, 1), "ms", flush=True)
