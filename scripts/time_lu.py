"""Stand-alone timing of the factorisation kernels on a C3-shaped ensemble."""
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
h = _lib.Handle(0)
h.set_network(S, *rd.flatten())
print("symbolic", h.symbolic(4), h.get_plan_stats())
rng = np.random.default_rng(1)
u = rng.uniform(0, 1e-2, (S, B)); k = 10 ** rng.uniform(-3, 3, (R, B)); hg = np.full(B, 1e4)
h.factor(u, k, hg, want_lu=False)
names = {1: "rhs", 2: "jac values", 3: "factor (as solve)", 4: "trisolve", 5: "panel assembly", 6: "panel LU", 7: "window LU", 8: "panel asm+LU"}
for w in (1, 2, 7, 3, 4, 5, 6, 8):
    try:
        print("%-20s %.3f ms" % (names[w], h.time_kernel(w, B, 5)), flush=True)
    except Exception as e:
        print(names[w], "failed:", e)
