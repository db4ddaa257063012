"""One A/B point of the window LU: loads the library named by KB2_LIB (default: the product build),
checks the window LU against the block-plan LU bit for bit on the C3 network (8 members), then times
the window LU and the triangular sweeps on 4096 members for the given ordering with and without
look-ahead for every front.  Usage: KB2_LIB=... python scripts/time_variant.py [ordering] [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kinetica_b200 import _lib
from kinetica_b200.synthetic import synthetic_crn, SEED_BASE
ordering = int(sys.argv[1]) if len(sys.argv) > 1 else 6
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
S, R = 1000, 5000
sd, rd, Ea, A = synthetic_crn(S, R, SEED_BASE + 3)
rng = np.random.default_rng(1)
tag = os.path.basename(os.environ.get("KB2_LIB", "product"))
for la_all in (0, 1):
    os.environ["KB2_LA_ALL"] = str(la_all)
    # bitwise check on a few members
    Bs = 8
    u = rng.uniform(0, 1, (S, Bs)); k = 10 ** rng.uniform(-3, 3, (R, Bs)); hg = 10 ** rng.uniform(1, 4, Bs)
    res = {}
    for mode in ("window", "panel"):
        os.environ["KB2_LU"] = mode
        h = _lib.Handle(0); h.set_network(S, *rd.flatten()); h.symbolic(ordering); h.set_tiling(4)
        res[mode] = h.factor(u, k, hg)
        h.close()
    os.environ.pop("KB2_LU")
    same = np.array_equal(res["window"], res["panel"])
    u = rng.uniform(0, 1e-2, (S, B)); k = 10 ** rng.uniform(-3, 3, (R, B)); hg = np.full(B, 1e4)
    h = _lib.Handle(0); h.set_network(S, *rd.flatten()); h.symbolic(ordering)
    h.factor(u, k, hg, want_lu=False)
    t_lu = min(h.time_kernel(7, B, 5) for _ in range(2))
    t_tri = h.time_kernel(4, B, 5)
    h.close()
    print("%-20s ordering %d la_all %d | bitwise %s | window LU %.3f ms  trisolve %.3f ms" % (tag, ordering, la_all, "OK" if same else "FAIL", t_lu, t_tri), flush=True)
