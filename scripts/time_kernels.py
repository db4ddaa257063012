"""Stand-alone kernel timings at a workload size for several tilings."""
import sys, os, time
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
print("symbolic", h.symbolic(4), h.get_plan_stats(), flush=True)
h.set_arrhenius(A, Ea, None, 1e12, 1.0)
rng = np.random.default_rng(0)
u = rng.uniform(0, 0.1, (S, B)); k = 10 ** rng.uniform(-3, 5, (R, B))
st = h.get_plan_stats()
peak = 6546.2
bytes_ = {"arrhenius": 8 * (R + 1), "rhs": 8 * (R + 2 * S), "jac": 8 * (R + S + h.nnzJ),
          "factor": 8 * (R + S + h.nnzJ + 2 * h.nnzLU), "trisolve": 8 * (h.nnzLU + 2 * S)}
only = os.environ.get("KB2_ONLY")
combos = [int(x) for x in os.environ.get("KB2_MB", "4,2,1").split(",")]
for mb in combos:
    var = 0
    h.set_tiling(mb, 0)
    h.eval_rhs(u, k)
    h.factor(u, k, np.full(B, 1e6), want_lu=False)
    out = []
    bytes_["assemble"] = 8 * (R + S + h.nnzJ + h.nnzLU); bytes_["lu_only"] = 8 * 2 * h.nnzLU
    t0 = time.time()
    for w, nm in enumerate(["arrhenius", "rhs", "jac", "factor", "trisolve", "assemble", "lu_only"]):
        if only and nm not in only.split(","):
            continue
        ms = h.time_kernel(w, B, 2 if only else 5)
        gb = bytes_[nm] * B / ms / 1e6
        extra = f" {2 * h.n_fma * B / ms / 1e9:.2f}TF(exact) {2 * st['fma_padded'] * B / ms / 1e9:.2f}TF(padded)" if nm == "factor" else ""
        out.append(f"{nm} {ms:.3f}ms {gb:.0f}GB/s({gb / peak * 100:.1f}%)" + extra)
    print(f"mb={mb}: " + " | ".join(out), flush=True)
