"""Quick timing probe: C3-shaped network, B members; prints solve time and step statistics."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kinetica_b200 as kb
from kinetica_b200.synthetic import synthetic_crn, synthetic_u0, SEED_BASE

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
R = 5 * S
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
mb = int(sys.argv[3]) if len(sys.argv) > 3 else 0
nt = int(sys.argv[4]) if len(sys.argv) > 4 else 0
rtol = float(sys.argv[5]) if len(sys.argv) > 5 else 1e-8
sd, rd, Ea, A = synthetic_crn(S, R, SEED_BASE + 3)
calc = kb.PrecalculatedArrheniusCalculator(Ea, A, k_max=1e12)
pars = kb.ODESimulationParams(tspan=(0.0, 1.0), u0=synthetic_u0(S), save_interval=0.1, low_k_cutoff="none",
                              solve_chunks=False, abstol=rtol * 1e-2, reltol=rtol)
conds = [kb.ConditionSet({"T": kb.LinearDirectProfile(rate=100.0, X_start=600.0 + 600.0 * b / max(B - 1, 1),
                                                      X_end=700.0 + 600.0 * b / max(B - 1, 1))}, ts_update=1e-2)
         for b in range(B)]
for cs in conds:
    cs.solve_variable_conditions(pars)
t = time.time()
es = kb.EnsembleSolver(sd, rd, calc)
print("symbolic+upload s", time.time() - t, "nnzJ nnzLU nfma", es.nnzJ, es.nnzLU, es.n_fma, flush=True)
es.h.set_tiling(mb, 0)
for rep in range(int(os.environ.get("KB2_REPS", "2"))):
    t = time.time()
    es.prepare(conds, pars, synthetic_u0(S))
    t1 = time.time()
    ms = es.run()
    t2 = time.time()
    out_u, umax, status, stats = es.fetch()
    t3 = time.time()
    print(f"rep{rep}: prepare {t1-t:.3f}s run {ms:.1f} ms (wall {t2-t1:.3f}) fetch {t3-t2:.3f}s  solves/s(device) {B/ms*1e3:.1f}")
print("launch", es.h.get_launch_info())
print("status counts", np.bincount(status), "steps acc mean/min/max", stats[:, 0].mean(), stats[:, 0].min(), stats[:, 0].max(),
      "rej mean", stats[:, 1].mean())
for w, name in enumerate(["arrhenius", "rhs", "jac", "factor", "trisolve"]):
    print(name, "ms", es.h.time_kernel(w, B, 5))
