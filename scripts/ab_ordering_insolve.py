"""In-solve A/B of two elimination orderings on the same box: a C3 ensemble (4096 members) runs
`maxiters` attempted steps per member from t = 0 with each ordering; prints the device time per
round (all phase kernels) and the sampled phase timings.  Usage: python scripts/ab_ordering_insolve.py [maxiters] [orderings...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kinetica_b200 as kb
from kinetica_b200.synthetic import synthetic_crn, synthetic_u0, SEED_BASE
S, R, B, tf = 1000, 5000, 4096, 0.01
maxit = int(sys.argv[1]) if len(sys.argv) > 1 else 64
orderings = [int(a) for a in sys.argv[2:]] or [4, 3]
sd, rd, Ea, A = synthetic_crn(S, R, SEED_BASE + 3)
calc = kb.PrecalculatedArrheniusCalculator(Ea, A, k_max=1e12)
pars = kb.ODESimulationParams(tspan=(0.0, tf), u0=synthetic_u0(S), save_interval=tf, low_k_cutoff="none", solve_chunks=False, maxiters=maxit)
conds = [kb.ConditionSet({"T": kb.LinearDirectProfile(rate=100.0, X_start=600.0 + 600.0 * b / (B - 1), X_end=700.0 + 600.0 * b / (B - 1))},
                         ts_update=1e-2) for b in range(B)]
for cs in conds:
    cs.solve_variable_conditions(pars)
for o in orderings:
    es = kb.EnsembleSolver(sd, rd, calc, ordering=o)
    es.prepare(conds, pars, synthetic_u0(S))
    ms = es.run()
    ph, rounds = es.h.get_phase_times()
    st = es.h.get_plan_stats()
    print("ordering %d (in use %d) padded %d: %.1f ms, %d rounds, %.3f ms/round | phases %s" %
          (o, st["ordering"], st["padded"], ms, rounds, ms / max(rounds, 1), {k: round(v["ms"], 3) for k, v in ph.items()}), flush=True)
    es.close()
