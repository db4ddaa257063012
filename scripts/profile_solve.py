"""Short solve for ncu / phase timing: C3 network, B members, a few stops."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kinetica_b200 as kb
from kinetica_b200.synthetic import synthetic_crn, synthetic_u0, SEED_BASE
S = int(os.environ.get('KB2_S', '1000')); R = 5 * S
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
tf = float(sys.argv[2]) if len(sys.argv) > 2 else 0.02
maxit = int(sys.argv[3]) if len(sys.argv) > 3 else 100000
sd, rd, Ea, A = synthetic_crn(S, R, SEED_BASE + int(os.environ.get('KB2_CID', '3')))
calc = kb.PrecalculatedArrheniusCalculator(Ea, A, k_max=1e12)
pars = kb.ODESimulationParams(tspan=(0.0, tf), u0=synthetic_u0(S), save_interval=tf / 2, low_k_cutoff="none",
                              solve_chunks=False, abstol=float(os.environ.get('KB2_ATOL', '1e-10')), reltol=float(os.environ.get('KB2_RTOL', '1e-8')), maxiters=maxit)
conds = [kb.ConditionSet({"T": kb.LinearDirectProfile(rate=100.0, X_start=600.0 + 600.0 * b / (B - 1), X_end=700.0 + 600.0 * b / (B - 1))},
                         ts_update=1e-2) for b in range(B)]
for cs in conds:
    cs.solve_variable_conditions(pars)
es = kb.EnsembleSolver(sd, rd, calc)
es.h.set_tiling(int(os.environ.get('KB2_MBSET', '0')), 0)
for rep in range(2):
    es.prepare(conds, pars, synthetic_u0(S))
    ms = es.run()
    out_u, umax, status, stats = es.fetch()
    ph, rounds = es.h.get_phase_times()
    print(f"rep{rep} run {ms:.1f} ms; attempts mean {stats[:,2].mean():.1f} max {stats[:,2].max()}; ok {np.sum(status==0)}/{B}; rounds {rounds}; ms/round {ms/max(rounds,1):.3f}")
    print("   phases:", {k: round(v["ms"], 3) for k, v in ph.items()}, "per round:", round(ph["jacobian"]["ms"] + ph["lu"]["ms"] + 6 * (ph["stage_rhs"]["ms"] + ph["stage_sweeps"]["ms"]) + ph["step_end"]["ms"], 3))
