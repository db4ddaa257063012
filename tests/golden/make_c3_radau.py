"""Regenerates tests/golden/c3_radau.npz: trajectories of two members of the C3 sweep (S = 1000 /
R = 5000 synthetic network, bench seed; members 4 and 7 of the 8-member sweep 600..1200 K used by
tests/test_gpu_baseline_configs.py::test_c3_network_at_bench_tolerances) from the INDEPENDENT
integrator of the oracle — scipy Radau, rtol 1e-8 / atol 1e-12, analytic sparse Jacobian, restarted at
every one of the 101 rate updates (zero-order hold) — at the 11 save points of tspan (0, 1).
About ten minutes per member on one core, which is why the vectors are committed instead of being
recomputed inside the GPU test.  Run in the authoring container:

    python tests/golden/make_c3_radau.py
"""
import os
import sys
from concurrent.futures import ProcessPoolExecutor

for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):      # the vectors are short: BLAS threads only spin
    os.environ.setdefault(_v, "1")

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

S, R, B = 1000, 5000, 8
MEMBERS = (4, 7)
SAVE_T = np.arange(11) / 10.0


def one(b):
    from kinetica_b200.synthetic import synthetic_crn, synthetic_u0, SEED_BASE
    from oracle import kinetica_oracle as ko
    sd, rd, Ea, A = synthetic_crn(S, R, SEED_BASE + 3)
    net = ko.Network(S, rd.id_reacs, rd.id_prods, rd.stoic_reacs, rd.stoic_prods)
    T0 = 600.0 + 600.0 * b / (B - 1)
    ts = ko.create_savepoints(0.0, 1.0, 1e-2)
    calc = ko.PrecalculatedArrheniusCalculator(Ea, A, k_max=1e12)
    ktab = np.array([calc(T0 + 100.0 * min(t, 1.0)) for t in ts])
    return ko.solve_trajectory(net, synthetic_u0(S), ktab, ts, (0.0, 1.0), SAVE_T, k_init=calc(T0), rtol=1e-8, atol=1e-12)


if __name__ == "__main__":
    with ProcessPoolExecutor(len(MEMBERS)) as ex:
        res = list(ex.map(one, MEMBERS))
    np.savez_compressed(os.path.join(HERE, "c3_radau.npz"), members=np.array(MEMBERS), save_t=SAVE_T,
                        T0=np.array([600.0 + 600.0 * b / (B - 1) for b in MEMBERS]), u=np.array(res))
    print("wrote c3_radau.npz", np.array(res).shape)
