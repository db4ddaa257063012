"""GPU parity, kernel level, on a hand-made network with the shapes the synthetic generator never
produces: three distinct reactants in one reaction (third slot of the Jacobian's derivative
table), a collision partner that nets out, a stoichiometry of 4 (general power path), an inert
species.  Same checks as test_gpu_kernels.py::test_rhs_and_jacobian."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_rhs_and_jacobian_edge_network(built):
    import kinetica_b200 as kb
    from kinetica_b200 import _lib
    from oracle import c_oracle as co, kinetica_oracle as ko
    #        A+B+C -> D      A+M -> B+M    4A -> E     D -> A+B     2E -> 3A + C
    reacs = [[0, 1, 2],      [0, 6],       [0],        [3],         [4]]
    prods = [[3],            [1, 6],       [4],        [0, 1],      [0, 2]]
    sr    = [[1, 1, 1],      [1, 1],       [4],        [1],         [2]]
    sp    = [[1],            [1, 1],       [1],        [1, 1],      [3, 1]]
    rd = kb.RxData(reacs, prods, sr, sp)
    S = 7                                            # species 5 is inert, species 6 = M
    h = _lib.Handle(0)
    try:
        h.set_network(S, *rd.flatten())
        h.symbolic(0)
        net = ko.Network(S, rd.id_reacs, rd.id_prods, rd.stoic_reacs, rd.stoic_prods)
        assert h.get_gather_tables()["jslots"] == 3
        colptr, rowval = net.pattern_csc()
        cp, rv = h.get_pattern()
        assert np.array_equal(cp, colptr) and np.array_equal(rv, rowval)
        rng = np.random.default_rng(5)
        for B in (1, 6, 37):
            u = rng.uniform(0, 1, (S, B)) * (rng.random((S, B)) > 0.2)        # exact zeros included
            k = 10 ** rng.uniform(-3, 4, (net.R, B))
            du = h.eval_rhs(u, k)
            J = h.eval_jac(u, k)
            for b in range(B):
                ref = co.rhs(net, u[:, b], k[:, b])
                assert np.max(np.abs(du[:, b] - ref) / (np.abs(ref) + 1e-9 * np.max(np.abs(ref)) + 1e-300)) < 1e-9
                Jd = net.jac_dense(u[:, b], k[:, b])
                Jref = np.array([Jd[rowval[p], l] for l in range(S) for p in range(colptr[l], colptr[l + 1])])
                tol = 1e-11 * (np.abs(Jref) + 1e-6 * np.max(np.abs(Jref))) + 1e-300
                assert np.all(np.abs(J[:, b] - Jref) <= tol)
    finally:
        h.close()
