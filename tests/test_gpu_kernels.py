"""GPU parity tests, kernel level: every kernel-level C-ABI entry point against the CPU oracle
on the same seeded inputs."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def _golden_arrhenius():
    d = json.load(open(os.path.join(HERE, "golden", "arrhenius_params.json")))
    return (np.array([float.fromhex(x) for x in d["Ea"]]), np.array([float.fromhex(x) for x in d["A"]]))


def _ulp_diff(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.spacing(np.maximum(np.abs(a), np.abs(b))), 5e-324)


@pytest.fixture(scope="module")
def small(built):
    from kinetica_b200 import _lib
    from kinetica_b200.synthetic import synthetic_crn, SEED_BASE
    from oracle import kinetica_oracle as ko
    S, R = 96, 400
    sd, rd, Ea, A = synthetic_crn(S, R, SEED_BASE + 11)
    # a few unusual stoichiometries: 3A -> B, A + 2B -> C
    rd.id_reacs[0], rd.stoic_reacs[0] = [5], [3]
    rd.id_reacs[1], rd.stoic_reacs[1] = [6, 7], [1, 2]
    h = _lib.Handle(0)
    h.set_network(S, *rd.flatten())
    h.symbolic(0)
    net = ko.Network(S, rd.id_reacs, rd.id_prods, rd.stoic_reacs, rd.stoic_prods)
    yield h, net, rd, Ea, A
    h.close()


def test_arrhenius_shipped_fixture(built):
    """k(T) on the reference's shipped Ea/A (examples/getting_started/arrhenius_params.bson) with
    k_max = 1e12 (docs/src/getting-started.md:152): <= 2 ulp against the restatement of
    calculator.jl:223-232, including exp underflow and Ea = 0."""
    from kinetica_b200 import _lib
    from kinetica_b200.synthetic import getting_started_standin
    from oracle import kinetica_oracle as ko
    Ea, A = _golden_arrhenius()
    sd, rd = getting_started_standin()
    h = _lib.Handle(0)
    h.set_network(sd.n, *rd.flatten())
    h.symbolic(0)
    T = np.array([300.0, 500.0, 850.0, 1200.0, 2000.0, 50.0])
    for k_max in (1e12, None):
        h.set_arrhenius(A, Ea, None, k_max, 1.0)
        k = h.eval_k(T)
        calc = ko.PrecalculatedArrheniusCalculator(Ea, A, k_max=k_max)
        for b, t in enumerate(T):
            ref = calc(t)
            assert np.all(np.isfinite(k[:, b]))
            assert np.max(_ulp_diff(k[:, b], ref)[ref > 1e-300]) <= 2.0
            assert np.all(k[ref == 0, b] == 0)
    # SURVEY §8c spot values (restatement-derived)
    h.set_arrhenius(A, Ea, None, 1e12, 1.0)
    k = h.eval_k(np.array([500.0, 850.0, 1200.0]))
    assert abs(k[0, 0] / 1.1476254735240006e-28 - 1) < 1e-13
    assert abs(k[1, 0] / 9.9999898092210303e+11 - 1) < 1e-13
    assert abs(k[4, 1] / 1.2031111957791838e+05 - 1) < 1e-13
    assert abs(k[0, 2] / 4.187600263086772e+07 - 1) < 1e-13
    h.close()


def test_arrhenius_tn_and_units(small):
    h, net, rd, Ea, A = small
    rng = np.random.default_rng(1)
    n = rng.uniform(-1, 2, len(A))
    h.set_arrhenius(A, Ea, n, 1e12, 60.0)
    T = np.linspace(400, 1500, 37)
    k = h.eval_k(T)
    for b, t in enumerate(T):
        kr = A * np.exp(-Ea / (8.314462618 * t)) * t ** n * 6.02214076e23 * 60.0
        ref = 1.0 / (1.0 / 1e12 + 1.0 / kr)
        assert np.max(np.abs(k[:, b] - ref) / ref) < 1e-13


def test_rhs_and_jacobian(small):
    from oracle import c_oracle as co
    h, net, rd, Ea, A = small
    rng = np.random.default_rng(2)
    B = 37                         # ragged: not a multiple of the tile width
    u = rng.uniform(0, 1, (net.S, B)) * (rng.random((net.S, B)) > 0.2)
    k = 10 ** rng.uniform(-6, 9, (net.R, B))
    du = h.eval_rhs(u, k)
    J = h.eval_jac(u, k)
    colptr, rowval = net.pattern_csc()
    cp, rv = h.get_pattern()
    assert np.array_equal(cp, colptr) and np.array_equal(rv, rowval)      # bit-exact pattern
    for b in range(B):
        ref = co.rhs(net, u[:, b], k[:, b])
        scale = np.abs(net.rhs(np.abs(u[:, b]), k[:, b])) + 1e-300
        assert np.max(np.abs(du[:, b] - ref) / (np.abs(ref) + 1e-9 * np.max(np.abs(ref)))) < 1e-9
        Jd = net.jac_dense(u[:, b], k[:, b])
        Jref = np.array([Jd[rowval[p], l] for l in range(net.S) for p in range(colptr[l], colptr[l + 1])])
        Jc = co.jac_csc(net, u[:, b], k[:, b])
        tol = 1e-11 * (np.abs(Jref) + 1e-6 * np.max(np.abs(Jref)))
        assert np.all(np.abs(J[:, b] - Jref) <= tol)
        assert np.all(np.abs(Jc - Jref) <= tol)


def test_rhs_deterministic(small):
    h, net, rd, Ea, A = small
    rng = np.random.default_rng(3)
    u = rng.uniform(0, 1, (net.S, 64)); k = 10 ** rng.uniform(-3, 6, (net.R, 64))
    a = h.eval_rhs(u, k); b = h.eval_rhs(u, k)
    assert np.array_equal(a, b)            # gather CSR, no atomics: bit-identical run to run


@pytest.mark.parametrize("S,R,ordering,B,mb", [(96, 400, 0, 19, 0), (420, 2100, 3, 9, 1), (420, 2100, 0, 16, 4), (30, 60, 1, 3, 0),
                                               (420, 2100, 3, 10, 2), (200, 1000, 3, 37, 4), (96, 400, 4, 8, 2), (420, 2100, 3, 19, 4), (96, 400, 0, 16, 4)])
def test_factor_and_trisolve_panels(built, S, R, ordering, B, mb):
    """Panel LU + panel triangular solves on networks whose hub rows span several column chunks
    (S = 420: widest panel > 3 chunks), for every ordering mode and ragged member counts."""
    from kinetica_b200 import _lib
    from kinetica_b200.synthetic import synthetic_crn, SEED_BASE
    from oracle import kinetica_oracle as ko
    sd, rd, Ea, A = synthetic_crn(S, R, SEED_BASE + 50 + S)
    h = _lib.Handle(0)
    h.set_network(S, *rd.flatten())
    h.symbolic(ordering)
    h.set_tiling(mb)        # members per warp tile: 0 auto (1 for these small ensembles), 1, 2, 4
    st = h.get_plan_stats()
    if S == 420:
        assert st["max_width"] > 96 and st["units"] > st["panels"]
    net = ko.Network(S, rd.id_reacs, rd.id_prods, rd.stoic_reacs, rd.stoic_prods)
    _factor_trisolve_check(h, net, B, seed=S)
    h.close()


def test_factor_and_trisolve(small):
    h, net, rd, Ea, A = small
    _factor_trisolve_check(h, net, 19, seed=4)


def _factor_trisolve_check(h, net, B, seed):
    rng = np.random.default_rng(seed)
    u = rng.uniform(0, 1, (net.S, B))
    k = 10 ** rng.uniform(-3, 3, (net.R, B))
    hg = 10 ** rng.uniform(1, 4, B)
    lu = h.factor(u, k, hg)
    rhs = rng.normal(size=(net.S, B))
    x = h.trisolve(rhs)
    rowptr, colidx, diagpos = h.get_lu_pattern()
    perm = h.get_ordering()
    for b in range(B):
        W = np.eye(net.S) * hg[b] - net.jac_dense(u[:, b], k[:, b])
        xr = np.linalg.solve(W, rhs[:, b])
        assert np.max(np.abs(x[:, b] - xr)) <= 1e-9 * np.max(np.abs(xr))
        # L'*U' (Crout form: pivots on L', unit diagonal on U') reproduces the permuted W on the
        # pattern and nothing outside it
        L = np.zeros((net.S, net.S)); U = np.eye(net.S)
        for i in range(net.S):
            for p in range(rowptr[i], rowptr[i + 1]):
                j = colidx[p]
                if j <= i:
                    L[i, j] = lu[p, b]
                else:
                    U[i, j] = lu[p, b]
        Wp = W[np.ix_(perm, perm)]
        assert np.max(np.abs(L @ U - Wp)) <= 1e-10 * np.max(np.abs(Wp))


@pytest.mark.parametrize("S,R,ordering,B,mb", [(96, 400, 0, 19, 0), (420, 2100, 3, 10, 2), (200, 1000, 3, 37, 4), (420, 2100, 0, 16, 4),
                                               (1000, 5000, 4, 8, 4), (1000, 5000, 3, 8, 4), (1000, 5000, 5, 12, 4), (1000, 5000, 7, 6, 2),
                                               (200, 1000, 7, 37, 4), (200, 1000, 6, 9, 2), (200, 1000, 5, 5, 1), (420, 2100, 7, 16, 4),
                                               (96, 400, 6, 19, 4), (64, 256, 4, 8, 4)])
def test_window_lu_matches_block_plan_lu_bitwise(built, monkeypatch, S, R, ordering, B, mb):
    """The window LU (right-looking, active submatrix in shared memory, built from the compact
    Jacobian values) and the block-plan LU (left-looking over the assembled padded storage) apply the
    same updates to every entry in the same order: factors and solutions must be identical bits."""
    from kinetica_b200 import _lib
    from kinetica_b200.synthetic import synthetic_crn, SEED_BASE
    sd, rd, Ea, A = synthetic_crn(S, R, SEED_BASE + (3 if S == 1000 else 50 + S))
    rng = np.random.default_rng(S + B)
    u = rng.uniform(0, 1, (S, B)); k = 10 ** rng.uniform(-3, 3, (R, B)); hg = 10 ** rng.uniform(1, 4, B)
    rhs = rng.normal(size=(S, B))
    res = {}
    for mode in ("window", "panel"):
        monkeypatch.setenv("KB2_LU", mode)
        h = _lib.Handle(0)
        h.set_network(S, *rd.flatten())
        h.symbolic(ordering)
        h.set_tiling(mb)
        res[mode] = (h.factor(u, k, hg), h.trisolve(rhs))
        h.close()
    assert np.array_equal(res["window"][0], res["panel"][0])
    assert np.array_equal(res["window"][1], res["panel"][1])


def test_profiles_on_device(built):
    """Device X_b(t) against the reference's profile known answers (test/Main/conditions.jl) and
    the host closed forms."""
    import kinetica_b200 as kb
    from kinetica_b200 import _lib
    profs = [
        kb.StaticConditionProfile(10.0),
        kb.NullDirectProfile(X_start=300.0, t_end=10.0),
        kb.LinearDirectProfile(rate=50.0, X_start=300.0, X_end=500.0),
        kb.LinearGradientProfile(rate=50.0, X_start=300.0, X_end=500.0),
        kb.DoubleRampGradientProfile(X_start=300.0, t_start_plateau=5.0, rate1=10.0, X_mid=500.0,
                                     t_mid_plateau=3.0, rate2=-20.0, X_end=200.0, t_end_plateau=5.0),
        kb.DoubleRampGradientProfile(X_start=300.0, t_start_plateau=5.0, rate1=10.0, X_mid=500.0,
                                     t_mid_plateau=3.0, rate2=-20.0, X_end=200.0, t_end_plateau=5.0,
                                     t_blend=0.1),
    ]
    kinds, params = zip(*[p.device_desc() for p in profs])
    h = _lib.Handle(0)
    h.set_profiles(np.array(kinds), np.array(params))
    t = np.array([-1.0, 0.0, 1.0, 2.0, 4.0, 4.95, 5.05, 15.0, 24.95, 25.0, 25.05, 27.0, 28.0, 35.0, 43.0, 45.0, 48.0, 100.0])
    X = h.eval_profile(len(profs), t)
    h.close()
    assert np.all(X[0] == 10.0) and np.all(X[1] == 300.0)
    assert X[2][list(t).index(2.0)] == pytest.approx(400.0)        # lineardirect.f(2.0) ≈ 400.0
    assert X[2][-1] == 500.0 and X[2][0] == 300.0
    assert X[3][list(t).index(2.0)] == pytest.approx(400.0)
    assert X[4][list(t).index(15.0)] == pytest.approx(400.0)
    assert X[4][list(t).index(27.0)] == pytest.approx(500.0)
    assert X[4][-1] == pytest.approx(200.0)
    for i in (4, 5):
        ref = np.array([profs[i].value_at(x) for x in t])
        assert np.max(np.abs(X[i] - ref)) < 1e-10
    assert X[5][-1] == pytest.approx(200.0, abs=1e-9)


def test_modified_arrhenius_forms_on_device(built):
    """k = A*T^n*exp(-Ea/RT)*N_A*t_mult with n = 1/2 (collision theory) and n = 1 (Eyring) on the device
    against the host calculators, a few ulp (pow on the device)."""
    import kinetica_b200 as kb
    from kinetica_b200 import _lib
    from kinetica_b200.synthetic import synthetic_crn, SEED_BASE
    S, R = 40, 120
    sd, rd, _, _ = synthetic_crn(S, R, SEED_BASE + 31)
    rng = np.random.default_rng(9)
    Ts = np.array([300.0, 512.5, 850.0, 1200.0, 2000.0])
    calcs = [kb.CollisionTheoryCalculator(rng.uniform(0, 2e5, R), rng.uniform(1, 40, R) * 1.66054e-27, rng.uniform(1, 9, R) * 1e-19,
                                          rng.uniform(0.01, 1, R), k_max=1e12),
             kb.EyringCalculator(rng.uniform(2e4, 2e5, R), rng.uniform(-80, 40, R))]
    for calc in calcs:
        h = _lib.Handle(0)
        h.set_network(S, *rd.flatten())
        h.symbolic(4)
        d = calc.device_arrhenius()
        h.set_arrhenius(d["A"], d["Ea"], d["n"], d["k_max"], d["t_mult"])
        k = h.eval_k(Ts)
        h.close()
        for b, T in enumerate(Ts):
            ref = calc(T=T)
            assert np.all(np.abs(k[:, b] - ref) <= 8 * np.spacing(np.abs(ref)) + 1e-300)
