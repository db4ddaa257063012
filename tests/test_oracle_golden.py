"""CPU tests pinning the oracle (and the host mirror of the reference API) against everything the
reference ships for this path: the profile known answers of reference test/Main/conditions.jl
(ported 1:1) and the Ea/A fixture examples/getting_started/arrhenius_params.bson.  The solver
boundary itself is unpinned upstream and is pinned here by closed forms,
literature values and conservation laws."""
import json
import math
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _impls():
    import kinetica_b200.conditions as host
    from oracle import kinetica_oracle as ko

    class HostAdapter:      # keyword constructors like the reference
        StaticConditionProfile = host.StaticConditionProfile
        NullDirectProfile = staticmethod(lambda X_start, t_end: host.NullDirectProfile(X_start=X_start, t_end=t_end))
        LinearDirectProfile = staticmethod(lambda rate, X_start, X_end: host.LinearDirectProfile(rate=rate, X_start=X_start, X_end=X_end))
        NullGradientProfile = staticmethod(lambda X_start, t_end: host.NullGradientProfile(X_start=X_start, t_end=t_end))
        LinearGradientProfile = staticmethod(lambda rate, X_start, X_end: host.LinearGradientProfile(rate=rate, X_start=X_start, X_end=X_end))
        DoubleRampGradientProfile = staticmethod(lambda *a, **k: host.DoubleRampGradientProfile(
            **dict(zip(["X_start", "t_start_plateau", "rate1", "X_mid", "t_mid_plateau", "rate2", "X_end", "t_end_plateau"], a)), **k))
        ConditionSet = host.ConditionSet
        f = staticmethod(lambda p, t: p.f(t, p))
        grad = staticmethod(lambda p, t: p.grad(t, p))

    class OracleAdapter:
        StaticConditionProfile = ko.StaticConditionProfile
        NullDirectProfile = ko.NullDirectProfile
        LinearDirectProfile = ko.LinearDirectProfile
        NullGradientProfile = ko.NullGradientProfile
        LinearGradientProfile = ko.LinearGradientProfile
        DoubleRampGradientProfile = ko.DoubleRampGradientProfile
        ConditionSet = ko.ConditionSet
        f = staticmethod(lambda p, t: p.f(t))
        grad = staticmethod(lambda p, t: p.grad(t))
    return [OracleAdapter, HostAdapter]


@pytest.mark.parametrize("impl", [0, 1])
def test_profile_construction(impl):
    """reference test/Main/conditions.jl:4-90"""
    M = _impls()[impl]
    assert M.StaticConditionProfile(10.0).value == 10.0
    nd = M.NullDirectProfile(300.0, 10.0)
    assert nd.X_start == 300.0 and nd.t_end == 10.0 and M.f(nd, 5.0) == pytest.approx(300.0)
    assert len(nd.tstops) == 1 and nd.tstops[0] == pytest.approx(10.0)
    ld = M.LinearDirectProfile(50.0, 300.0, 500.0)
    assert (ld.rate, ld.X_start, ld.X_end) == (50.0, 300.0, 500.0)
    assert ld.t_end == pytest.approx(4.0) and M.f(ld, 2.0) == pytest.approx(400.0)
    assert len(ld.tstops) == 1 and ld.tstops[0] == pytest.approx(4.0)
    ng = M.NullGradientProfile(300.0, 10.0)
    assert ng.X_start == 300.0 and ng.t_end == 10.0 and M.grad(ng, 5.0) == 0.0
    assert len(ng.tstops) == 1 and ng.tstops[0] == pytest.approx(10.0)
    lg = M.LinearGradientProfile(50.0, 300.0, 500.0)
    assert lg.t_end == pytest.approx(4.0) and M.grad(lg, 2.0) == 50.0 and M.grad(lg, 5.0) == 0.0
    assert M.grad(lg, -1.0) == 50.0                       # Appendix A.1: rate for ALL t <= t_end
    assert len(lg.tstops) == 1 and lg.tstops[0] == pytest.approx(4.0)
    dr = M.DoubleRampGradientProfile(300.0, 5.0, 10.0, 500.0, 3.0, -20.0, 200.0, 5.0)
    assert (dr.rate1, dr.rate2, dr.X_start, dr.X_mid, dr.X_end) == (10.0, -20.0, 300.0, 500.0, 200.0)
    assert (dr.t_start_plateau, dr.t_mid_plateau, dr.t_end_plateau, dr.t_blend) == (5.0, 3.0, 5.0, 0.0)
    assert dr.t_end == pytest.approx(48.0)
    assert np.allclose(dr.tstops, [5.0, 25.0, 28.0, 43.0, 48.0])
    for t, g in [(1.0, 0.0), (15.0, 10.0), (27.0, 0.0), (35.0, -20.0), (45.0, 0.0), (100.0, 0.0)]:
        assert M.grad(dr, t) == g
    drb = M.DoubleRampGradientProfile(300.0, 5.0, 10.0, 500.0, 3.0, -20.0, 200.0, 5.0, t_blend=0.1)
    assert drb.t_blend == 0.1
    assert np.allclose(drb.tstops, [4.9, 5.1, 24.9, 25.1, 27.9, 28.1, 42.9, 43.1, 48.0])


@pytest.mark.parametrize("impl", [0, 1])
def test_conditionset_construction(impl):
    """reference test/Main/conditions.jl:92-135"""
    M = _impls()[impl]

    def mk():
        return {"T": M.LinearDirectProfile(50.0, 300.0, 500.0),
                "P": M.DoubleRampGradientProfile(1e5, 1.0, 1e3, 2e5, 10.0, -1e3, 1e5, 1.0, t_blend=0.1),
                "V": 1e3}
    csc = M.ConditionSet(mk())
    assert set(csc.symbols) == {"T", "P", "V"} and len(csc.profiles) == 3
    assert csc.discrete_updates is False and csc.ts_update is None
    csd = M.ConditionSet(mk(), ts_update=1e-3)
    assert csd.discrete_updates is True and csd.ts_update == pytest.approx(1e-3)
    with pytest.raises(ValueError):
        M.ConditionSet({"X": "abc"})
    # discrete tstops: LinearDirect 0:1e-3:4 -> 4001 points ending exactly at t_end
    ts = csd.get_profile("T").tstops
    assert len(ts) == 4001 and ts[0] == 0.0 and ts[-1] == 4.0 and ts[1234] == 1.234
    allts = csd.get_tstops()
    assert np.all(np.diff(allts) > 0) and allts[0] == 0.0 and allts[-1] == pytest.approx(212.0)
    with pytest.raises(ValueError):
        M.ConditionSet({"T": M.LinearDirectProfile(50.0, 300.0, 500.0)}, ts_update=5.0)


def test_julia_range_semantics():
    """`0.0:0.001:14.0` has exactly 14001 correctly rounded elements."""
    from oracle import kinetica_oracle as ko
    from kinetica_b200.conditions import create_savepoints
    for fn in (ko.create_savepoints, create_savepoints):
        r = fn(0.0, 14.0, 0.001)
        assert len(r) == 14001 and r[-1] == 14.0 and r[7] == 0.007 and r[13999] == 13.999
        r = fn(0.0, 1.0, 0.3)
        assert np.array_equal(r, [0.0, 0.3, 0.6, 0.9, 1.0])           # final time appended
        r = fn(0.0, 0.35, 0.1)
        assert np.allclose(r, [0.0, 0.1, 0.2, 0.3, 0.35]) and r[3] == 0.3
        assert len(fn(0.0, 5.0, 1.0000000001)) == 6                   # sigdigits=9 step clean-up


def test_profile_solutions_agree():
    """Host mirror vs oracle: tabulated profile solutions and interpolated values at tstops."""
    import kinetica_b200 as kb
    from oracle import kinetica_oracle as ko
    pars = kb.ODESimulationParams(tspan=(0.0, 48.0), u0=[1.0], save_interval=None, solve_chunks=False)
    h = kb.DoubleRampGradientProfile(X_start=300.0, t_start_plateau=5.0, rate1=10.0, X_mid=500.0, t_mid_plateau=3.0,
                                     rate2=-20.0, X_end=200.0, t_end_plateau=5.0, t_blend=0.1)
    o = ko.DoubleRampGradientProfile(300.0, 5.0, 10.0, 500.0, 3.0, -20.0, 200.0, 5.0, t_blend=0.1)
    h.solve(pars); o.solve((0.0, 48.0), None)
    assert np.array_equal(h.sol.t, o.sol.t) and np.allclose(h.sol.u, o.sol.u, rtol=1e-14, atol=0)
    assert h.sol(48.0) == pytest.approx(200.0, abs=1e-9) and h.sol(26.0) == pytest.approx(500.0, abs=1e-9)
    assert o.maximum() == pytest.approx(500.0, abs=1e-9) and h.minimum() == pytest.approx(200.0, abs=1e-9)


def _fixture():
    d = json.load(open(os.path.join(HERE, "golden", "arrhenius_params.json")))
    return (np.array([float.fromhex(x) for x in d["Ea"]]), np.array([float.fromhex(x) for x in d["A"]]))


def test_arrhenius_fixture():
    """Shipped Ea/A (30 reactions; 8 barrierless; max Ea 595 362.9 J/mol) through the restated
    formula of calculator.jl:223-232; numpy restatement == plain-C restatement == host mirror."""
    import kinetica_b200 as kb
    from oracle import c_oracle as co, kinetica_oracle as ko
    Ea, A = _fixture()
    assert len(Ea) == 30 and int(np.sum(Ea == 0)) == 8 and Ea.max() == pytest.approx(595362.9117963027)
    assert A.min() == pytest.approx(596071344.862709) and A.max() == pytest.approx(2055992694198.0383)
    for k_max in (1e12, None):
        calc = ko.PrecalculatedArrheniusCalculator(Ea, A, k_max=k_max)
        host = kb.PrecalculatedArrheniusCalculator(Ea, A, k_max=k_max)
        for T in (300.0, 500.0, 850.0, 1200.0):
            k = calc(T)
            kc = co.arrhenius(A, Ea, T, k_max)
            assert np.all(np.abs(kc - k) <= 2 * np.spacing(k))
            assert np.array_equal(host(T=T), k)
    calc = ko.PrecalculatedArrheniusCalculator(Ea, A, k_max=1e12)
    assert calc(500.0)[0] == pytest.approx(1.1476254735240006e-28, rel=1e-13)
    assert calc(500.0)[1] == pytest.approx(9.9999898092210303e+11, rel=1e-13)
    assert calc(850.0)[4] == pytest.approx(1.2031111957791838e+05, rel=1e-13)
    assert calc(1200.0)[0] == pytest.approx(4.187600263086772e+07, rel=1e-13)
    # t_unit scaling and exp underflow
    assert ko.PrecalculatedArrheniusCalculator(Ea, A, t_unit="mins").t_mult == 60.0
    assert ko.PrecalculatedArrheniusCalculator(np.array([6e5]), np.array([1.0]))(10.0)[0] == 0.0


def test_rodas4_tableau_order_conditions():
    """The Rodas4 coefficients used by both the CUDA path and the C oracle satisfy the eight
    order-4 conditions (and the embedded solution the order-3 ones)."""
    g = 0.25
    A = np.zeros((6, 6)); C = np.zeros((6, 6))
    A[1, 0] = 0.1544000000000000e+01
    A[2, :2] = [0.9466785280815826e+00, 0.2557011698983284e+00]
    A[3, :3] = [0.3314825187068521e+01, 0.2896124015972201e+01, 0.9986419139977817e+00]
    A[4, :4] = [0.1221224509226641e+01, 0.6019134481288629e+01, 0.1253708332932087e+02, -0.6878860361058950e+00]
    A[5, :5] = list(A[4, :4]) + [1.0]
    C[1, 0] = -0.5668800000000000e+01
    C[2, :2] = [-0.2430093356833875e+01, -0.2063599157091915e+00]
    C[3, :3] = [-0.1073529058151375e+00, -0.9594562251023355e+01, -0.2047028614809616e+02]
    C[4, :4] = [0.7496443313967647e+01, -0.1024680431464352e+02, -0.3399990352819905e+02, 0.1170890893206160e+02]
    C[5, :5] = [0.8083246795921522e+01, -0.7981132988064893e+01, -0.3152159432874371e+02, 0.1631930543123136e+02, -0.6058818238834054e+01]
    G = np.linalg.inv(np.eye(6) / g - C)
    alpha = A @ G
    beta = alpha + G
    bi, ai = beta.sum(1), alpha.sum(1)
    b = np.array(list(A[4, :4]) + [1.0, 1.0]) @ G
    bh = np.array(list(A[4, :4]) + [1.0, 0.0]) @ G
    conds = lambda w: [w.sum() - 1, w @ bi - 0.5, w @ ai ** 2 - 1 / 3, w @ (beta @ bi) - 1 / 6, w @ ai ** 3 - 0.25,
                       w @ (ai * (alpha @ bi)) - 1 / 8, w @ (beta @ ai ** 2) - 1 / 12, w @ (beta @ (beta @ bi)) - 1 / 24]
    assert np.max(np.abs(conds(b))) < 1e-13
    assert np.max(np.abs(conds(bh)[:4])) < 1e-13 and abs(conds(bh)[4]) > 1e-3
    assert np.allclose(ai, [0, 0.386, 0.21, 0.63, 1, 1], atol=1e-14)
    # non-autonomous form (continuous rate updates, kb2_kernels.cuh cT / cD): stage times c_i = row sums
    # of alpha, time-derivative weights d_i = gamma_i = row sums of Gamma = G's inverse relation
    # (beta - alpha = G), and c = A d in the transformed tableau
    d = G.sum(1)
    assert np.allclose(d, [0.25, -0.1043, 0.1035, -0.3620000000000023e-01, 0.0, 0.0], atol=1e-13)
    assert np.allclose(A @ d, ai, atol=1e-13)


def test_c_oracle_known_answers(built):
    """The plain-C oracle (the CUDA path's algorithmic twin) against closed forms and Robertson."""
    from oracle import c_oracle as co, kinetica_oracle as ko
    NA = ko.N_A
    # 2A -> B non-combinatoric: A = A0/(1 + 2 k A0 t)
    net = ko.Network(2, [[0]], [[1]], [[2]], [[1]])
    save = np.linspace(0, 1, 9)
    out, st, _, _ = co.solve_rodas4(net, np.array([3.0]) / NA, np.zeros(1), None, 1.0, [300.0], None, None, [0.5, 0.0], (0.0, 1.0), save)
    assert st[0] == 0 and np.allclose(out[0][:, 0], 0.5 / (1 + 3.0 * save), rtol=1e-6)
    # Robertson (literature: Hairer & Wanner)
    net = ko.Network(3, [[0], [1], [1, 2]], [[1], [1, 2], [0, 2]], [[1], [2], [1, 1]], [[1], [1, 1], [1, 1]])
    k = np.array([0.04, 3e7, 1e4])
    out, st, _, _ = co.solve_rodas4(net, k / NA, np.zeros(3), None, 1.0, [300.0], None, None, [1.0, 0, 0], (0.0, 40.0),
                                    np.array([0.0, 0.4, 4.0, 40.0]))
    assert st[0] == 0
    assert np.allclose(out[0][1], [0.98517, 3.3864e-5, 1.4794e-2], rtol=2e-5)
    assert np.allclose(out[0][3], [0.7158, 9.185e-6, 0.2842], rtol=2e-4)
    assert np.max(np.abs(out[0].sum(axis=1) - 1)) < 1e-13
    # zero-order hold: k switches from 1 to 3 at t = 0.5 for A -> B
    net = ko.Network(2, [[0]], [[1]], [[1]], [[1]])
    ref = ko.solve_trajectory(net, [1.0, 0.0], np.array([[1.0], [3.0]]), np.array([0.0, 0.5]), (0.0, 1.0), np.array([0.5, 1.0]),
                              rtol=1e-12, atol=1e-15)
    assert ref[0, 0] == pytest.approx(math.exp(-0.5), rel=1e-9) and ref[1, 0] == pytest.approx(math.exp(-0.5 - 1.5), rel=1e-9)


def test_c3_radau_fixture_agrees_with_the_c_twin(built):
    """tests/golden/c3_radau.npz (scipy Radau on two members of the C3 sweep, full horizon; written by
    tests/golden/make_c3_radau.py) against the plain-C Rodas4 twin at the bench tolerances: two
    independent integrators of the oracle on the bench network, bound 1e-4 relative + 1e-9 (the
    Rodas4 global error level at abstol 1e-10 / reltol 1e-8).  The GPU test
    (tests/test_gpu_baseline_configs.py) checks the CUDA path against the same fixture."""
    from kinetica_b200.synthetic import synthetic_crn, synthetic_u0, SEED_BASE
    from oracle import c_oracle as co, kinetica_oracle as ko
    gold = np.load(os.path.join(HERE, "golden", "c3_radau.npz"))
    S, R = 1000, 5000
    sd, rd, Ea, A = synthetic_crn(S, R, SEED_BASE + 3)
    net = ko.Network(S, rd.id_reacs, rd.id_prods, rd.stoic_reacs, rd.stoic_prods)
    Ts = [float(t) for t in gold["T0"]]
    assert gold["u"].shape == (len(Ts), 11, S) and list(gold["members"]) == [4, 7]
    ts = ko.create_savepoints(0.0, 1.0, 1e-2)
    ref, st, stats, _ = co.solve_rodas4(net, A, Ea, 1e12, 1.0, Ts, ts, lambda b, t: Ts[b] + 100.0 * min(t, 1.0),
                                        synthetic_u0(S), (0.0, 1.0), gold["save_t"], nthreads=len(Ts),
                                        abstol=1e-10, reltol=1e-8)
    assert np.all(st == 0)
    worst = np.max(np.abs(ref - gold["u"]) / (1e-4 * np.abs(gold["u"]) + 1e-9))
    assert worst < 1.0, worst
    # conservation of the synthetic networks' mass balance is not needed here: positivity and the initial state
    assert np.array_equal(gold["u"][:, 0, :], np.tile(synthetic_u0(S), (len(Ts), 1)))
    assert gold["u"].min() > -1e-12
