"""GPU parity tests, solve level: solve_network through the C ABI against (a) closed forms under
the reference rate-law convention, (b) the plain-C oracle running the same Rodas4 algorithm,
(c) the independent scipy Radau oracle.  Stated tolerance: relative 1e-6 on
species >= 1e-9, absolute 1e-9 below."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
RTOL, FLOOR = 1e-6, 1e-9


def _check(got, ref, rtol=RTOL, floor=FLOOR):
    """|got - ref| <= rtol * (|ref| + floor/rtol*...) — relative `rtol` above the abstol floor,
    absolute `floor * rtol / 1e-6`-scaled below it: err <= rtol*|ref| + floor."""
    got, ref = np.asarray(got), np.asarray(ref)
    lim = rtol * np.abs(ref) + floor
    worst = np.max(np.abs(got - ref) / lim)
    assert worst < 1.0, f"parity violated: worst |d|/(rtol|ref|+floor) = {worst:.3g}"


def _solve_const(kb, species, reacs, prods, sr, sp, rates, u0, tspan, save_interval, **kw):
    sd = kb.SpeciesData(species)
    rd = kb.RxData(reacs, prods, sr, sp)
    calc = kb.DummyKineticCalculator(rates)
    pars = kb.ODESimulationParams(tspan=tspan, u0=u0, save_interval=save_interval, low_k_cutoff="none",
                                  solve_chunks=False, **kw)
    conds = kb.ConditionSet({"T": 300.0})
    return kb.solve_network(kb.StaticODESolve(pars, conds, calc), sd, rd)


def test_closed_forms(built):
    import kinetica_b200 as kb
    # A -> B : u_A = exp(-k t)
    res = _solve_const(kb, ["A", "B"], [[0]], [[1]], [[1]], [[1]], [2.0], [1.0, 0.0], (0.0, 2.0), 0.25)
    t = res.sol.t
    U = np.array(res.sol.u)
    assert res.sol.retcode == "Success" and len(t) == 9
    _check(U[:, 0], np.exp(-2.0 * t)); _check(U[:, 1], 1 - np.exp(-2.0 * t))
    # 2A -> B, non-combinatoric rate law: dA/dt = -2 k A^2  =>  A = A0/(1 + 2 k A0 t)
    res = _solve_const(kb, ["A", "B"], [[0]], [[1]], [[2]], [[1]], [3.0], [0.5, 0.0], (0.0, 1.0), 0.125)
    t = res.sol.t; U = np.array(res.sol.u)
    _check(U[:, 0], 0.5 / (1 + 2 * 3.0 * 0.5 * t)); _check(U[:, 1], (0.5 - U[:, 0]) / 2, rtol=1e-5)
    # same reaction written with the species listed twice (id_reacs [A, A], stoich [1, 1])
    res2 = _solve_const(kb, ["A", "B"], [[0, 0]], [[1]], [[1, 1]], [[1]], [3.0], [0.5, 0.0], (0.0, 1.0), 0.125)
    _check(np.array(res2.sol.u), U, rtol=1e-9)
    # A <-> B
    res = _solve_const(kb, ["A", "B"], [[0], [1]], [[1], [0]], [[1], [1]], [[1], [1]], [3.0, 1.0], [1.0, 0.0], (0.0, 3.0), 0.5)
    t = res.sol.t; U = np.array(res.sol.u)
    _check(U[:, 0], 0.25 + 0.75 * np.exp(-4.0 * t))
    # A + B -> C with A0 != B0
    res = _solve_const(kb, ["A", "B", "C"], [[0, 1]], [[2]], [[1, 1]], [[1]], [2.0], [1.0, 0.5, 0.0], (0.0, 2.0), 0.25)
    t = res.sol.t; U = np.array(res.sol.u)
    a0, b0, k = 1.0, 0.5, 2.0
    x = a0 * b0 * (np.exp((a0 - b0) * k * t) - 1) / (a0 * np.exp((a0 - b0) * k * t) - b0)
    _check(U[:, 2], x); _check(U[:, 0], a0 - x); _check(U[:, 1], b0 - x)


def test_robertson(built):
    """Robertson (net-zero catalyst species and stoichiometry 2) against literature values and the
    Radau oracle."""
    import kinetica_b200 as kb
    from oracle import kinetica_oracle as ko
    res = _solve_const(kb, ["A", "B", "C"], [[0], [1], [1, 2]], [[1], [1, 2], [0, 2]], [[1], [2], [1, 1]],
                       [[1], [1, 1], [1, 1]], [0.04, 3e7, 1e4], [1.0, 0.0, 0.0], (0.0, 40.0), 4.0)
    U = np.array(res.sol.u)
    assert abs(U[-1, 0] - 0.7158) < 1e-4 and abs(U[-1, 1] - 9.185e-6) < 1e-9 and abs(U[-1, 2] - 0.2842) < 1e-4
    net = ko.Network(3, [[0], [1], [1, 2]], [[1], [1, 2], [0, 2]], [[1], [2], [1, 1]], [[1], [1, 1], [1, 1]])
    ref = ko.solve_trajectory(net, [1.0, 0, 0], np.array([0.04, 3e7, 1e4]), None, (0.0, 40.0), res.sol.t,
                              rtol=1e-12, atol=1e-16)
    _check(U, ref)
    assert np.max(np.abs(U.sum(axis=1) - 1.0)) < 1e-12          # conservation


@pytest.fixture(scope="module")
def ensemble_case(built):
    import kinetica_b200 as kb
    from kinetica_b200.synthetic import synthetic_crn, synthetic_u0, SEED_BASE
    S, R, B = 80, 320, 21
    sd, rd, Ea, A = synthetic_crn(S, R, SEED_BASE + 21)
    calc = kb.PrecalculatedArrheniusCalculator(Ea, A, k_max=1e12)
    # solver tolerances two decades below the reference defaults so that the solver's own global
    # error sits below the stated parity tolerance (1e-6 relative above the 1e-9 floor)
    pars = kb.ODESimulationParams(tspan=(0.0, 1.0), u0=synthetic_u0(S), save_interval=0.1,
                                  low_k_cutoff="none", solve_chunks=False, abstol=1e-12, reltol=1e-10)
    Ts = [600.0 + 600.0 * b / (B - 1) for b in range(B)]
    conds = [kb.ConditionSet({"T": kb.LinearDirectProfile(rate=100.0, X_start=T, X_end=T + 100.0)},
                             ts_update=1e-2) for T in Ts]
    outs = kb.solve_network(kb.B200EnsembleODESolve(pars, conds, calc), sd, rd)
    return sd, rd, Ea, A, Ts, conds, outs


def test_ensemble_vs_c_oracle(ensemble_case):
    """Same algorithm on CPU (oracle/crn_oracle.c): trajectories and step counts."""
    from oracle import c_oracle as co, kinetica_oracle as ko
    sd, rd, Ea, A, Ts, conds, outs = ensemble_case
    net = ko.Network(sd.n, rd.id_reacs, rd.id_prods, rd.stoic_reacs, rd.stoic_prods)
    ts = conds[0].get_tstops()
    assert len(ts) == 101
    u0 = np.zeros(sd.n); u0[8:18] = 0.1
    ref, st, stats, save_t = co.solve_rodas4(net, A, Ea, 1e12, 1.0, Ts, ts,
                                             lambda b, t: Ts[b] + 100.0 * min(t, 1.0), u0, (0.0, 1.0), outs[0].sol.t,
                                             abstol=1e-12, reltol=1e-10)
    assert np.all(st == 0)
    assert np.array_equal(save_t, outs[0].sol.t) and len(save_t) == 11
    for b, o in enumerate(outs):
        assert o.sol.retcode == "Success"
        _check(np.array(o.sol.u), ref[b])
        assert abs(int(o.sol.stats[0]) - int(stats[b, 0])) <= max(5, 0.05 * stats[b, 0])
        assert np.array_equal(o.umax, np.max(np.array(o.sol.u), axis=0))


@pytest.mark.parametrize("mb", [2, 4])
def test_tile_sizes_agree(ensemble_case, monkeypatch, mb):
    """The ensemble solved with 2 and 4 members per warp tile (the sizes large ensembles run at;
    this small one defaults to 1) against the 1-member-per-tile solution, ragged last tile included."""
    import kinetica_b200 as kb
    from kinetica_b200.synthetic import synthetic_u0
    sd, rd, Ea, A, Ts, conds, outs = ensemble_case
    monkeypatch.setenv("KB2_MB", str(mb))
    calc = kb.PrecalculatedArrheniusCalculator(Ea, A, k_max=1e12)
    pars = kb.ODESimulationParams(tspan=(0.0, 1.0), u0=synthetic_u0(sd.n), save_interval=0.1,
                                  low_k_cutoff="none", solve_chunks=False, abstol=1e-12, reltol=1e-10)
    got = kb.solve_network(kb.B200EnsembleODESolve(pars, conds, calc), sd, rd)
    for a, b in zip(got, outs):
        assert a.sol.retcode == "Success"
        _check(np.array(a.sol.u), np.array(b.sol.u))
        assert np.array_equal(a.umax, np.max(np.array(a.sol.u), axis=0))


def test_default_tolerances_vs_tight(ensemble_case):
    """Reference-default tolerances (abstol 1e-10, reltol 1e-8) against the tight solution: the
    solver's own global error, bounded at 1e-4 relative."""
    import kinetica_b200 as kb
    from kinetica_b200.synthetic import synthetic_u0
    sd, rd, Ea, A, Ts, conds, outs = ensemble_case
    calc = kb.PrecalculatedArrheniusCalculator(Ea, A, k_max=1e12)
    pars = kb.ODESimulationParams(tspan=(0.0, 1.0), u0=synthetic_u0(sd.n), save_interval=0.1,
                                  low_k_cutoff="none", solve_chunks=False)
    loose = kb.solve_network(kb.B200EnsembleODESolve(pars, conds, calc), sd, rd)
    for a, b in zip(loose, outs):
        assert a.sol.retcode == "Success"
        _check(np.array(a.sol.u), np.array(b.sol.u), rtol=1e-4)


def test_ensemble_vs_radau(ensemble_case):
    """Independent integrator (scipy Radau, rtol 1e-10) on three members."""
    from oracle import kinetica_oracle as ko
    sd, rd, Ea, A, Ts, conds, outs = ensemble_case
    net = ko.Network(sd.n, rd.id_reacs, rd.id_prods, rd.stoic_reacs, rd.stoic_prods)
    calc = ko.PrecalculatedArrheniusCalculator(Ea, A, k_max=1e12)
    u0 = np.zeros(sd.n); u0[8:18] = 0.1
    ts = conds[0].get_tstops()
    cons = net.conservation_basis()
    for b in (0, 20):
        ktab = np.array([calc(Ts[b] + 100.0 * min(t, 1.0)) for t in ts])
        ref = ko.solve_trajectory(net, u0, ktab, ts, (0.0, 1.0), outs[b].sol.t, k_init=calc(Ts[b]),
                                  rtol=1e-10, atol=1e-14)
        U = np.array(outs[b].sol.u)
        _check(U, ref)
        if len(cons):
            assert np.max(np.abs((U - u0) @ cons.T)) < 1e-11      # conservation laws
        assert U.min() > -1e-9


def test_getting_started_standin(built):
    """BASELINE configs[0]/[1] on the stand-in 30-reaction methane CRN with the shipped Ea/A and
    k_max = 1e12 (raw reference formula incl. N_A): static 1000 K, then the docs' ramp
    (LinearGradientProfile 50 K/s, ts_update 1e-3, shortened to 0.2 s)."""
    import kinetica_b200 as kb
    from kinetica_b200.synthetic import getting_started_standin
    from oracle import c_oracle as co, kinetica_oracle as ko
    d = json.load(open(os.path.join(HERE, "golden", "arrhenius_params.json")))
    Ea = np.array([float.fromhex(x) for x in d["Ea"]]); A = np.array([float.fromhex(x) for x in d["A"]])
    sd, rd = getting_started_standin()
    net = ko.Network(sd.n, rd.id_reacs, rd.id_prods, rd.stoic_reacs, rd.stoic_prods)
    calc = kb.PrecalculatedArrheniusCalculator(Ea, A, k_max=1e12)
    u0 = np.zeros(sd.n); u0[0] = 1.0
    # static
    pars = kb.ODESimulationParams(tspan=(0.0, 0.05), u0={"C": 1.0}, save_interval=0.01, low_k_cutoff="none",
                                  solve_chunks=False, ban_negatives=True)
    res = kb.solve_network(kb.StaticODESolve(pars, kb.ConditionSet({"T": 1000.0}), calc), sd, rd)
    ref, st, _, _ = co.solve_rodas4(net, A, Ea, 1e12, 1.0, [1000.0], None, None, u0, (0.0, 0.05), res.sol.t,
                                    ban_negatives=True)
    assert st[0] == 0 and res.sol.retcode == "Success"
    _check(np.array(res.sol.u), ref[0], rtol=1e-5)
    # variable, discrete updates
    cs = kb.ConditionSet({"T": kb.LinearGradientProfile(rate=50.0, X_start=500.0, X_end=1200.0)}, ts_update=1e-3)
    pars = kb.ODESimulationParams(tspan=(0.0, 0.2), u0={"C": 1.0}, save_interval=0.05, low_k_cutoff="none",
                                  solve_chunks=False, ban_negatives=True)
    res = kb.solve_network(kb.VariableODESolve(pars, cs, calc), sd, rd)
    ts = cs.get_tstops()
    ref, st, _, _ = co.solve_rodas4(net, A, Ea, 1e12, 1.0, [500.0], ts, lambda b, t: 500.0 + 50.0 * t, u0,
                                    (0.0, 0.2), res.sol.t, ban_negatives=True)
    assert st[0] == 0 and res.sol.retcode == "Success"
    _check(np.array(res.sol.u), ref[0], rtol=1e-5)


def test_low_k_cutoff_and_filter(built):
    """Host pre-processing defines R and the indexing (solve_utils.jl:213-245, filters.jl:40-52)."""
    import kinetica_b200 as kb
    from kinetica_b200.synthetic import synthetic_crn, synthetic_u0, SEED_BASE
    from oracle import kinetica_oracle as ko
    S, R = 40, 120
    sd, rd, Ea, A = synthetic_crn(S, R, SEED_BASE + 31)
    calc = kb.PrecalculatedArrheniusCalculator(Ea, A, k_max=1e12)
    pars = kb.ODESimulationParams(tspan=(0.0, 1.0), u0=synthetic_u0(S), save_interval=0.5, solve_chunks=False)
    flt = kb.RxFilter([lambda sd_, rd_: [j % 7 == 0 for j in range(rd_.nr)]])
    res = kb.solve_network(kb.StaticODESolve(pars, kb.ConditionSet({"T": 700.0}), calc, flt), sd, rd)
    keep = np.array([j for j in range(R) if j % 7 != 0])
    k = ko.PrecalculatedArrheniusCalculator(Ea[keep], A[keep], k_max=1e12)(700.0)
    low = ko.low_k_removal_set(k, 1e-8, 1.0)
    assert res.rd.nr == len(keep) - len(low) and len(low) > 0
    keep2 = np.delete(keep, low)
    assert res.rd.id_reacs == [rd.id_reacs[j] for j in keep2]          # order preserved, compacted
    assert rd.nr == R                                                   # copy_network=True left the input alone


def test_failure_raises(built):
    import kinetica_b200 as kb
    with pytest.raises(RuntimeError, match="ODE solution failed."):
        _solve_const(kb, ["A", "B"], [[0]], [[1]], [[1]], [[1]], [2.0], [1.0, 0.0], (0.0, 2.0), 0.25,
                     maxiters=3, adaptive_tols=False)


def test_ragged_ensemble_sizes(built):
    import kinetica_b200 as kb
    from kinetica_b200.synthetic import synthetic_crn, synthetic_u0, SEED_BASE
    S, R = 30, 100
    sd, rd, Ea, A = synthetic_crn(S, R, SEED_BASE + 41)
    calc = kb.PrecalculatedArrheniusCalculator(Ea, A, k_max=1e12)
    pars = kb.ODESimulationParams(tspan=(0.0, 0.5), u0=synthetic_u0(S), save_interval=0.25, low_k_cutoff="none",
                                  solve_chunks=False)
    base = None
    for B in (1, 5, 33, 64):
        conds = [kb.ConditionSet({"T": 900.0}) for _ in range(B)]
        outs = kb.solve_network(kb.B200EnsembleODESolve(pars, conds, calc), sd, rd)
        U = np.array([np.array(o.sol.u) for o in outs])
        if base is None:
            base = U[0]
        assert np.all(U == base[None])       # identical members: bit-identical, any tile position


def test_more_tiles_than_resident_warps(built, monkeypatch):
    """More tiles than the phase kernels keep resident is fine as well (one member per tile, 6000
    tiles: every CTA walks several tiles), with members that need different numbers of steps (two
    temperatures) in the same launches: finished tiles are skipped, results stay bit-identical."""
    import kinetica_b200 as kb
    from kinetica_b200.synthetic import synthetic_crn, synthetic_u0, SEED_BASE
    S, R, B = 24, 80, 6000
    sd, rd, Ea, A = synthetic_crn(S, R, SEED_BASE + 43)
    calc = kb.PrecalculatedArrheniusCalculator(Ea, A, k_max=1e12)
    pars = kb.ODESimulationParams(tspan=(0.0, 0.2), u0=synthetic_u0(S), save_interval=0.1, low_k_cutoff="none",
                                  solve_chunks=False)
    conds = [kb.ConditionSet({"T": 700.0 if b % 3 else 1100.0}) for b in range(B)]
    monkeypatch.setenv("KB2_MB", "1")
    outs = kb.solve_network(kb.B200EnsembleODESolve(pars, conds, calc), sd, rd)
    U = np.array([np.array(o.sol.u) for o in outs])
    assert all(o.sol.retcode == "Success" for o in outs)
    hot, cold = U[0], U[1]
    assert not np.array_equal(hot, cold)
    for b in range(B):
        assert np.array_equal(U[b], cold if b % 3 else hot)
    monkeypatch.setenv("KB2_MB", "4")
    ref = kb.solve_network(kb.B200EnsembleODESolve(pars, conds, calc), sd, rd)
    # another tile size only changes the order of the per-member reductions
    _check(np.array([np.array(o.sol.u) for o in ref]), U, rtol=1e-5)


def test_batch_tiling_matches_one_shot(built):
    """kb2_solve walks an ensemble in batch tiles when it does not fit the device memory (here forced:
    40 members in tiles of 16, last tile ragged): every member's result is what the one-shot solve gives."""
    import kinetica_b200 as kb
    from kinetica_b200.synthetic import synthetic_crn, synthetic_u0, SEED_BASE
    S, R, B = 48, 160, 40
    sd, rd, Ea, A = synthetic_crn(S, R, SEED_BASE + 47)
    calc = kb.PrecalculatedArrheniusCalculator(Ea, A, k_max=1e12)
    pars = kb.ODESimulationParams(tspan=(0.0, 0.5), u0=synthetic_u0(S), save_interval=0.125, low_k_cutoff="none", solve_chunks=False)
    conds = [kb.ConditionSet({"T": kb.LinearDirectProfile(rate=100.0, X_start=650.0 + 10.0 * b, X_end=700.0 + 10.0 * b)},
                             ts_update=0.05) for b in range(B)]
    for cs in conds:
        cs.solve_variable_conditions(pars)
    es = kb.EnsembleSolver(sd, rd, calc)
    plan = es.h.memory_plan(5)
    assert plan["bytes_per_member"] > 8 * 12 * S and plan["b_tile"] >= B
    es.h.set_tiling(1)
    whole = es.solve(conds, pars, synthetic_u0(S))
    assert es.h.last_batch_tiles == 1
    es.h.set_batch_tile(16)
    tiled = es.solve(conds, pars, synthetic_u0(S))
    assert es.h.last_batch_tiles == 3
    es.close()
    for a, b in zip(whole[:4], tiled[:4]):
        assert np.array_equal(a, b)
    assert np.all(whole[2] == 0)
