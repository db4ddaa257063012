"""GPU parity on the BASELINE configurations themselves: the C3 bench network at
the bench tolerances against the plain-C twin and scipy Radau, the getting-started stand-in over
its full horizons (C1: tspan (0, 1); C2: the docs' 14 s ramp with ts_update 1e-3 = 14 001 tstops),
a C4-shaped network (S = 10 000: state vector and y do not fit shared memory) and the per-species
maxima against the oracle's.  Stated tolerances: 1e-4 relative + 1e-9 at the reference-default
solver tolerances (the solver's own global error level), 1e-6 + 1e-9 against the tight
references."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _check(got, ref, rtol, floor=1e-9):
    got, ref = np.asarray(got), np.asarray(ref)
    worst = np.max(np.abs(got - ref) / (rtol * np.abs(ref) + floor))
    assert worst < 1.0, f"parity violated: worst |d|/(rtol|ref|+floor) = {worst:.3g}"


def _sweep(kb, Ts, ts_update):
    return [kb.ConditionSet({"T": kb.LinearDirectProfile(rate=100.0, X_start=T, X_end=T + 100.0)}, ts_update=ts_update)
            for T in Ts]


def _c3_solve(monkeypatch, B=8):
    import kinetica_b200 as kb
    from kinetica_b200.synthetic import synthetic_crn, synthetic_u0, SEED_BASE
    S, R = 1000, 5000
    sd, rd, Ea, A = synthetic_crn(S, R, SEED_BASE + 3)
    calc = kb.PrecalculatedArrheniusCalculator(Ea, A, k_max=1e12)
    pars = kb.ODESimulationParams(tspan=(0.0, 1.0), u0=synthetic_u0(S), save_interval=0.1, low_k_cutoff="none",
                                  solve_chunks=False)           # abstol 1e-10 / reltol 1e-8: the bench's
    Ts = [600.0 + 600.0 * b / (B - 1) for b in range(B)]
    conds = _sweep(kb, Ts, 1e-2)
    monkeypatch.setenv("KB2_MB", "4")
    outs = kb.solve_network(kb.B200EnsembleODESolve(pars, conds, calc), sd, rd)
    return S, R, rd, Ea, A, pars, Ts, conds, outs


def test_c3_network_at_bench_tolerances(built, monkeypatch):
    """S = 1000 / R = 5000, seed 20261018 + 3, auto ordering, four members per tile (the bench layout:
    window LU, staged state vector), 8 members spread over the sweep, against the plain-C twin."""
    from kinetica_b200.synthetic import synthetic_u0
    from oracle import c_oracle as co, kinetica_oracle as ko
    S, R, rd, Ea, A, pars, Ts, conds, outs = _c3_solve(monkeypatch)
    net = ko.Network(S, rd.id_reacs, rd.id_prods, rd.stoic_reacs, rd.stoic_prods)
    ts = conds[0].get_tstops()
    assert len(ts) == 101
    ref, st, stats, save_t = co.solve_rodas4(net, A, Ea, 1e12, 1.0, Ts, ts, lambda b, t: Ts[b] + 100.0 * min(t, 1.0),
                                             synthetic_u0(S), (0.0, 1.0), outs[0].sol.t, nthreads=os.cpu_count() or 1,
                                             abstol=pars.abstol, reltol=pars.reltol)
    assert np.all(st == 0) and len(save_t) == 11
    for b, o in enumerate(outs):
        assert o.sol.retcode == "Success"
        U = np.array(o.sol.u)
        _check(U, ref[b], rtol=1e-4)
        _check(o.umax, ref[b].max(axis=0), rtol=1e-4)                     # per-species maxima vs the ORACLE's
        assert abs(int(o.sol.stats[0]) - int(stats[b, 0])) <= 0.05 * stats[b, 0]
        assert o.sol_k is not None and o.sol_k.u.shape == (101, R) and o.sol_vcs is None


def test_c3_network_against_radau_fixture(built, monkeypatch):
    """The same solve against the INDEPENDENT integrator of the oracle (scipy Radau, rtol 1e-8, restarted at
    every rate update) over the full horizon, two members: ten minutes of CPU per member, so the
    trajectories are a committed fixture (tests/golden/c3_radau.npz, written by
    tests/golden/make_c3_radau.py, pinned against the C twin in tests/test_oracle_golden.py)."""
    from oracle import kinetica_oracle as ko
    S, R, rd, Ea, A, pars, Ts, conds, outs = _c3_solve(monkeypatch)
    ts = conds[0].get_tstops()
    b = len(Ts) // 2
    ocalc = ko.PrecalculatedArrheniusCalculator(Ea, A, k_max=1e12)
    ktab = np.array([ocalc(Ts[b] + 100.0 * min(t, 1.0)) for t in ts])
    assert np.allclose(outs[b].sol_k.u, ktab, rtol=1e-14, atol=0)         # res.sol_k = the reference's k_precalc table
    gold = np.load(os.path.join(HERE, "golden", "c3_radau.npz"))
    assert np.allclose(outs[0].sol.t, gold["save_t"], rtol=0, atol=1e-15)
    for q, bm in enumerate(gold["members"]):
        assert gold["T0"][q] == Ts[bm] and outs[bm].sol.retcode == "Success"
        _check(np.array(outs[bm].sol.u), gold["u"][q], rtol=1e-4)


def _standin():
    import kinetica_b200 as kb
    from kinetica_b200.synthetic import getting_started_standin
    from oracle import kinetica_oracle as ko
    d = json.load(open(os.path.join(HERE, "golden", "arrhenius_params.json")))
    Ea = np.array([float.fromhex(x) for x in d["Ea"]]); A = np.array([float.fromhex(x) for x in d["A"]])
    sd, rd = getting_started_standin()
    net = ko.Network(sd.n, rd.id_reacs, rd.id_prods, rd.stoic_reacs, rd.stoic_prods)
    u0 = np.zeros(sd.n); u0[0] = 1.0
    return kb, sd, rd, net, Ea, A, u0


def test_c1_getting_started_static_full_horizon(built):
    """BASELINE configs[0]: StaticODESolve at 1000 K over the full tspan (0, 1)."""
    from oracle import c_oracle as co
    kb, sd, rd, net, Ea, A, u0 = _standin()
    calc = kb.PrecalculatedArrheniusCalculator(Ea, A, k_max=1e12)
    pars = kb.ODESimulationParams(tspan=(0.0, 1.0), u0={"C": 1.0}, save_interval=0.1, low_k_cutoff="none",
                                  solve_chunks=False, ban_negatives=True)
    res = kb.solve_network(kb.StaticODESolve(pars, kb.ConditionSet({"T": 1000.0}), calc), sd, rd)
    ref, st, _, _ = co.solve_rodas4(net, A, Ea, 1e12, 1.0, [1000.0], None, None, u0, (0.0, 1.0), res.sol.t,
                                    ban_negatives=True)
    assert st[0] == 0 and res.sol.retcode == "Success" and len(res.sol.t) == 11
    _check(np.array(res.sol.u), ref[0], rtol=1e-5)
    assert res.sol_k is None and res.sol_vcs is None


def test_c2_getting_started_ramp_full_14s(built):
    """BASELINE configs[1]: VariableODESolve, LinearGradientProfile 50 K/s from 500 K to 1200 K (14 s),
    ts_update 1e-3 -> 14 001 discrete rate updates, with the reference's DEFAULT parameters of
    docs/src/getting-started.md:43-49,66-70: solve_chunks = true (14 000 chunks of 1e-3 s, maxiters per
    chunk), save_interval = nothing (one save per chunk: 14 001 points)."""
    from oracle import c_oracle as co
    kb, sd, rd, net, Ea, A, u0 = _standin()
    calc = kb.PrecalculatedArrheniusCalculator(Ea, A, k_max=1e12)
    cs = kb.ConditionSet({"T": kb.LinearGradientProfile(rate=50.0, X_start=500.0, X_end=1200.0)}, ts_update=1e-3)
    pars = kb.ODESimulationParams(tspan=(0.0, 14.0), u0={"C": 1.0}, low_k_cutoff="none", ban_negatives=True)
    assert pars.solve_chunks and pars.solve_chunkstep == 1e-3 and pars.save_interval is None
    res = kb.solve_network(kb.VariableODESolve(pars, cs, calc), sd, rd)
    ts = cs.get_tstops()
    assert len(ts) == 14001 and res.sol.retcode == "Success"
    assert len(res.sol.t) == 14001 and res.sol.t[0] == 0.0 and abs(res.sol.t[-1] - 14.0) < 1e-9     # (2 - 1) * 14000 + 1
    ref, st, _, _ = co.solve_rodas4(net, A, Ea, 1e12, 1.0, [500.0], ts, lambda b, t: 500.0 + 50.0 * min(t, 14.0), u0,
                                    (0.0, 14.0), res.sol.t, ban_negatives=True, maxiters=10 ** 7)
    assert st[0] == 0
    _check(np.array(res.sol.u), ref[0], rtol=1e-4)
    assert res.sol_k.u.shape == (14001, rd.nr) and np.array_equal(res.sol_k.t, ts)
    # the same solve as ONE chunk runs out of the reference's maxiters = 100000 (>= 14 001 forced steps
    # with a restart after each): "ODE solution failed." as in the reference
    pars1 = kb.ODESimulationParams(tspan=(0.0, 14.0), u0={"C": 1.0}, low_k_cutoff="none", ban_negatives=True,
                                   solve_chunks=False, save_interval=0.5, adaptive_tols=False)
    with pytest.raises(RuntimeError, match="ODE solution failed."):
        kb.solve_network(kb.VariableODESolve(pars1, cs, calc), sd, rd)


def test_chunkwise_matches_complete_and_retries(built):
    """Chunk boundaries only re-initialise the integrator: a chunkwise solve agrees with the complete one
    within tolerance on its (finer) save grid; a chunk that cannot finish within maxiters is repeated
    with tightened tolerances five times on the device, then the failure stands."""
    import kinetica_b200 as kb
    from kinetica_b200.synthetic import synthetic_crn, synthetic_u0, SEED_BASE
    S, R, B = 64, 256, 6
    sd, rd, Ea, A = synthetic_crn(S, R, SEED_BASE + 100)
    calc = kb.PrecalculatedArrheniusCalculator(Ea, A, k_max=1e12)
    conds = _sweep(kb, [650.0 + 90.0 * b for b in range(B)], 0.05)
    chunked = kb.ODESimulationParams(tspan=(0.0, 1.0), u0=synthetic_u0(S), low_k_cutoff="none", solve_chunks=True,
                                     solve_chunkstep=0.125, save_interval=0.0625, abstol=1e-12, reltol=1e-10)
    whole = kb.ODESimulationParams(tspan=(0.0, 1.0), u0=synthetic_u0(S), low_k_cutoff="none", solve_chunks=False,
                                   save_interval=0.0625, abstol=1e-12, reltol=1e-10)
    a = kb.solve_network(kb.B200EnsembleODESolve(chunked, conds, calc), sd, rd)
    b = kb.solve_network(kb.B200EnsembleODESolve(whole, conds, calc), sd, rd)
    assert len(a[0].sol.t) == (3 - 1) * 8 + 1 and np.allclose(a[0].sol.t, b[0].sol.t, rtol=0, atol=1e-15)
    for x, y in zip(a, b):
        assert x.sol.retcode == "Success"
        _check(np.array(x.sol.u), np.array(y.sol.u), rtol=1e-6)
    # a chunk that needs more than maxiters attempts: five attempts (four retries), then failure
    tight = kb.ODESimulationParams(tspan=(0.0, 1.0), u0=synthetic_u0(S), low_k_cutoff="none", solve_chunks=True,
                                   solve_chunkstep=0.5, save_interval=0.5, maxiters=5)
    es = kb.EnsembleSolver(sd, rd, calc)
    out_u, umax, status, stats, _ = es.solve(conds, tight, synthetic_u0(S))
    es.close()
    assert np.all(status == 1) and np.all(stats[:, 7] == 4)
    with pytest.raises(RuntimeError, match="ODE solution failed."):
        kb.solve_network(kb.B200EnsembleODESolve(tight, conds, calc), sd, rd)


def test_c4_shape_state_vector_outside_shared_memory(built, monkeypatch):
    """S = 10 000 / R = 50 000 (BASELINE configs[3] shape): the tile's state vector (320 KB) and the
    sweeps' y do not fit shared memory, gathers go to HBM / L2; four members, short horizon."""
    import kinetica_b200 as kb
    from kinetica_b200.synthetic import synthetic_crn, synthetic_u0, SEED_BASE
    from oracle import c_oracle as co, kinetica_oracle as ko
    S, R, B = 10000, 50000, 4
    sd, rd, Ea, A = synthetic_crn(S, R, SEED_BASE + 4)
    calc = kb.PrecalculatedArrheniusCalculator(Ea, A, k_max=1e12)
    pars = kb.ODESimulationParams(tspan=(0.0, 0.03), u0=synthetic_u0(S), save_interval=0.01, low_k_cutoff="none",
                                  solve_chunks=False)
    Ts = [700.0 + 120.0 * b for b in range(B)]
    conds = _sweep(kb, Ts, 1e-2)
    monkeypatch.setenv("KB2_MB", "4")
    outs = kb.solve_network(kb.B200EnsembleODESolve(pars, conds, calc), sd, rd)
    net = ko.Network(S, rd.id_reacs, rd.id_prods, rd.stoic_reacs, rd.stoic_prods)
    ts = conds[0].get_tstops()
    ref, st, stats, _ = co.solve_rodas4(net, A, Ea, 1e12, 1.0, Ts, ts, lambda b, t: Ts[b] + 100.0 * min(t, 1.0),
                                        synthetic_u0(S), (0.0, 0.03), outs[0].sol.t, nthreads=os.cpu_count() or 1,
                                        abstol=pars.abstol, reltol=pars.reltol)
    assert np.all(st == 0)
    for b, o in enumerate(outs):
        assert o.sol.retcode == "Success"
        _check(np.array(o.sol.u), ref[b], rtol=1e-4)
        _check(o.umax, ref[b].max(axis=0), rtol=1e-4)


def test_per_member_retry_and_status(built):
    """adaptive_solve! per member: a member that fails (here: maxiters too small for the hot members only)
    is re-solved with tightened tolerances while the others keep their first result; without
    adaptive_tols the failure raises like the reference (solve_utils.jl:405-411)."""
    import kinetica_b200 as kb
    from kinetica_b200.synthetic import synthetic_crn, synthetic_u0, SEED_BASE
    S, R, B = 64, 256, 64
    sd, rd, Ea, A = synthetic_crn(S, R, SEED_BASE + 100)
    calc = kb.PrecalculatedArrheniusCalculator(Ea, A, k_max=1e12)
    conds = [kb.ConditionSet({"T": 650.0 + (500.0 if b == 17 else 0.0)}) for b in range(B)]
    base = kb.ODESimulationParams(tspan=(0.0, 1.0), u0=synthetic_u0(S), save_interval=0.5, low_k_cutoff="none",
                                  solve_chunks=False)
    ref = kb.solve_network(kb.B200EnsembleODESolve(base, conds, calc), sd, rd)
    n_cold, n_hot = int(ref[0].sol.stats[2]), int(ref[17].sol.stats[2])
    assert n_hot > n_cold + 5
    pars = kb.ODESimulationParams(tspan=(0.0, 1.0), u0=synthetic_u0(S), save_interval=0.5, low_k_cutoff="none",
                                  solve_chunks=False, maxiters=(n_cold + n_hot) // 2, adaptive_tols=False)
    with pytest.raises(RuntimeError, match="ODE solution failed."):
        kb.solve_network(kb.B200EnsembleODESolve(pars, conds, calc), sd, rd)
    es = kb.EnsembleSolver(sd, rd, calc)
    out_u, umax, status, stats, _ = es.solve(conds, pars, synthetic_u0(S))
    es.close()
    assert status[17] == 1 and np.all(np.delete(status, 17) == 0)          # MaxIters on the hot member only
