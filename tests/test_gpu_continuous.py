"""Continuous rate updates on the device (reference src/solving/methods.jl:363-458):
k follows the member's condition profile inside every Rodas4 stage, with the non-autonomous df/dt
term.  Parity: against scipy Radau integrating du/dt = f(u, k(T(t))) directly, and as the limit of
the discrete mode for ts_update -> 0."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check(got, ref, rtol, floor=1e-9):
    got, ref = np.asarray(got), np.asarray(ref)
    worst = np.max(np.abs(got - ref) / (rtol * np.abs(ref) + floor))
    assert worst < 1.0, f"parity violated: worst |d|/(rtol|ref|+floor) = {worst:.3g}"


@pytest.fixture(scope="module")
def case(built):
    import kinetica_b200 as kb
    from kinetica_b200.synthetic import synthetic_crn, synthetic_u0, SEED_BASE
    S, R = 80, 320
    sd, rd, Ea, A = synthetic_crn(S, R, SEED_BASE + 21)
    return kb, sd, rd, Ea, A, synthetic_u0(S)


def test_continuous_vs_radau(case):
    """A ramp that ends inside tspan (kink at t_end = 0.6) and a double ramp with blending, three
    members; tolerances two decades below the reference defaults, stated bound 1e-6 + 1e-9."""
    from oracle import kinetica_oracle as ko
    kb, sd, rd, Ea, A, u0 = case
    calc = kb.PrecalculatedArrheniusCalculator(Ea, A, k_max=1e12)
    ocalc = ko.PrecalculatedArrheniusCalculator(Ea, A, k_max=1e12)
    net = ko.Network(sd.n, rd.id_reacs, rd.id_prods, rd.stoic_reacs, rd.stoic_prods)
    pars = kb.ODESimulationParams(tspan=(0.0, 1.0), u0=u0, save_interval=0.1, low_k_cutoff="none", solve_chunks=False,
                                  abstol=1e-12, reltol=1e-10)
    profs = [kb.LinearDirectProfile(rate=300.0, X_start=700.0, X_end=880.0),
             kb.LinearGradientProfile(rate=150.0, X_start=900.0, X_end=1050.0),
             kb.DoubleRampGradientProfile(X_start=650.0, t_start_plateau=0.1, rate1=900.0, X_mid=920.0, t_mid_plateau=0.2,
                                          rate2=-600.0, X_end=800.0, t_end_plateau=0.2, t_blend=0.02)]
    conds = [kb.ConditionSet({"T": p}) for p in profs]                 # no ts_update: continuous
    assert not conds[0].discrete_updates
    pars2 = kb.ODESimulationParams(tspan=(0.0, 0.9), u0=u0, save_interval=0.1, low_k_cutoff="none", solve_chunks=False,
                                   abstol=1e-12, reltol=1e-10)
    for cs, p in zip(conds, (pars, pars, pars2)):
        res = kb.solve_network(kb.VariableODESolve(p, cs, calc), sd, rd)
        prof = cs.get_profile("T")
        assert res.sol.retcode == "Success" and res.sol_k is None
        assert np.allclose(res.sol_vcs["T"].u, prof.values_at(res.sol.t))          # sol_vcs: T at the save times
        ref = ko.solve_trajectory_continuous(net, u0, lambda t: ocalc(float(prof.value_at(t))), cs.get_tstops(),
                                             p.tspan, res.sol.t, rtol=1e-10, atol=1e-14)
        _check(np.array(res.sol.u), ref, rtol=1e-6)


def test_discrete_converges_to_continuous(case):
    """Zero-order hold with ts_update -> 0 approaches the continuous solve: the distance shrinks
    about linearly with ts_update (first-order hold error)."""
    kb, sd, rd, Ea, A, u0 = case
    calc = kb.PrecalculatedArrheniusCalculator(Ea, A, k_max=1e12)
    pars = kb.ODESimulationParams(tspan=(0.0, 0.5), u0=u0, save_interval=0.1, low_k_cutoff="none", solve_chunks=False,
                                  abstol=1e-12, reltol=1e-10)
    mk = lambda: kb.LinearDirectProfile(rate=200.0, X_start=800.0, X_end=900.0)
    cont = kb.solve_network(kb.VariableODESolve(pars, kb.ConditionSet({"T": mk()}), calc), sd, rd)
    Uc = np.array(cont.sol.u)
    errs = []
    for dt in (2e-2, 5e-3, 1.25e-3):
        d = kb.solve_network(kb.VariableODESolve(pars, kb.ConditionSet({"T": mk()}, ts_update=dt), calc), sd, rd)
        U = np.array(d.sol.u)
        errs.append(np.max(np.abs(U - Uc) / (np.abs(Uc) + 1e-6)))
    assert errs[0] > errs[1] > errs[2] and errs[2] < 0.15 * errs[0]
    assert 2.0 < errs[0] / errs[1] < 8.0                                  # ~4x per 4x finer updates


def test_continuous_ensemble_and_chunks(case):
    """An ensemble of ramps in continuous mode, chunkwise (the reference default) against complete."""
    kb, sd, rd, Ea, A, u0 = case
    calc = kb.PrecalculatedArrheniusCalculator(Ea, A, k_max=1e12)
    conds = [kb.ConditionSet({"T": kb.LinearDirectProfile(rate=100.0, X_start=650.0 + 40.0 * b, X_end=750.0 + 40.0 * b)})
             for b in range(9)]
    whole = kb.ODESimulationParams(tspan=(0.0, 1.0), u0=u0, save_interval=0.125, low_k_cutoff="none", solve_chunks=False,
                                   abstol=1e-12, reltol=1e-10)
    chunk = kb.ODESimulationParams(tspan=(0.0, 1.0), u0=u0, save_interval=0.125, low_k_cutoff="none", solve_chunks=True,
                                   solve_chunkstep=0.25, abstol=1e-12, reltol=1e-10)
    a = kb.solve_network(kb.B200EnsembleODESolve(whole, conds, calc), sd, rd)
    b = kb.solve_network(kb.B200EnsembleODESolve(chunk, conds, calc), sd, rd)
    for x, y in zip(a, b):
        assert x.sol.retcode == y.sol.retcode == "Success" and set(x.sol_vcs) == {"T"}
        _check(np.array(y.sol.u), np.array(x.sol.u), rtol=1e-6)
