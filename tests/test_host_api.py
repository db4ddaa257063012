"""CPU tests of the host logic above the C ABI and of the ABI itself (no compute calls: there is
no GPU here and no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built):
    from kinetica_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "kinetica_b200.h")).read()
    declared = set(re.findall(r"\b(kb2_[a-z_A-Z0-9]+)\s*\(", hdr))
    assert len(declared) >= 28
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/kinetica_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), "ctypes binding out of sync with the header"


def test_no_cpu_fallback(built):
    """Without a CUDA device the product path fails loudly."""
    import torch
    from kinetica_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.Kb2Error, match="no CPU fallback"):
        _lib.Handle(0)
    h = _lib.Handle(-1)                      # host-only: symbolic analysis only
    h.set_network(2, [0, 1], [0], [1], [0, 1], [1], [1])
    h.symbolic(0)
    with pytest.raises(_lib.Kb2Error, match="no CPU fallback"):
        h.eval_rhs(np.zeros((2, 1)), np.zeros((1, 1)))


@pytest.mark.parametrize("S,R,seed", [(12, 30, 1), (150, 700, 2), (400, 2000, 3)])
def test_symbolic_bit_exact_vs_oracle(built, S, R, seed):
    """Sparsity pattern, ordering, L\\U pattern and FMA count of kb2_symbolic are bit-identical to
    the numpy restatement (pattern contract)."""
    from kinetica_b200 import _lib
    from kinetica_b200.synthetic import synthetic_crn
    from oracle import kinetica_oracle as ko
    sd, rd, Ea, A = synthetic_crn(S, R, seed)
    h = _lib.Handle(-1)
    h.set_network(S, *rd.flatten())
    nnzJ, nnzLU, nfma = h.symbolic(0)
    net = ko.Network(S, rd.id_reacs, rd.id_prods, rd.stoic_reacs, rd.stoic_prods)
    colptr, rowval = net.pattern_csc()
    cp, rv = h.get_pattern()
    assert np.array_equal(cp, colptr) and np.array_equal(rv, rowval) and nnzJ == len(rowval)
    perm = ko.min_degree_order(S, colptr, rowval)
    assert np.array_equal(h.get_ordering(), perm)
    rowptr, colidx, diagpos, n = ko.symbolic_lu(S, colptr, rowval, perm)
    rp, ci, dp = h.get_lu_pattern()
    assert np.array_equal(rp, rowptr) and np.array_equal(ci, colidx) and np.array_equal(dp, diagpos)
    assert n == nfma and nnzLU == len(colidx)
    # natural and caller-supplied orderings
    h.symbolic(1)
    assert np.array_equal(h.get_ordering(), np.arange(S))
    p2 = np.random.default_rng(seed).permutation(S)
    h.symbolic(perm=p2)
    assert np.array_equal(h.get_ordering(), p2)
    rowptr, colidx, diagpos, n = ko.symbolic_lu(S, colptr, rowval, p2)
    rp, ci, dp = h.get_lu_pattern()
    assert np.array_equal(rp, rowptr) and np.array_equal(ci, colidx) and np.array_equal(dp, diagpos)


def test_pattern_edge_cases(built):
    """Net-zero species create columns but no rows (Appendix A.10); empty networks; bad input."""
    from kinetica_b200 import _lib
    h = _lib.Handle(-1)
    # B + C -> A + C : C is a reactant with zero net stoichiometry
    h.set_network(3, [0, 2], [1, 2], [1, 1], [0, 2], [0, 2], [1, 1])
    h.symbolic(1)
    cp, rv = h.get_pattern()
    assert list(cp) == [0, 0, 2, 4] and list(rv) == [0, 1, 0, 1]
    h.set_network(4, [0], [], [], [0], [], [])          # no reactions: identity W
    assert h.symbolic(0) == (0, 4, 0)
    with pytest.raises(_lib.Kb2Error, match="out of range"):
        h.set_network(2, [0, 1], [5], [1], [0, 1], [1], [1])
    with pytest.raises(_lib.Kb2Error, match="more than 3"):
        h.set_network(5, [0, 4], [0, 1, 2, 3], [1, 1, 1, 1], [0, 1], [4], [1])


def test_params_validation():
    """reference src/solving/params.jl:76-104"""
    import kinetica_b200 as kb
    p = kb.ODESimulationParams(tspan=(0.0, 14.0), u0={"C": 1.0})
    assert (p.abstol, p.reltol, p.maxiters, p.solve_chunks, p.solve_chunkstep) == (1e-10, 1e-8, 100000, True, 1e-3)
    assert p.low_k_cutoff == "auto" and p.low_k_maxconc == 2.0 and p.ban_negatives is False and p.save_interval is None
    with pytest.raises(ValueError, match="Invalid time span"):
        kb.ODESimulationParams(tspan=(1.0, 1.0), u0=[1.0])
    with pytest.raises(ValueError, match="low_k_cutoff"):
        kb.ODESimulationParams(tspan=(0.0, 1.0), u0=[1.0], low_k_cutoff="sometimes")
    with pytest.raises(ValueError, match="low_k_cutoff"):
        kb.ODESimulationParams(tspan=(0.0, 1.0), u0=[1.0], low_k_cutoff=-1.0)
    with pytest.raises(ValueError, match="not divisible"):
        kb.ODESimulationParams(tspan=(0.0, 1.0), u0=[1.0], solve_chunkstep=0.3)
    with pytest.raises(ValueError, match="save interval"):
        kb.ODESimulationParams(tspan=(0.0, 1.0), u0=[1.0], save_interval=0.1)


def test_method_constructors_and_host_preprocessing():
    """methods.jl:12-20,49-57; calculator.jl:200-209; solve_utils.jl:19-54,213-297; filters.jl:40-52."""
    import kinetica_b200 as kb
    from kinetica_b200 import solve as sv
    from kinetica_b200.synthetic import getting_started_standin
    sd, rd = getting_started_standin()
    assert sd.n == 10 and rd.nr == 30
    Ea = np.linspace(0, 3e5, 30); A = np.full(30, 1e10)
    calc = kb.PrecalculatedArrheniusCalculator(Ea, A, k_max=1e12)
    pars = kb.ODESimulationParams(tspan=(0.0, 4.0), u0={"C": 1.0}, solve_chunks=False)
    var = kb.ConditionSet({"T": kb.LinearGradientProfile(rate=50.0, X_start=300.0, X_end=500.0)}, ts_update=1e-2)
    with pytest.raises(ValueError, match="static"):
        kb.StaticODESolve(pars, var, calc)
    with pytest.raises(ValueError, match="does not support"):
        kb.VariableODESolve(pars, kb.ConditionSet({"T": 300.0, "P": 1e5}), calc)
    with pytest.raises(ValueError, match="Number of parameters"):
        kb.PrecalculatedArrheniusCalculator(Ea[:5], A[:5]).setup_network(sd, rd)
    # max rates: the permutation with the largest MEAN k
    var.solve_variable_conditions(pars)
    assert var.get_profile("T").minimum() == 300.0 and var.get_profile("T").maximum() == 500.0
    assert np.array_equal(sv.get_max_rates(var, calc), calc(T=500.0))
    assert np.array_equal(sv.get_initial_rates(var, calc), calc(T=300.0))
    low = sv.low_k_removal(calc(T=500.0), pars)
    assert np.array_equal(low, np.nonzero(calc(T=500.0) * 4.0 < 1e-8 / 4.0)[0])
    # discrete rate table: one row per tstop, condition read from the interpolated profile
    ts, ktab = sv.calculate_discrete_rates(var, calc, 30)
    assert len(ts) == 401 and ktab.shape == (401, 30) and np.array_equal(ktab[100], calc(T=var.get_profile("T").sol(1.0)))
    # u0
    assert sv.make_u0(sd, pars)[0] == 1.0 and sv.make_u0(sd, pars).sum() == 1.0
    with pytest.raises(RuntimeError, match="not in SpeciesData"):
        sv.make_u0(sd, kb.ODESimulationParams(tspan=(0.0, 1.0), u0={"XYZ": 1.0}, solve_chunks=False))
    with pytest.raises(RuntimeError, match="does not match"):
        sv.make_u0(sd, kb.ODESimulationParams(tspan=(0.0, 1.0), u0=[1.0, 2.0], solve_chunks=False))
    short = sv.make_u0(sd, kb.ODESimulationParams(tspan=(0.0, 1.0), u0=[1.0, 2.0], solve_chunks=False, allow_short_u0=True))
    assert list(short[:3]) == [1.0, 2.0, 0.0] and len(short) == 10
    # filters + splice keep order and compact every field
    flt = kb.RxFilter([lambda s, r: [j < 3 for j in range(r.nr)], lambda s, r: [j == 29 for j in range(r.nr)]])
    mask = sv.get_filter_mask(flt, sd, rd)
    assert list(np.nonzero(mask)[0]) == [0, 1, 2, 29]
    keep = kb.RxFilter(flt.filters, keep_filtered=True)
    assert int(np.sum(sv.get_filter_mask(keep, sd, rd))) == 26
    rd2 = sv.copy.deepcopy(rd)
    rd2.splice(np.nonzero(mask)[0])
    assert rd2.nr == 26 and rd2.id_reacs[0] == rd.id_reacs[3] and rd2.id_prods[-1] == rd.id_prods[28]
    calc.splice(np.nonzero(mask)[0])
    assert len(calc.Ea) == 26 and calc.Ea[0] == Ea[3]


def test_stop_merging():
    from kinetica_b200.solve import merge_stops
    t, f = merge_stops(np.array([0.0, 0.25, 0.5, 0.75, 1.0]), np.array([0.0, 0.5, 1.0]), 0.0, 1.0)
    assert list(t) == [0.0, 0.25, 0.5, 0.75, 1.0] and list(f) == [3, 1, 3, 1, 3]
    t, f = merge_stops(None, np.array([0.0, 0.4, 0.8]), 0.0, 1.0)       # tf always a stop
    assert list(t) == [0.0, 0.4, 0.8, 1.0] and list(f) == [2, 2, 2, 0]
    t, f = merge_stops(np.array([0.5, 2.0]), np.array([0.0, 1.0]), 0.0, 1.0)  # stops beyond tf dropped
    assert list(t) == [0.0, 0.5, 1.0] and list(f) == [2, 1, 2]


def test_chunk_grid_follows_the_reference_formula():
    """methods.jl:214-222: n_chunks = Int(tspan[2] / chunkstep), saveat_local = 0:save_interval:chunkstep,
    (len(saveat_local) - 1) * n_chunks + 1 points, global time = local + nc * chunkstep."""
    import kinetica_b200 as kb
    from kinetica_b200.solve import chunk_grid, merge_stops, STOP_CHUNK, STOP_SAVE, STOP_RATE
    p = kb.ODESimulationParams(tspan=(0.0, 14.0), u0=[1.0])
    bounds, save = chunk_grid(p)
    assert len(save) == 14001 and len(bounds) == 13999 and save[0] == 0.0 and save[-1] == 0.001 + 13999 * 0.001
    p = kb.ODESimulationParams(tspan=(0.0, 1.0), u0=[1.0], solve_chunkstep=0.25, save_interval=0.05)
    bounds, save = chunk_grid(p)
    assert len(save) == (6 - 1) * 4 + 1 and np.allclose(save, np.arange(21) * 0.05) and bounds.tolist() == [0.25, 0.5, 0.75]
    st, fl = merge_stops([0.0, 0.5, 1.0], save, 0.0, 1.0, bounds)
    assert fl[list(st).index(0.5)] == (STOP_RATE | STOP_SAVE | STOP_CHUNK) and fl[list(st).index(0.25)] == (STOP_SAVE | STOP_CHUNK)
    assert fl[-1] & STOP_CHUNK == 0                      # tf is not the start of a chunk


def test_modified_arrhenius_calculators_match_their_documented_rate_laws():
    """Collision theory (docs/src/tutorials/kinetic-calculators.md:144-148) and Eyring (:73-77) written
    out directly against the A*T^n*exp(-E/RT)*N_A*t_mult form the device evaluates."""
    import kinetica_b200 as kb
    from kinetica_b200.calculator import K_B, H_PLANCK, N_A, R_GAS
    rng = np.random.default_rng(5)
    R = 12
    Ea = rng.uniform(0, 2e5, R); mu = rng.uniform(1, 40, R) * 1.66054e-27; sig = rng.uniform(1, 9, R) * 1e-19; rho = rng.uniform(0.01, 1, R)
    calc = kb.CollisionTheoryCalculator(Ea, mu, sig, rho, k_max=1e12)
    for T in (300.0, 850.0, 1500.0):
        kr = sig * rho * N_A * np.sqrt(8 * K_B * T / (np.pi * mu)) * np.exp(-Ea / (R_GAS * T))
        assert np.allclose(calc(T=T), 1.0 / (1.0 / 1e12 + 1.0 / kr), rtol=1e-13)
    dev = calc.device_arrhenius()
    assert np.all(dev["n"] == 0.5) and dev["k_max"] == 1e12 and calc.has_conditions(["T"]) and calc.allows_continuous()
    dH = rng.uniform(2e4, 2e5, R); dS = rng.uniform(-80, 40, R)
    ey = kb.EyringCalculator(dH, dS)
    for T in (300.0, 1200.0):
        assert np.allclose(ey(T=T, P=1e5), K_B * T / H_PLANCK * np.exp(dS / R_GAS) * np.exp(-dH / (R_GAS * T)), rtol=1e-13)
    assert np.all(ey.device_arrhenius()["n"] == 1.0) and ey.has_conditions(["T", "P"])
    ey.splice([0, 3])
    assert len(ey.Ea) == R - 2 and len(ey.n) == R - 2
