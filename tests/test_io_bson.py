"""save_output / load_output (reference src/analysis/io.jl:50-255): round trip of an ODESolveOutput
through the BSON dictionary tree, and the Julia tagging of bits arrays against the form the
reference's own shipped file uses (tests/golden: arrhenius_params.bson was parsed with the same
reader).  CPU only."""
import numpy as np

import kinetica_b200 as kb
from kinetica_b200 import io as kio
from kinetica_b200.solve import ODESolveOutput, RateSolution, Solution


def _fake_output(continuous=False):
    sd = kb.SpeciesData(["C", "[H]", "[CH3]", "CC"])
    rd = kb.RxData([[0], [1, 2], [2]], [[1, 2], [0], [3]], [[1], [1, 1], [2]], [[1, 1], [1], [1]])
    pars = kb.ODESimulationParams(tspan=(0.0, 2.0), u0={"C": 1.0}, solve_chunks=False, save_interval=0.5, low_k_cutoff="none")
    prof = kb.LinearGradientProfile(rate=50.0, X_start=500.0, X_end=600.0)
    cs = kb.ConditionSet({"T": prof, "P": 1.0e5}, ts_update=None if continuous else 0.5)
    cs.solve_variable_conditions(pars)
    t = np.arange(5) * 0.5
    rng = np.random.default_rng(1)
    sol = Solution(t=t, u=[rng.uniform(0, 1, sd.n) for _ in t])
    ts = cs.get_tstops()
    sol_k = None if continuous else RateSolution(ts, rng.uniform(1, 2, (len(ts), rd.nr)))
    from kinetica_b200.conditions import _Sol
    sol_vcs = {"T": _Sol(t, prof.values_at(t))} if continuous else None
    return ODESolveOutput(sd=sd, rd=rd, sol=sol, sol_k=sol_k, sol_vcs=sol_vcs, pars=pars, conditions=cs)


def test_round_trip_discrete(tmp_path):
    out = _fake_output()
    p = str(tmp_path / "out.bson")
    kio.save_output(out, p)
    back = kio.load_output(p)
    assert back.sd.toInt == out.sd.toInt and back.sd.n == 4
    assert back.rd.id_reacs == out.rd.id_reacs and back.rd.stoic_prods == out.rd.stoic_prods and back.rd.nr == 3
    assert np.array_equal(back.sol.t, out.sol.t) and all(np.array_equal(a, b) for a, b in zip(back.sol.u, out.sol.u))
    assert np.array_equal(back.sol_k.t, out.sol_k.t) and np.array_equal(back.sol_k.u, out.sol_k.u) and back.sol_vcs is None
    assert back.pars.tspan == (0.0, 2.0) and back.pars.u0 == {"C": 1.0} and back.pars.low_k_cutoff == "none"
    assert back.pars.save_interval == 0.5 and back.pars.solve_chunks is False and back.pars.maxiters == 100000
    assert back.conditions.symbols == ["T", "P"] and back.conditions.discrete_updates and back.conditions.ts_update == 0.5
    prof = back.conditions.get_profile("T")
    assert type(prof).__name__ == "LinearGradientProfile" and prof.rate == 50.0 and prof.X_end == 600.0
    assert np.array_equal(prof.tstops, out.conditions.get_profile("T").tstops)
    assert prof.value_at(1.0) == 550.0 and back.conditions.get_profile("P").value == 1.0e5
    assert np.array_equal(back.umax, np.max(np.array(out.sol.u), axis=0))


def test_round_trip_continuous(tmp_path):
    out = _fake_output(continuous=True)
    p = str(tmp_path / "out.bson")
    kio.save_output(out, p)
    back = kio.load_output(p)
    assert back.sol_k is None and np.allclose(back.sol_vcs["T"].u, out.sol_vcs["T"].u)
    assert not back.conditions.discrete_updates and back.conditions.ts_update is None


def test_tree_layout_matches_the_reference_keys_and_tagging(tmp_path):
    """Top-level keys of io.jl:112-166, 1-based indices on disk, and bits arrays tagged exactly like
    the arrays of the reference's shipped arrhenius_params.bson."""
    out = _fake_output()
    p = str(tmp_path / "out.bson")
    kio.save_output(out, p)
    raw, _ = kio._dec_doc(open(p, "rb").read())
    assert set(raw) == {"KineticaCoreVersion", "sd", "rd", "pars", "sol", "conditions"}
    assert set(raw["rd"]) == {"nr", "mapped_rxns", "id_reacs", "id_prods", "stoic_reacs", "stoic_prods", "dH", "rhash", "level_found"}
    assert set(raw["sol"]) == {"u", "t", "vcs", "k"} and raw["sol"]["vcs"] is None
    t = raw["sol"]["t"]
    assert t["tag"] == "array" and t["type"] == {"tag": "datatype", "name": ["Core", "Float64"], "params": []} and t["size"] == [5]
    assert np.array_equal(np.frombuffer(t["data"], dtype="<f8"), out.sol.t)
    first = raw["rd"]["id_reacs"]["data"][1]
    assert first["type"]["name"] == ["Core", "Int64"] and list(np.frombuffer(first["data"], dtype="<i8")) == [2, 3]   # 1-based
    assert raw["pars"]["solver"] == {"tag": "symbol", "name": "B200Rodas4"} and raw["pars"]["tspan"]["tag"] == "tuple"
    assert raw["sd"]["toInt"]["tag"] == "dict" and raw["conditions"]["symbols"]["data"][0] == {"tag": "symbol", "name": "T"}
