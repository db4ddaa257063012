"""`identify_next_seeds` (reference src/exploration/explore_utils.jl:338-409) on top of the
per-species maxima of a solve: selection rule, `ignore`, `elim_small_na`, the two-argument form,
the `seeds.out` file layout, and the ensemble extension.  Host logic only, no GPU."""
import numpy as np
import pytest

import kinetica_b200 as kb
from kinetica_b200.seeds import julia_float_repr
from kinetica_b200.solve import ODESolveOutput, Solution


def _case():
    sd = kb.SpeciesData(["C", "[H]", "[CH3]", "CC", "C=C"])
    sd.xyz = {0: {"N_atoms": 5}, 1: {"N_atoms": 1}, 2: {"N_atoms": 4}, 3: {"N_atoms": 8}, 4: {"N_atoms": 6}}
    t = np.array([0.0, 0.5, 1.0])
    u = [np.array([1.0, 0.0, 0.0, 0.0, 0.0]), np.array([0.6, 0.02, 0.3, 0.05, 1e-5]), np.array([0.4, 0.001, 0.1, 0.2, 0.01])]
    return sd, Solution(t=t, u=u)


def test_threshold_rule_and_order():
    sd, sol = _case()
    assert kb.identify_next_seeds(sol, sd, 0.05) == ["C", "[CH3]", "CC"]          # max >= seed_conc, species-id order
    assert kb.identify_next_seeds(sol, sd, 0.2) == ["C", "[CH3]", "CC"]           # CC reaches exactly 0.2: >= keeps it
    assert kb.identify_next_seeds(sol, sd, 0.2000001) == ["C", "[CH3]"]
    assert kb.identify_next_seeds(sol, sd) == ["C", "[H]", "[CH3]", "CC", "C=C"]  # two-argument form: everything


def test_ignore_and_small_species():
    sd, sol = _case()
    assert kb.identify_next_seeds(sol, sd, 0.01, ignore=["C"]) == ["[H]", "[CH3]", "CC", "C=C"]
    assert kb.identify_next_seeds(sol, sd, 0.01, elim_small_na=5) == ["C", "CC", "C=C"]
    assert kb.identify_next_seeds(sol, sd, elim_small_na=2, ignore=["CC"]) == ["C", "[CH3]", "C=C"]
    bare = kb.SpeciesData(["A", "B"])
    with pytest.raises(KeyError):
        kb.identify_next_seeds(Solution(t=np.zeros(1), u=[np.ones(2)]), bare, 0.1, elim_small_na=3)


def test_uses_device_maxima_when_present():
    sd, sol = _case()
    out = ODESolveOutput(sd=sd, rd=None, sol=sol, umax=np.array([1.0, 0.5, 0.3, 0.2, 0.01]))
    assert kb.identify_next_seeds(out, sd, 0.4) == ["C", "[H]"]                    # umax, not the coarse saves
    out.umax = None
    assert kb.identify_next_seeds(out, sd, 0.4) == ["C"]


def test_seeds_out_layout(tmp_path):
    sd, sol = _case()
    f = tmp_path / "seeds.out"
    seeds = kb.identify_next_seeds(sol, sd, 0.05, saveto=str(f))
    lines = f.read_text().splitlines()
    assert lines[0] == "3"
    assert lines[1] == "SID   SMILES   Max. Conc."
    assert lines[2] == "1     C       1.0"          # rpad(sid, 5) + " " + rpad(smi, 5) + "   " + conc
    assert lines[3] == "2     [CH3]   0.3"
    assert lines[4] == "3     CC      0.2"
    assert seeds == ["C", "[CH3]", "CC"]
    with pytest.raises(ValueError):
        kb.identify_next_seeds(sol, sd, 5.0, saveto=str(f))       # no seeds: the reference's maximum() over an empty list throws


def test_ensemble_takes_the_maximum_over_members():
    sd, sol = _case()
    other = Solution(t=sol.t, u=[np.array([1.0, 0.0, 0.0, 0.0, 0.0]), np.array([0.9, 0.3, 0.0, 0.0, 0.0])])
    assert kb.identify_next_seeds(other, sd, 0.25) == ["C", "[H]"]
    assert kb.identify_next_seeds_ensemble([sol, other], sd, 0.25) == ["C", "[H]", "[CH3]"]
    assert kb.identify_next_seeds_ensemble([sol], sd, 0.25) == kb.identify_next_seeds(sol, sd, 0.25)


@pytest.mark.parametrize("x,s", [(0.1, "0.1"), (1.0, "1.0"), (1e-5, "1.0e-5"), (1e-4, "0.0001"), (123456.7, "123456.7"),
                                 (1e6, "1.0e6"), (999999.0, "999999.0"), (2.5e-7, "2.5e-7"), (1.2345e10, "1.2345e10"),
                                 (0.00012, "0.00012"), (3e-10, "3.0e-10"), (0.0, "0.0"), (-0.5, "-0.5"),
                                 (1234567.0, "1.234567e6"), (0.30000000000000004, "0.30000000000000004")])
def test_julia_float_repr(x, s):
    assert julia_float_repr(x) == s
