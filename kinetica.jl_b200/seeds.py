"""Seed selection from a kinetic solve — the consumer of the per-species maxima the device keeps.

Mirrors `identify_next_seeds` of the reference (src/exploration/explore_utils.jl:338-409), which
`explore_network(::IterativeExplore, ...)` calls on every `solve_network` result
(src/exploration/methods.jl:221-240): a species becomes a seed of the next exploration level if its
maximum concentration over the saved trajectory reaches `seed_conc` (or, in the two-argument form,
every species is returned), skipping `ignore`d species and, with `elim_small_na > 0`, species with
fewer than that many atoms (`sd.xyz[species]["N_atoms"]`).  The reference recomputes the maximum
from `reduce(vcat, sol.u')`; here it is the `umax` vector the solve kernel already wrote
(`ODESolveOutput.umax`, checked equal to the maximum over `sol.u` in tests/test_gpu_solve.py), with
the trajectory as the fallback for outputs that do not carry it.

`identify_next_seeds_ensemble` is this build's extension for `B200EnsembleODESolve`: the maximum is
taken over all members first (a species that is abundant under ANY condition of the sweep seeds the
next level); with one member it is `identify_next_seeds`.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import numpy as np


def julia_float_repr(x: float) -> str:
    """`string(x::Float64)` of Julia (shortest round-trip digits; fixed notation for
    1e-4 <= |x| < 1e6, `d.ddde±x` outside ... written the way Base.Ryu.writeshortest does with its default
    arguments), which is what the reference interpolates into `seeds.out`."""
    x = float(x)
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "Inf" if x > 0 else "-Inf"
    if x == 0.0:
        return "-0.0" if np.signbit(x) else "0.0"
    r = repr(abs(x))                       # shortest round-trip digits, same as Ryu
    if "e" in r:
        mant, exp = r.split("e")
        e10 = int(exp)
    else:
        mant, e10 = r, 0
    ip, _, fp = mant.partition(".")
    digits = (ip + fp).lstrip("0")
    # decimal exponent of the first significant digit
    lead = len(ip.lstrip("0")) - 1 if ip.strip("0") else -(len(fp) - len(fp.lstrip("0")) + 1)
    e10 += lead
    digits = digits.rstrip("0") or "0"
    sign = "-" if x < 0 else ""
    if -4 <= e10 < 6:
        if e10 >= 0:
            whole = digits[:e10 + 1].ljust(e10 + 1, "0")
            frac = digits[e10 + 1:] or "0"
            return f"{sign}{whole}.{frac}"
        return f"{sign}0.{'0' * (-e10 - 1)}{digits}"
    frac = digits[1:] or "0"
    return f"{sign}{digits[0]}.{frac}e{e10}"


def _species_max(res) -> np.ndarray:
    umax = getattr(res, "umax", None)
    if umax is not None:
        return np.asarray(umax, dtype=np.float64)
    sol = getattr(res, "sol", res)
    return np.max(np.asarray(sol.u, dtype=np.float64), axis=0)      # maximum(umat[:, species])


def _select(maxc: np.ndarray, sd, seed_conc: Optional[float], elim_small_na: int, ignore: Iterable[str],
            saveto: Optional[str]) -> List[str]:
    ignore = set(ignore or ())
    seeds: List[str] = []
    concs: List[float] = []
    for species in range(len(maxc)):
        name = sd.toStr[species]
        if name in ignore:
            continue
        c = float(maxc[species])
        if seed_conc is not None and not (c >= seed_conc):
            continue
        if elim_small_na > 0:
            xyz = getattr(sd, "xyz", None)
            if xyz is None:
                raise KeyError("SpeciesData has no xyz table: elim_small_na needs sd.xyz[species]['N_atoms']")
            if xyz[species]["N_atoms"] < elim_small_na:
                continue
        seeds.append(name)
        concs.append(c)
    if saveto is not None:
        if not seeds:
            raise ValueError("reducing over an empty collection is not allowed")     # maximum(length.(String[])) in the reference
        w = max(len(s) for s in seeds)
        with open(saveto, "w") as f:
            f.write(f"{len(seeds)}\n")
            f.write(f"SID   {'SMILES'.ljust(w)}   Max. Conc.\n")
            for sid, (smi, conc) in enumerate(zip(seeds, concs), start=1):
                f.write(f"{str(sid).ljust(5)} {smi.ljust(w)}   {julia_float_repr(conc)}\n")
    return seeds


def identify_next_seeds(res, sd, seed_conc: Optional[float] = None, *, elim_small_na: int = 0,
                        ignore: Sequence[str] = (), saveto: Optional[str] = None) -> List[str]:
    """Species (SMILES, in species-id order) whose maximum concentration in `res` is at least
    `seed_conc`; all species when `seed_conc` is None (the reference's second method).
    `res`: an `ODESolveOutput` (uses its device-side `umax`) or anything with `.u` time-major."""
    return _select(_species_max(res), sd, seed_conc, elim_small_na, ignore, saveto)


def identify_next_seeds_ensemble(results: Sequence, sd, seed_conc: Optional[float] = None, *, elim_small_na: int = 0,
                                 ignore: Sequence[str] = (), saveto: Optional[str] = None) -> List[str]:
    """The same over an ensemble: maximum over members first."""
    if len(results) == 0:
        raise ValueError("empty ensemble")
    maxc = _species_max(results[0]).copy()
    for r in results[1:]:
        np.maximum(maxc, _species_max(r), out=maxc)
    return _select(maxc, sd, seed_conc, elim_small_na, ignore, saveto)
