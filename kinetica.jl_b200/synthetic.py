"""Deterministic synthetic CRN generator + the stand-in getting-started network.

The synthetic generator follows SURVEY.md §8(d): reactions are produced as R/2
forward reactions plus their R/2 reverses appended as a block (the layout
`ingest_cde_run` produces, reference src/exploration/cde.jl:299-313), with a
locality window and a handful of hub species (radical pool).  Duplicate and
no-op reactions are rejected like `RxData` does (reference
src/exploration/network.jl:268-287).

Everything here is host-side input preparation; nothing is on the hot path.
"""
from __future__ import annotations

import numpy as np

from .network import RxData, SpeciesData

N_A = 6.02214076e23  # reference src/constants.jl:5

SEED_BASE = 20261018
# Rate-parameter ranges of the synthetic CRNs.  SURVEY.md §8(d) proposed the shipped fixture's
# ranges (log10 A in 8.7..12.3, Ea up to 4e5 J/mol, 25 % barrierless); with abundant species at
# 0.1 mol/dm3 that gives opposing fluxes of ~1e10 mol/dm3/s whose FP64 cancellation noise
# (~1e-6 /s) sits above abstol = 1e-10 and pins ANY integrator at h ~ 1e-5 (measured with the CPU
# oracle: > 1e5 steps per member).  The ranges below keep the network stiff (rate constants span
# ~1e-11..1e7, fastest time scale ~1e-6 s over a 1 s horizon) but well conditioned in FP64.
LOG10_A = (3.0, 7.0)
EA_MAX = 2.0e5


def synthetic_crn(S: int, R: int, seed: int, w: float = 16.0, n_hubs: int = 8,
                  p_hub: float = 0.1):
    """Return (sd, rd, Ea, A) for a synthetic stiff mass-action CRN.

    `A` is already divided by N_A so that the reference formula
    k = A*exp(-Ea/RT)*N_A*t_mult (reference src/solving/calculator.jl:223-232)
    lands in a physical range with t_mult = 1.
    """
    if R % 2:
        raise ValueError("R must be even (forward + reverse blocks)")
    rng = np.random.Generator(np.random.PCG64(seed))
    nf = R // 2
    seen = set()
    fwd = []
    # every species carries an integer "mass" and every reaction balances it, so sum_i m_i u_i is a
    # positive conservation law: concentrations stay bounded (random unbalanced A -> B + C networks
    # are autocatalytic and blow up, which no real CRN does)
    mass = rng.integers(1, 5, S)
    attempts = 0
    while len(fwd) < nf:
        attempts += 1
        if attempts > 2000 * nf + 20000:
            raise ValueError(f"synthetic_crn: no {nf} distinct mass-balanced forward reactions among {S} species "
                             f"(window {w}, {n_hubs} hubs): ask for fewer reactions")
        kind = rng.random()
        c = int(rng.integers(0, S))

        def pick():
            if rng.random() < p_hub:
                return int(rng.integers(0, n_hubs))
            v = c + int(np.rint(rng.normal(0.0, w)))
            return min(max(v, 0), S - 1)

        if kind < 0.60:      # A + B -> C + D
            reac = [pick(), pick()]
            prod = [pick(), pick()]
        elif kind < 0.85:    # A -> B + C
            reac = [pick()]
            prod = [pick(), pick()]
        elif kind < 0.95:    # A -> B
            reac = [pick()]
            prod = [pick()]
        else:                # 2A -> B + C
            a = pick()
            reac = [a, a]
            prod = [pick(), pick()]
        reac.sort()
        prod.sort()
        if reac == prod:
            continue
        if sum(int(mass[x]) for x in reac) != sum(int(mass[x]) for x in prod):
            continue
        key = (tuple(reac), tuple(prod))
        rkey = (tuple(prod), tuple(reac))
        if key in seen or rkey in seen:
            continue
        seen.add(key)
        fwd.append((reac, prod))

    def compress(lst):
        ids, nu = [], []
        for s in lst:
            if ids and ids[-1] == s:
                nu[-1] += 1
            else:
                ids.append(s)
                nu.append(1)
        return ids, nu

    id_reacs, id_prods, st_reacs, st_prods = [], [], [], []
    for reac, prod in fwd:
        i, n = compress(reac)
        id_reacs.append(i); st_reacs.append(n)
        i, n = compress(prod)
        id_prods.append(i); st_prods.append(n)
    for reac, prod in fwd:
        i, n = compress(prod)
        id_reacs.append(i); st_reacs.append(n)
        i, n = compress(reac)
        id_prods.append(i); st_prods.append(n)

    Ea_f = rng.uniform(0.0, EA_MAX, nf)
    Ea_r = rng.uniform(0.0, EA_MAX, nf)
    zero_sel = rng.random(nf) < 0.5        # half the pairs get one barrierless direction => 25 % zeros
    zero_dir = rng.random(nf) < 0.5
    Ea_f[zero_sel & zero_dir] = 0.0
    Ea_r[zero_sel & ~zero_dir] = 0.0
    Ea = np.concatenate([Ea_f, Ea_r])
    A = 10.0 ** rng.uniform(LOG10_A[0], LOG10_A[1], R) / N_A

    sd = SpeciesData([f"S{i}" for i in range(S)])
    rd = RxData(id_reacs, id_prods, st_reacs, st_prods)
    return sd, rd, Ea, A


def synthetic_u0(S: int) -> np.ndarray:
    """Species 8..17 start at 0.1, everything else at 0."""
    u0 = np.zeros(S)
    u0[8:18] = 0.1
    return u0


# ----------------------------------------------------------------------------
# Stand-in for the docs' getting-started methane CRN.  The real network is built
# by the external CDE binary at docs-build time and is NOT shipped;
# only its 30 Ea/A values are (examples/getting_started/arrhenius_params.bson).
# This hand-written 15-reversible-reaction methane pyrolysis network has the
# same size (30 reactions: forward block + reverse block) and is labelled a
# stand-in everywhere it is used.
# ----------------------------------------------------------------------------
GETTING_STARTED_SPECIES = ["C", "[CH3]", "[H]", "[H][H]", "CC", "C[CH2]", "C=C",
                           "C=[CH]", "C#C", "[CH2]"]

_GS_FWD = [
    (["C"], ["[CH3]", "[H]"]),
    (["C", "[H]"], ["[CH3]", "[H][H]"]),
    (["[CH3]", "[CH3]"], ["CC"]),
    (["CC"], ["C[CH2]", "[H]"]),
    (["CC", "[H]"], ["C[CH2]", "[H][H]"]),
    (["CC", "[CH3]"], ["C[CH2]", "C"]),
    (["C[CH2]"], ["C=C", "[H]"]),
    (["C=C", "[H]"], ["C=[CH]", "[H][H]"]),
    (["C=C", "[CH3]"], ["C=[CH]", "C"]),
    (["C=[CH]"], ["C#C", "[H]"]),
    (["[H]", "[H]"], ["[H][H]"]),
    (["[CH3]"], ["[CH2]", "[H]"]),
    (["[CH2]", "C"], ["[CH3]", "[CH3]"]),
    (["C=C"], ["C#C", "[H][H]"]),
    (["[CH2]", "[CH2]"], ["C=C"]),
]


def getting_started_standin():
    """Return (sd, rd) of the stand-in getting-started CRN (30 reactions, 10 species)."""
    sd = SpeciesData(list(GETTING_STARTED_SPECIES))

    def enc(names):
        ids = sorted(sd.toInt[n] for n in names)
        out_i, out_n = [], []
        for s in ids:
            if out_i and out_i[-1] == s:
                out_n[-1] += 1
            else:
                out_i.append(s); out_n.append(1)
        return out_i, out_n

    ir, ip, sr, sp = [], [], [], []
    for reac, prod in _GS_FWD:
        a, b = enc(reac); ir.append(a); sr.append(b)
        a, b = enc(prod); ip.append(a); sp.append(b)
    for reac, prod in _GS_FWD:
        a, b = enc(prod); ir.append(a); sr.append(b)
        a, b = enc(reac); ip.append(a); sp.append(b)
    return sd, RxData(ir, ip, sr, sp)
