"""Input schema of the kinetic solve: `SpeciesData` and `RxData`.

Mirrors the fields the reference's solve path reads
(reference src/exploration/network.jl:1-8 `SpeciesData`, :193-203 `RxData`,
:514-529 `splice!`).  Only the id/stoichiometry arrays matter to the solver;
chemistry-side fields (xyz, hashes, dH, levels) are carried opaquely.

Indices are 0-based on this side of the boundary (the Julia shim converts its
1-based ids, see julia/KineticaB200.jl).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Sequence

import numpy as np


class SpeciesData:
    """Species table: `n`, `toInt` (SMILES -> id), `toStr` (id -> SMILES)."""

    def __init__(self, smiles: Sequence[str]):
        self.toStr: Dict[int, str] = {i: s for i, s in enumerate(smiles)}
        self.toInt: Dict[str, int] = {s: i for i, s in enumerate(smiles)}
        if len(self.toInt) != len(self.toStr):
            raise ValueError("duplicate species in SpeciesData")
        self.n = len(smiles)


@dataclass
class RxData:
    """Reaction table (ragged, one entry per reaction)."""
    id_reacs: List[List[int]]
    id_prods: List[List[int]]
    stoic_reacs: List[List[int]]
    stoic_prods: List[List[int]]
    mapped_rxns: List[str] = field(default_factory=list)
    dH: List[float] = field(default_factory=list)
    rhash: List[bytes] = field(default_factory=list)
    level_found: List[int] = field(default_factory=list)

    def __post_init__(self):
        n = len(self.id_reacs)
        if not (len(self.id_prods) == len(self.stoic_reacs) == len(self.stoic_prods) == n):
            raise ValueError("RxData arrays disagree in length")
        for name in ("mapped_rxns", "dH", "rhash", "level_found"):
            v = getattr(self, name)
            if len(v) not in (0, n):
                raise ValueError(f"RxData.{name} has wrong length")

    @property
    def nr(self) -> int:
        return len(self.id_reacs)

    def splice(self, rids: Sequence[int]) -> None:
        """Remove reactions `rids` from every field, compacting in order
        (reference src/exploration/network.jl:514-529)."""
        if len(rids) == 0:
            return
        drop = set(int(r) for r in rids)
        keep = [i for i in range(self.nr) if i not in drop]
        for name in ("id_reacs", "id_prods", "stoic_reacs", "stoic_prods",
                     "mapped_rxns", "dH", "rhash", "level_found"):
            v = getattr(self, name)
            if len(v):
                setattr(self, name, [v[i] for i in keep])

    def flatten(self):
        """Ragged -> CSR: (reac_ptr, reac_idx, reac_nu, prod_ptr, prod_idx, prod_nu), int64."""
        def csr(ids, nus):
            ptr = np.zeros(len(ids) + 1, dtype=np.int64)
            for j, row in enumerate(ids):
                ptr[j + 1] = ptr[j] + len(row)
            idx = np.fromiter((s for row in ids for s in row), dtype=np.int64, count=int(ptr[-1]))
            nu = np.fromiter((s for row in nus for s in row), dtype=np.int64, count=int(ptr[-1]))
            return ptr, idx, nu
        return csr(self.id_reacs, self.stoic_reacs) + csr(self.id_prods, self.stoic_prods)
