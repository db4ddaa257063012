"""Host-side mirror of the reference calculator API (src/solving/calculator.jl).

A calculator is a functor `calc(T=..., P=...) -> k[R]`; conditions arrive as keyword
arguments and ALL R rate constants are returned (reference
docs/src/development/calculator-interface.md).  Calculators that can describe themselves
to the device (`device_arrhenius()`) have their k(T) evaluated by the CUDA kernel at every
rate-update stop; any other calculator is tabulated on the host with
`calculate_discrete_rates` and uploaded as a k table — exactly the reference's discrete
design.
"""
from __future__ import annotations

import numpy as np

from .conditions import tconvert

R_GAS = 8.314462618      # reference src/constants.jl:4
N_A = 6.02214076e23      # reference src/constants.jl:5


class AbstractKineticCalculator:
    """reference src/solving/calculator.jl:1-66"""

    def setup_network(self, sd, rd):          # setup_network!
        raise NotImplementedError

    def splice(self, rids):                    # Base.splice!(calc, rids)
        raise NotImplementedError

    def has_conditions(self, symbols):
        raise NotImplementedError

    def allows_continuous(self):
        return False

    def device_arrhenius(self):
        """Return dict(A, Ea, n, k_max, t_mult) if k(T) can be evaluated by the device kernel."""
        return None


class DummyKineticCalculator(AbstractKineticCalculator):
    """reference src/solving/calculator.jl:72-158"""

    def __init__(self, rates, k_max=None, t_unit="s"):
        self.rates = np.array(rates, dtype=np.float64)
        self.k_max = k_max
        self.t_unit = t_unit
        self.t_mult = tconvert(t_unit, "s")

    def setup_network(self, sd, rd):
        if len(self.rates) != rd.nr:
            raise ValueError(f"Number of rates ({len(self.rates)}) does not match number of reactions in `RxData` ({rd.nr})")

    def splice(self, rids):
        self.rates = np.delete(self.rates, np.asarray(rids, dtype=np.int64))

    def __call__(self, **conditions):
        k_r = self.rates * self.t_mult
        if self.k_max is None:
            return k_r
        return 1.0 / ((1.0 / self.k_max) + (1.0 / k_r))

    def has_conditions(self, symbols):
        return all(s in ("T", "V") for s in symbols)

    def allows_continuous(self):
        return True


class PrecalculatedArrheniusCalculator(AbstractKineticCalculator):
    """reference src/solving/calculator.jl:164-238.  `n` (T^n prefactor) is this build's
    extension and defaults to None = the reference formula."""

    def __init__(self, Ea, A, k_max=None, t_unit="s", n=None):
        self.Ea = np.array(Ea, dtype=np.float64)
        self.A = np.array(A, dtype=np.float64)
        self.n = None if n is None else np.array(n, dtype=np.float64)
        self.k_max = k_max
        self.t_unit = t_unit
        self.t_mult = tconvert(t_unit, "s")

    def setup_network(self, sd, rd):
        if len(self.Ea) != rd.nr or len(self.A) != rd.nr:
            raise ValueError(f"Number of parameters (Ea: {len(self.Ea)}, A: {len(self.A)}) does not match "
                             f"number of reactions in `RxData` ({rd.nr})")     # ArgumentError, :201-203

    def splice(self, rids):
        rids = np.asarray(rids, dtype=np.int64)
        self.Ea = np.delete(self.Ea, rids)
        self.A = np.delete(self.A, rids)
        if self.n is not None:
            self.n = np.delete(self.n, rids)

    def __call__(self, *, T):
        """Host evaluation (used for low-k pruning only; the solve evaluates k on the device)."""
        with np.errstate(divide="ignore", over="ignore"):
            k_r = self.A * np.exp(-self.Ea / (R_GAS * T))
            if self.n is not None:
                k_r = k_r * T ** self.n
            k_r = k_r * N_A * self.t_mult
            if self.k_max is None:
                return k_r
            return 1.0 / ((1.0 / self.k_max) + (1.0 / k_r))

    def has_conditions(self, symbols):
        return all(s in ("T",) for s in symbols)

    def allows_continuous(self):
        return True

    def device_arrhenius(self):
        return dict(A=self.A, Ea=self.Ea, n=self.n, k_max=self.k_max, t_mult=self.t_mult)
