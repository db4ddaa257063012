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


K_B = 1.380649e-23        # J/K
H_PLANCK = 6.62607015e-34  # J s


class _ModifiedArrheniusForm(PrecalculatedArrheniusCalculator):
    """Calculators whose temperature dependence is A*T^n*exp(-E/RT): they run on the device through
    the T^n factor of kb2_set_arrhenius.  The device multiplies by N_A*t_mult like the
    reference's Arrhenius functor (calculator.jl:223-232), so prefactors are stored accordingly."""

    def __call__(self, *, T, **_other_conditions):
        return super().__call__(T=T)


class CollisionTheoryCalculator(_ModifiedArrheniusForm):
    """Rate law of the reference ecosystem's `KPMCollisionCalculator`
    (docs/src/tutorials/kinetic-calculators.md:129-150):
        k_i = 1 / (1/k_max + 1/(sigma_i rho_i N_A sqrt(8 k_B T / (pi mu_i)) exp(-E_i / RT)))
    with the activation energies E_i [J/mol], reduced masses mu_i [kg], collision cross-sections
    sigma_i [m^2] and steric factors rho_i supplied by the caller (in the reference they come from
    the external KPM model and the species' geometries, which are outside the hot path)."""

    def __init__(self, Ea, mu, sigma, rho=None, k_max=None, t_unit="s"):
        mu, sigma = np.asarray(mu, dtype=np.float64), np.asarray(sigma, dtype=np.float64)
        rho = np.ones_like(mu) if rho is None else np.asarray(rho, dtype=np.float64)
        A = sigma * rho * np.sqrt(8.0 * K_B / (np.pi * mu))
        super().__init__(Ea, A, k_max=k_max, t_unit=t_unit, n=np.full(len(A), 0.5))


class EyringCalculator(_ModifiedArrheniusForm):
    """Eyring equation of the reference's `ASENEBCalculator` (docs/src/tutorials/kinetic-calculators.md:73-77),
        k = k_B T / h * exp(dS/R) * exp(-dH/RT),
    for caller-supplied entropies [J/mol/K] and enthalpies [J/mol] of activation (in the reference
    they come from NEB + vibrational analysis, outside the hot path; taken as temperature-independent)."""

    def __init__(self, dH, dS, k_max=None, t_unit="s"):
        dH, dS = np.asarray(dH, dtype=np.float64), np.asarray(dS, dtype=np.float64)
        A = (K_B / H_PLANCK) * np.exp(dS / R_GAS) / N_A          # the device formula carries the N_A of the Arrhenius functor
        super().__init__(dH, A, k_max=k_max, t_unit=t_unit, n=np.ones(len(A)))

    def has_conditions(self, symbols):
        return all(s in ("T", "P") for s in symbols)
