"""CPU ORACLE — test infrastructure only, never the product path.

A numpy/scipy restatement of the kinetic-solve hot path of Kinetica.jl v0.7.2
(`solve_network(StaticODESolve|VariableODESolve, sd, rd)` with discrete rate
updates).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.

PARITY STATUS: **parity unpinned** at the solver boundary.  The reference's own
tests never call `solve_network`, a calculator functor, `make_rs` or any ODE
solve, Julia is not installed here, and the arithmetic of RHS /
Jacobian / implicit step lives in un-vendored Julia packages (Catalyst 14.4,
ModelingToolkit 9, OrdinaryDiffEq 6.95; reference Project.toml:34-58).  What IS
pinned against the reference:
  * condition-profile known answers of reference test/Main/conditions.jl:4-135
    (tests/test_oracle_golden.py ports them 1:1),
  * the shipped Ea/A fixture examples/getting_started/arrhenius_params.bson
    (tests/golden/arrhenius_params.json, made by tests/golden/make_golden.py).
Everything else is pinned by closed-form known-answer CRNs, conservation laws
and agreement between two independent implicit integrators (Radau, BDF).

Every function cites the reference file:line whose behaviour it restates.
Indices are 0-based here; the reference is 1-based.
"""
from __future__ import annotations

import itertools
import math
from fractions import Fraction

import numpy as np

# reference src/constants.jl:4-5
R_GAS = 8.314462618
N_A = 6.02214076e23

# reference src/utils.jl:77-97
T_UNIT_MAP = {
    "picoseconds": 1.0e-12, "ps": 1.0e-12, "nanoseconds": 1.0e-9, "ns": 1.0e-9,
    "microseconds": 1.0e-6, "us": 1.0e-6, "milliseconds": 1.0e-3, "ms": 1.0e-3,
    "seconds": 1.0, "s": 1.0, "minutes": 60.0, "mins": 60.0, "hours": 3600.0,
    "hrs": 3600.0, "days": 86400.0, "months": 2.6297368e06, "mts": 2.6297368e06,
    "years": 3.15576e07, "yrs": 3.15576e07,
}


def tconvert(t, from_unit, to_unit):
    """reference src/utils.jl:21-42"""
    if from_unit not in T_UNIT_MAP or to_unit not in T_UNIT_MAP:
        raise ValueError("Unknown unit specified in time conversion!")
    return float(t) * T_UNIT_MAP[from_unit] / T_UNIT_MAP[to_unit]


# ----------------------------------------------------------------------------
# Julia float ranges, restated with exact rationals  [upstream Base semantics]
# ----------------------------------------------------------------------------
def _simplest_rational(x: float) -> Fraction:
    """Rational with the shortest decimal that round-trips to `x` (what
    Base.rat's continued-fraction search finds for decimal literals)."""
    return Fraction(repr(float(x)))


def julia_range(start: float, step: float, stop: float) -> np.ndarray:
    """`collect(start:step:stop)` for Float64.  Julia lifts start/step/stop to
    rationals so every element is the correctly rounded value of
    start + i*step and the length is floor((stop-start)/step)+1 evaluated
    exactly (Base `range_start_step_stop`, twice-precision fallback) —
    SURVEY.md §7 "Julia range semantics"."""
    if step == 0:
        raise ValueError("step cannot be zero")
    a, s, b = (_simplest_rational(v) for v in (start, step, stop))
    if float(a) != start or float(s) != step or float(b) != stop:
        a, s, b = Fraction(start), Fraction(step), Fraction(stop)
    n = math.floor((b - a) / s)
    if n < 0:
        return np.zeros(0)
    return np.array([float(a + i * s) for i in range(n + 1)], dtype=np.float64)


def create_savepoints(start: float, stop: float, step: float) -> np.ndarray:
    """reference src/utils.jl:108-115"""
    if step > 1e-9 and abs(step - math.floor(step)) < 1e-9:
        cstep = float(f"{step:.9g}")           # round(step; sigdigits=9)
    else:
        cstep = step
    r = julia_range(start, cstep, stop)
    if r[-1] < stop:
        r = np.append(r, stop)
    return r


class LinearInterp:
    """`DiffEqArray(u, t)(t_interp)` — linear interpolation functor
    (reference src/utils.jl:135-139 on SciMLBase.LinearInterpolation)."""

    def __init__(self, t, u):
        self.t = np.asarray(t, dtype=np.float64)
        self.u = np.asarray(u, dtype=np.float64)

    def __call__(self, tq):
        t, u = self.t, self.u
        if tq <= t[0]:
            return u[0] if tq == t[0] else u[0] + (u[1] - u[0]) * (tq - t[0]) / (t[1] - t[0])
        i = int(np.searchsorted(t, tq, side="left"))   # continuity=:left
        if i >= len(t):
            i = len(t) - 1
        if t[i] == tq:
            return u[i]
        i0 = i - 1
        th = (tq - t[i0]) / (t[i] - t[i0])
        return (1.0 - th) * u[i0] + th * u[i]


# ----------------------------------------------------------------------------
# Condition profiles   (reference src/conditions/*.jl)
# ----------------------------------------------------------------------------
class StaticConditionProfile:
    """reference src/conditions/static.jl:7-9"""
    static = True

    def __init__(self, value):
        self.value = float(value)


class _Variable:
    static = False
    sol = None

    def minimum(self):
        """reference src/conditions/abstract_profiles.jl:113-118"""
        if self.sol is None:
            raise RuntimeError("Condition profile is missing a solution.")
        return float(np.min(self.sol.u))

    def maximum(self):
        """reference src/conditions/abstract_profiles.jl:134-139"""
        if self.sol is None:
            raise RuntimeError("Condition profile is missing a solution.")
        return float(np.max(self.sol.u))


class _Direct(_Variable):
    def solve(self, tspan, save_interval):
        """reference src/conditions/direct_variable.jl:34-43"""
        si = tspan[1] / 1000 if save_interval is None else save_interval
        t = create_savepoints(tspan[0], tspan[1], si)
        self.sol = LinearInterp(t, [self.f(tp) for tp in t])


class _Gradient(_Variable):
    def _kinks(self):
        return []

    def X(self, t):
        raise NotImplementedError

    def solve(self, tspan, save_interval):
        """reference src/conditions/gradient_variable.jl:35-64.  The reference
        integrates D(X) ~ grad(t) with OwrenZen5 (abstol 1e-6, reltol 1e-4)
        stopping at profile.tstops and saving at savepoints ∪ tstops; every
        shipped gradient is piecewise constant or piecewise linear with its
        kinks inside tstops, which a 5th-order RK integrates exactly, so the
        closed-form integral X(t) is used here."""
        si = tspan[1] / 1000 if save_interval is None else save_interval
        t = np.sort(np.concatenate([create_savepoints(tspan[0], tspan[1], si),
                                    np.asarray(self.tstops, dtype=np.float64)]))
        t = t[(t >= tspan[0]) & (t <= tspan[1])]
        self.sol = LinearInterp(t, [self.X(tp) for tp in t])


class NullDirectProfile(_Direct):
    """reference src/conditions/direct_variable.jl:49-92"""

    def __init__(self, X_start, t_end):
        self.X_start, self.t_end = float(X_start), float(t_end)
        self.tstops = np.array([self.t_end])

    def f(self, t):
        return self.X_start

    def create_discrete_tstops(self, ts_update):
        if ts_update > self.t_end:
            raise ValueError("Error defining tstops, `ts_update` is too large.")
        self.tstops = julia_range(0.0, ts_update, self.t_end)


class LinearDirectProfile(_Direct):
    """reference src/conditions/direct_variable.jl:98-155"""

    def __init__(self, rate, X_start, X_end):
        rate, X_start, X_end = float(rate), float(X_start), float(X_end)
        if (X_end < X_start and rate > 0) or (X_end > X_start and rate < 0):
            raise RuntimeError("Impossible temperature ramp defined. Check heating rates have the correct signs.")
        self.rate, self.X_start, self.X_end = rate, X_start, X_end
        self.t_end = (X_end - X_start) / rate
        self.tstops = np.array([self.t_end])

    def f(self, t):
        # reference :144-150 — sum of masked branches
        return ((t <= 0.0) * self.X_start
                + ((t > 0.0 and t <= self.t_end) * (self.X_start + self.rate * t))
                + ((t > self.t_end) * self.X_end))

    def create_discrete_tstops(self, ts_update):
        if ts_update > self.t_end:
            raise ValueError("Error defining tstops, `ts_update` is too large.")
        self.tstops = create_savepoints(0.0, self.t_end, ts_update)


class NullGradientProfile(_Gradient):
    """reference src/conditions/gradient_variable.jl:70-114"""

    def __init__(self, X_start, t_end):
        self.X_start, self.t_end = float(X_start), float(t_end)
        self.tstops = np.array([self.t_end])

    def grad(self, t):
        return 0.0

    def X(self, t):
        return self.X_start

    def create_discrete_tstops(self, ts_update):
        if ts_update > self.t_end:
            raise ValueError("Error defining tstops, `ts_update` is too large.")
        self.tstops = julia_range(0.0, ts_update, self.t_end)


class LinearGradientProfile(_Gradient):
    """reference src/conditions/gradient_variable.jl:120-175"""

    def __init__(self, rate, X_start, X_end):
        rate, X_start, X_end = float(rate), float(X_start), float(X_end)
        if (X_end < X_start and rate > 0) or (X_end > X_start and rate < 0):
            raise RuntimeError("Impossible condition ramp defined. Check heating rates have the correct signs.")
        self.rate, self.X_start, self.X_end = rate, X_start, X_end
        self.t_end = (X_end - X_start) / rate
        self.tstops = np.array([self.t_end])

    def grad(self, t):
        # reference :165-170 — `rate` for ALL t <= t_end (negative t included)
        return (t <= self.t_end) * self.rate + (t > self.t_end) * 0.0

    def X(self, t):
        return self.X_start + self.rate * min(t, self.t_end)

    def create_discrete_tstops(self, ts_update):
        if ts_update > self.t_end:
            raise ValueError("Error defining tstops, `ts_update` is too large.")
        self.tstops = create_savepoints(0.0, self.t_end, ts_update)


class DoubleRampGradientProfile(_Gradient):
    """reference src/conditions/gradient_variable.jl:181-310"""

    def __init__(self, X_start, t_start_plateau, rate1, X_mid, t_mid_plateau, rate2, X_end,
                 t_end_plateau, t_blend=None):
        X_start, X_mid, X_end = float(X_start), float(X_mid), float(X_end)
        rate1, rate2 = float(rate1), float(rate2)
        if ((X_mid > X_start and rate1 < 0) or (X_mid < X_start and rate1 > 0)
                or (X_end > X_mid and rate2 < 0) or (X_end < X_mid and rate2 > 0)):
            raise RuntimeError("Impossible condition ramp defined. Check heating rates have the correct signs.")
        self.rate1, self.rate2 = rate1, rate2
        self.X_start, self.X_mid, self.X_end = X_start, X_mid, X_end
        self.t_start_plateau = float(t_start_plateau)
        self.t_mid_plateau = float(t_mid_plateau)
        self.t_end_plateau = float(t_end_plateau)
        self.t_startr1 = self.t_start_plateau
        self.t_endr1 = self.t_startr1 + ((X_mid - X_start) / rate1)
        self.t_startr2 = self.t_endr1 + self.t_mid_plateau
        self.t_endr2 = self.t_startr2 + ((X_end - X_mid) / rate2)
        self.t_end = self.t_endr2 + self.t_end_plateau
        if t_blend is None:
            self.blended = False
            self.t_blend = 0.0
            self.tstops = np.array([self.t_startr1, self.t_endr1, self.t_startr2, self.t_endr2, self.t_end])
        else:
            self.blended = True
            tb = self.t_blend = float(t_blend)
            self.tstops = np.array([
                self.t_startr1 - tb, self.t_startr1 + tb, self.t_endr1 - tb, self.t_endr1 + tb,
                self.t_startr2 - tb, self.t_startr2 + tb, self.t_endr2 - tb, self.t_endr2 + tb,
                self.t_end])

    def grad(self, t):
        p = self
        if not p.blended:       # reference :277-285
            return (((t < p.t_startr1) * 0.0)
                    + ((t >= p.t_startr1 and t < p.t_endr1) * p.rate1)
                    + ((t >= p.t_endr1 and t < p.t_startr2) * 0.0)
                    + ((t >= p.t_startr2 and t < p.t_endr2) * p.rate2)
                    + ((t >= p.t_endr2) * 0.0))
        tb = p.t_blend          # reference :287-299
        return (((t < p.t_startr1 - tb) * 0.0)
                + ((t >= p.t_startr1 - tb and t < p.t_startr1 + tb) * (p.rate1 * (t - p.t_startr1 - tb) / (2 * tb) + p.rate1))
                + ((t >= p.t_startr1 + tb and t < p.t_endr1 - tb) * p.rate1)
                + ((t >= p.t_endr1 - tb and t < p.t_endr1 + tb) * (-p.rate1 * (t - p.t_endr1 - tb) / (2 * tb)))
                + ((t >= p.t_endr1 + tb and t < p.t_startr2 - tb) * 0.0)
                + ((t >= p.t_startr2 - tb and t < p.t_startr2 + tb) * (p.rate2 * (t - p.t_startr2 - tb) / (2 * tb) + p.rate2))
                + ((t >= p.t_startr2 + tb and t < p.t_endr2 - tb) * p.rate2)
                + ((t >= p.t_endr2 - tb and t < p.t_endr2 + tb) * (-p.rate2 * (t - p.t_endr2 - tb) / (2 * tb)))
                + ((t >= p.t_endr2 + tb) * 0.0))

    def _segments(self):
        """Breakpoints of grad(t) and its (value at left end, slope) per segment."""
        p, tb = self, self.t_blend
        if not p.blended:
            return [(-math.inf, 0.0, 0.0), (p.t_startr1, p.rate1, 0.0), (p.t_endr1, 0.0, 0.0),
                    (p.t_startr2, p.rate2, 0.0), (p.t_endr2, 0.0, 0.0)]
        segs = [(-math.inf, 0.0, 0.0)]
        for ts, te, r in ((p.t_startr1, p.t_endr1, p.rate1), (p.t_startr2, p.t_endr2, p.rate2)):
            segs.append((ts - tb, 0.0, r / (2 * tb)))
            segs.append((ts + tb, r, 0.0))
            segs.append((te - tb, r, -r / (2 * tb)))
            segs.append((te + tb, 0.0, 0.0))
        return segs

    def X(self, t):
        """Closed-form integral of grad from 0 (X(0) = X_start)."""
        segs = self._segments()
        starts = [s[0] for s in segs] + [math.inf]

        def integral_to(tt):
            acc = 0.0
            for (a, g0, sl), b in zip(segs, starts[1:]):
                lo = max(a, 0.0) if a != -math.inf else 0.0
                hi = min(b, tt)
                if hi <= lo:
                    continue
                aa = lo if a == -math.inf else a
                acc += g0 * (hi - lo) + 0.5 * sl * ((hi - aa) ** 2 - (lo - aa) ** 2)
            return acc
        return self.X_start + (integral_to(t) if t > 0 else 0.0)

    def create_discrete_tstops(self, ts_update):
        """reference :301-310"""
        if ts_update > self.t_end:
            raise ValueError("Error defining tstops, `ts_update` is too large.")
        tb = self.t_blend
        self.tstops = np.concatenate([
            [0.0],
            create_savepoints(self.t_startr1 - tb, self.t_endr1 + tb, ts_update),
            create_savepoints(self.t_startr2 - tb, self.t_endr2 + tb, ts_update),
            [self.t_end]])


class ConditionSet:
    """reference src/conditions/condition_set.jl:1-58 (+ accessors :61-191)"""

    def __init__(self, d, ts_update=None):
        self.symbols, self.profiles = [], []
        for sym, v in d.items():
            if isinstance(v, (int, float)) and not isinstance(v, bool):
                self.profiles.append(StaticConditionProfile(v))
            elif isinstance(v, (StaticConditionProfile, _Variable)):
                if ts_update is not None and not v.static:
                    v.create_discrete_tstops(ts_update)
                self.profiles.append(v)
            else:
                raise ValueError(f"Condition {sym} does not have a valid profile.")
            self.symbols.append(sym)
        self.discrete_updates = ts_update is not None
        self.ts_update = ts_update

    def isstatic(self):
        return all(p.static for p in self.profiles)

    def get_profile(self, sym):
        if sym not in self.symbols:
            raise KeyError(f"Condition {sym} does not exist in this ConditionSet")
        return self.profiles[self.symbols.index(sym)]

    def get_initial_conditions(self):
        """reference :111-121"""
        return {s: (p.value if p.static else p.X_start) for s, p in zip(self.symbols, self.profiles)}

    def get_static_conditions(self):
        return {s: p.value for s, p in zip(self.symbols, self.profiles) if p.static}

    def get_variable_conditions(self):
        return {s: p.sol for s, p in zip(self.symbols, self.profiles) if not p.static}

    def get_tstops(self):
        """reference :172-176 — sort(unique(vcat(...)))"""
        if self.isstatic():
            raise RuntimeError("No tstops available, all conditions in ConditionSet are static.")
        return np.unique(np.concatenate([np.asarray(p.tstops, dtype=np.float64)
                                         for p in self.profiles if not p.static]))

    def get_t_final(self):
        """reference :187-191"""
        if self.isstatic():
            raise RuntimeError("No t_end available, all conditions in ConditionSet are static.")
        return max(p.t_end for p in self.profiles if not p.static)

    def solve_variable_conditions(self, tspan, save_interval=None):
        """reference :260-268"""
        for p in self.profiles:
            if not p.static:
                p.solve(tspan, save_interval)


# ----------------------------------------------------------------------------
# Calculators   (reference src/solving/calculator.jl)
# ----------------------------------------------------------------------------
class PrecalculatedArrheniusCalculator:
    """reference src/solving/calculator.jl:164-238"""

    def __init__(self, Ea, A, k_max=None, t_unit="s"):
        self.Ea = np.array(Ea, dtype=np.float64)
        self.A = np.array(A, dtype=np.float64)
        self.k_max = k_max
        self.t_unit = t_unit
        self.t_mult = tconvert(1.0, t_unit, "s")

    def __call__(self, T):
        # reference :223-232 — this exact operation order
        with np.errstate(divide="ignore", over="ignore"):
            k_r = self.A * np.exp(-self.Ea / (R_GAS * T)) * N_A * self.t_mult
            if self.k_max is None:
                return k_r
            return 1.0 / ((1.0 / self.k_max) + (1.0 / k_r))

    def splice(self, rids):
        keep = np.setdiff1d(np.arange(len(self.Ea)), np.asarray(rids, dtype=np.int64))
        self.Ea, self.A = self.Ea[keep], self.A[keep]


class DummyKineticCalculator:
    """reference src/solving/calculator.jl:72-158"""

    def __init__(self, rates, k_max=None, t_unit="s"):
        self.rates = np.array(rates, dtype=np.float64)
        self.k_max = k_max
        self.t_mult = tconvert(1.0, t_unit, "s")

    def __call__(self, T=None, **_):
        k_r = self.rates * self.t_mult
        if self.k_max is None:
            return k_r
        return 1.0 / ((1.0 / self.k_max) + (1.0 / k_r))


def get_max_rates(conditions: ConditionSet, calc):
    """reference src/solving/solve_utils.jl:19-54 (single-condition :T calculators)"""
    static = conditions.get_static_conditions()
    variable = [(s, p) for s, p in zip(conditions.symbols, conditions.profiles) if not p.static]
    if not variable:
        return calc(**static)
    perms = []
    for bits in itertools.product((0, 1), repeat=len(variable)):     # '00','01',... like lpad(string(i, base=2))
        kw = dict(static)
        for (s, p), b in zip(variable, bits):
            kw[s] = p.maximum() if b else p.minimum()
        perms.append(calc(**kw))
    means = [float(np.mean(p)) for p in perms]
    return perms[int(np.argmax(means))]                              # findmax: first maximum


def low_k_removal_set(max_rates, reltol, t_final, low_k_cutoff="auto", low_k_maxconc=2.0):
    """reference src/solving/solve_utils.jl:213-245 — indices (0-based) removed."""
    if low_k_cutoff == "none":
        return np.zeros(0, dtype=np.int64)
    cutoff = reltol / t_final if low_k_cutoff == "auto" else float(low_k_cutoff)
    scaled = np.asarray(max_rates) * low_k_maxconc ** 2
    return np.nonzero(scaled < cutoff)[0].astype(np.int64)


def calculate_discrete_rates(conditions: ConditionSet, calc):
    """reference src/solving/solve_utils.jl:91-109 — returns (tstops, k[Nt, R]).
    The variable condition is read off the *interpolated profile solution*
    (`vpair.second(tstop)[1]`, :101-104)."""
    if not conditions.discrete_updates:
        raise RuntimeError("Cannot calculate discrete rates for a continuous ConditionSet.")
    tstops = conditions.get_tstops()
    static = conditions.get_static_conditions()
    vcs = conditions.get_variable_conditions()
    out = []
    for ts in tstops:
        kw = dict(static)
        for s, sol in vcs.items():
            kw[s] = float(sol(ts))
        out.append(calc(**kw))
    return tstops, np.array(out)


# ----------------------------------------------------------------------------
# Mass action (reference src/solving/solve_utils.jl:318-334 `make_rs`, with
# Catalyst's non-combinatoric rate law [upstream])
# ----------------------------------------------------------------------------
class Network:
    """Flattened reaction table: 0-based CSR of reactants and products."""

    def __init__(self, S, id_reacs, id_prods, stoic_reacs, stoic_prods):
        self.S = int(S)
        self.R = len(id_reacs)
        self.id_reacs = [list(map(int, r)) for r in id_reacs]
        self.id_prods = [list(map(int, r)) for r in id_prods]
        self.stoic_reacs = [list(map(int, r)) for r in stoic_reacs]
        self.stoic_prods = [list(map(int, r)) for r in stoic_prods]
        # substrate exponents (merged if a species is listed twice) and net stoichiometry
        self.sub = []
        self.net = []
        for j in range(self.R):
            sub = {}
            for s, n in zip(self.id_reacs[j], self.stoic_reacs[j]):
                sub[s] = sub.get(s, 0) + n
            net = {}
            for s, n in zip(self.id_prods[j], self.stoic_prods[j]):
                net[s] = net.get(s, 0) + n
            for s, n in sub.items():
                net[s] = net.get(s, 0) - n
            self.sub.append(sorted(sub.items()))
            self.net.append(sorted((s, n) for s, n in net.items() if n != 0))

    def _vec(self):
        """numpy/scipy tables for vectorised evaluation (same arithmetic, different summation order
        than the scalar loops below)."""
        if getattr(self, "_v", None) is None:
            import scipy.sparse as sp
            idx = np.full((self.R, 3), -1, dtype=np.int64)
            ex = np.zeros((self.R, 3), dtype=np.int64)
            for j, sub in enumerate(self.sub):
                assert len(sub) <= 3
                for q, (s_, n_) in enumerate(sub):
                    idx[j, q], ex[j, q] = s_, n_
            rows = [s_ for j in range(self.R) for s_, _ in self.net[j]]
            cols = [j for j in range(self.R) for _ in self.net[j]]
            vals = [float(n_) for j in range(self.R) for _, n_ in self.net[j]]
            N = sp.csr_matrix((vals, (rows, cols)), shape=(self.S, self.R))
            self._v = (idx, ex, N)
        return self._v

    def rates(self, u, k):
        idx, ex, _ = self._vec()
        u = np.asarray(u, dtype=np.float64)
        f = np.where(idx >= 0, u[np.maximum(idx, 0)] ** ex, 1.0)
        return np.asarray(k, dtype=np.float64) * f[:, 0] * f[:, 1] * f[:, 2]

    def rhs(self, u, k):
        """du_i = sum_j net[i,j] * k_j * prod_m u_m^nu_mj."""
        return self._vec()[2] @ self.rates(u, k)

    def rhs_scalar(self, u, k):
        """Same, scalar loops, summed in ascending j (the order the CUDA gather uses)."""
        du = np.zeros(self.S)
        for j in range(self.R):
            r = k[j]
            for s_, n_ in self.sub[j]:
                r *= u[s_] ** n_
            for s_, n_ in self.net[j]:
                du[s_] += n_ * r
        return du

    def jac_sparse(self, u, k):
        import scipy.sparse as sp
        idx, ex, N = self._vec()
        u = np.asarray(u, dtype=np.float64)
        k = np.asarray(k, dtype=np.float64)
        with np.errstate(divide="ignore", invalid="ignore"):
            f = np.where(idx >= 0, u[np.maximum(idx, 0)] ** ex, 1.0)
            rr, cc, vv = [], [], []
            for q in range(3):
                has = idx[:, q] >= 0
                e = ex[:, q]
                dq = np.where(e >= 1, e * u[np.maximum(idx[:, q], 0)] ** np.maximum(e - 1, 0), 0.0)
                others = np.ones(self.R)
                for q2 in range(3):
                    if q2 != q:
                        others = others * f[:, q2]
                d = k * dq * others
                rr.append(np.nonzero(has)[0]); cc.append(idx[has, q]); vv.append(d[has])
        D = sp.csr_matrix((np.concatenate(vv), (np.concatenate(rr), np.concatenate(cc))), shape=(self.R, self.S))
        return (N @ D).tocsc()

    def jac_dense(self, u, k):
        return np.asarray(self.jac_sparse(u, k).todense())

    def pattern_csc(self):
        """P_J = {(i,l): exists j, net[i,j] != 0 and nu_lj > 0}, CSC with rows ascending
        (SURVEY.md §8a R5 pattern contract)."""
        cols = [set() for _ in range(self.S)]
        for j in range(self.R):
            for l, _ in self.sub[j]:
                for i, _ in self.net[j]:
                    cols[l].add(i)
        colptr = np.zeros(self.S + 1, dtype=np.int64)
        rows = []
        for l in range(self.S):
            rr = sorted(cols[l])
            rows.extend(rr)
            colptr[l + 1] = colptr[l] + len(rr)
        return colptr, np.array(rows, dtype=np.int64)

    def conservation_basis(self):
        """Left null space of the net stoichiometric matrix (conservation laws)."""
        N = np.zeros((self.S, self.R))
        for j in range(self.R):
            for s, n in self.net[j]:
                N[s, j] = n
        u_, sv, vt = np.linalg.svd(N.T, full_matrices=True)
        rank = int(np.sum(sv > 1e-10 * max(sv.max(), 1.0))) if sv.size else 0
        return vt[rank:]


# ----------------------------------------------------------------------------
# Symbolic analysis restatement (bit-exact contract with kb2_symbolic)
# ----------------------------------------------------------------------------
def min_degree_order(S, colptr, rowval):
    """Minimum-degree ordering on the symmetrised pattern with an explicit
    elimination graph; ties -> smallest species index.  perm[k] = species
    eliminated k-th."""
    adj = [set() for _ in range(S)]
    for l in range(S):
        for p in range(colptr[l], colptr[l + 1]):
            i = int(rowval[p])
            if i != l:
                adj[i].add(l)
                adj[l].add(i)
    alive = [True] * S
    perm = []
    import heapq
    heap = [(len(adj[v]), v) for v in range(S)]
    heapq.heapify(heap)
    while heap:
        d, v = heapq.heappop(heap)
        if not alive[v] or d != len(adj[v]):
            continue
        alive[v] = False
        perm.append(v)
        nb = sorted(adj[v])
        for a in nb:
            adj[a].discard(v)
        for a in nb:
            before = len(adj[a])
            adj[a].update(x for x in nb if x != a)
            heapq.heappush(heap, (len(adj[a]), a))
        adj[v] = set()
    return np.array(perm, dtype=np.int64)


def symbolic_lu(S, colptr, rowval, perm):
    """Row-wise symbolic LU (no pivoting) of P (P_J ∪ diag) P^T.
    Returns (rowptr, colidx, diagpos, n_fma) of the combined L\\U pattern,
    columns ascending inside each row."""
    iperm = np.empty(S, dtype=np.int64)
    iperm[perm] = np.arange(S)
    rows = [set([a]) for a in range(S)]
    for l in range(S):
        for p in range(colptr[l], colptr[l + 1]):
            rows[iperm[int(rowval[p])]].add(int(iperm[l]))
    upper = [None] * S
    rowptr = np.zeros(S + 1, dtype=np.int64)
    colidx = []
    diagpos = np.zeros(S, dtype=np.int64)
    n_fma = 0
    import heapq
    for i in range(S):
        pat = rows[i]
        heap = [c for c in pat if c < i]
        heapq.heapify(heap)
        done = set()
        while heap:
            k = heapq.heappop(heap)
            if k in done:
                continue
            done.add(k)
            n_fma += len(upper[k])
            for j in upper[k]:
                if j not in pat:
                    pat.add(j)
                    if j < i:
                        heapq.heappush(heap, j)
        srt = sorted(pat)
        upper[i] = [c for c in srt if c > i]
        diagpos[i] = rowptr[i] + srt.index(i)
        colidx.extend(srt)
        rowptr[i + 1] = rowptr[i] + len(srt)
    return rowptr, np.array(colidx, dtype=np.int64), diagpos, n_fma


# ----------------------------------------------------------------------------
# Trajectory oracle: independent high-accuracy implicit integration
# ----------------------------------------------------------------------------
def solve_trajectory(net: Network, u0, k_table, tstops, tspan, saveat, k_init=None,
                     method="Radau", rtol=1e-10, atol=1e-14):
    """Integrate du/dt = rhs(u, k(t)) with zero-order-hold rate constants:
    k = k_init on [t0, tstops[0]) and k = k_table[i] on [tstops[i], tstops[i+1])
    (reference src/solving/solve_utils.jl:445-450 `CompleteRateUpdateAffect`
    driven by PresetTimeCallback, src/solving/methods.jl:678-679; initial k
    src/solving/methods.jl:668).  Restarted at every tstop.  Returns u[Ns, S]."""
    from scipy.integrate import solve_ivp
    import scipy.sparse as sp

    t0, tf = float(tspan[0]), float(tspan[1])
    saveat = np.asarray(saveat, dtype=np.float64)
    tstops = np.asarray(tstops, dtype=np.float64) if tstops is not None else np.zeros(0)
    k_table = np.asarray(k_table, dtype=np.float64)
    if k_table.ndim == 1:
        k_table = k_table[None, :]
    colptr, rowval = net.pattern_csc()
    have_pat = len(rowval) > 0
    brk = [t0] + [float(t) for t in tstops if t0 < t < tf] + [tf]
    u = np.array(u0, dtype=np.float64).copy()
    out = np.zeros((len(saveat), net.S))
    k = np.array(k_init if k_init is not None else k_table[0], dtype=np.float64)
    for a, b in zip(brk[:-1], brk[1:]):
        hit = np.nonzero(tstops == a)[0]
        if hit.size:
            k = k_table[int(hit[0])]
        sel = np.nonzero((saveat >= a) & ((saveat < b) | ((b == tf) & (saveat <= b))))[0]
        kk = k

        def f(t, y, kk=kk):
            return net.rhs(y, kk)

        def jac(t, y, kk=kk):
            return net.jac_sparse(y, kk) if (have_pat and net.S > 64) else net.jac_dense(y, kk)

        te = np.unique(np.append(saveat[sel], b))
        sol = solve_ivp(f, (a, b), u, method=method, jac=jac, rtol=rtol, atol=atol, t_eval=te)
        if not sol.success:
            raise RuntimeError("oracle integration failed: " + sol.message)
        if sel.size:
            out[sel] = sol.y.T[np.searchsorted(te, saveat[sel])]
        u = sol.y[:, -1]
    return out


def solve_trajectory_continuous(net: Network, u0, k_of_t, breaks, tspan, saveat, method="Radau", rtol=1e-10, atol=1e-14):
    """Integrate du/dt = rhs(u, k(t)) with CONTINUOUS rate constants k(t) = calculator(X(t)) — the
    reference's continuous rate update mode (src/solving/methods.jl:363-458: k is an algebraic
    variable bound to the condition profile).  `breaks`: the profile's tstops (its kinks), where the
    integration is restarted like the reference's `tstops` keyword (:446).  Returns u[Ns, S]."""
    from scipy.integrate import solve_ivp

    t0, tf = float(tspan[0]), float(tspan[1])
    saveat = np.asarray(saveat, dtype=np.float64)
    colptr, rowval = net.pattern_csc()
    have_pat = len(rowval) > 0
    brk = [t0] + [float(t) for t in np.asarray(breaks, dtype=np.float64) if t0 < t < tf] + [tf]
    u = np.array(u0, dtype=np.float64).copy()
    out = np.zeros((len(saveat), net.S))

    def f(t, y):
        return net.rhs(y, k_of_t(t))

    def jac(t, y):
        kk = k_of_t(t)
        return net.jac_sparse(y, kk) if (have_pat and net.S > 64) else net.jac_dense(y, kk)

    for a, b in zip(brk[:-1], brk[1:]):
        sel = np.nonzero((saveat >= a) & ((saveat < b) | ((b == tf) & (saveat <= b))))[0]
        te = np.unique(np.append(saveat[sel], b))
        sol = solve_ivp(f, (a, b), u, method=method, jac=jac, rtol=rtol, atol=atol, t_eval=te)
        if not sol.success:
            raise RuntimeError("oracle integration failed: " + sol.message)
        if sel.size:
            out[sel] = sol.y.T[np.searchsorted(te, saveat[sel])]
        u = sol.y[:, -1]
    return out
