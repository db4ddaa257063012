"""Host-side mirror of the reference's condition API (src/conditions/*.jl, src/utils.jl).

Same names, argument meaning and error behaviour as the reference; profiles also know how
to describe themselves to the device (`device_desc()` -> (kind, params[16])), where the
B200 path evaluates X_b(t) analytically instead of pre-solving each profile on the host.
"""
from __future__ import annotations

import math
from fractions import Fraction

import numpy as np

# device profile kinds (include/kinetica_b200.h)
KIND_STATIC, KIND_NULL, KIND_LINEAR_DIRECT, KIND_LINEAR_GRADIENT, KIND_DOUBLE_RAMP = 0, 1, 2, 3, 4

T_UNIT_MAP = {   # reference src/utils.jl:77-97
    "picoseconds": 1.0e-12, "ps": 1.0e-12, "nanoseconds": 1.0e-9, "ns": 1.0e-9,
    "microseconds": 1.0e-6, "us": 1.0e-6, "milliseconds": 1.0e-3, "ms": 1.0e-3,
    "seconds": 1.0, "s": 1.0, "minutes": 60.0, "mins": 60.0, "hours": 3600.0,
    "hrs": 3600.0, "days": 86400.0, "months": 2.6297368e06, "mts": 2.6297368e06,
    "years": 3.15576e07, "yrs": 3.15576e07,
}


def tconvert(t, from_unit=None, to_unit=None):
    """reference src/utils.jl:21-75 — `tconvert(from, to)` returns the factor."""
    if to_unit is None:
        t, from_unit, to_unit = 1.0, t, from_unit
    if from_unit not in T_UNIT_MAP or to_unit not in T_UNIT_MAP:
        raise RuntimeError("Unknown unit specified in time conversion!")
    return np.asarray(t, dtype=np.float64) * T_UNIT_MAP[from_unit] / T_UNIT_MAP[to_unit] if np.ndim(t) \
        else float(t) * T_UNIT_MAP[from_unit] / T_UNIT_MAP[to_unit]


def _float_range(start, step, stop):
    """Julia `collect(start:step:stop)` for Float64: rational lifting makes every element the
    correctly rounded start + i*step and fixes the length [upstream Base range semantics]."""
    a, s, b = (Fraction(repr(float(v))) for v in (start, step, stop))
    if float(a) != start or float(s) != step or float(b) != stop:
        a, s, b = Fraction(start), Fraction(step), Fraction(stop)
    n = math.floor((b - a) / s)
    return np.array([float(a + i * s) for i in range(n + 1)], dtype=np.float64) if n >= 0 else np.zeros(0)


def create_savepoints(start, stop, step):
    """reference src/utils.jl:108-115"""
    start, stop, step = float(start), float(stop), float(step)
    cstep = float(f"{step:.9g}") if (step > 1e-9 and abs(step - math.floor(step)) < 1e-9) else step
    r = _float_range(start, cstep, stop)
    if r[-1] < stop:
        r = np.append(r, stop)
    return r


class _Sol:
    """Linear-interpolating stand-in for `profile.sol` (DiffEqArray / ODESolution)."""

    def __init__(self, t, u):
        self.t = np.asarray(t, dtype=np.float64)
        self.u = np.asarray(u, dtype=np.float64)

    def __call__(self, tq):
        return float(np.interp(tq, self.t, self.u))


class AbstractConditionProfile:
    pass


class StaticConditionProfile(AbstractConditionProfile):
    """reference src/conditions/static.jl:7-9"""

    def __init__(self, value):
        self.value = float(value)

    def device_desc(self):
        p = np.zeros(16); p[0] = self.value
        return KIND_STATIC, p


class AbstractVariableProfile(AbstractConditionProfile):
    sol = None

    def _check_ts(self, ts_update):
        if ts_update > self.t_end:
            raise ValueError("Error defining tstops, `ts_update` is too large.")   # ArgumentError

    def value_at(self, t):
        raise NotImplementedError

    def values_at(self, ts):
        return np.array([self.value_at(float(t)) for t in ts])

    def solve(self, pars, reset=False):
        """solve_variable_condition! (direct_variable.jl:34-43, gradient_variable.jl:35-64):
        tabulates the profile on the save grid (∪ tstops for gradient profiles)."""
        if self.sol is not None and not reset:
            return
        si = pars.tspan[1] / 1000 if pars.save_interval is None else pars.save_interval
        t = create_savepoints(pars.tspan[0], pars.tspan[1], si)
        if isinstance(self, AbstractGradientProfile):
            ts = np.asarray(self.tstops, dtype=np.float64)
            t = np.sort(np.concatenate([t, ts[(ts >= pars.tspan[0]) & (ts <= pars.tspan[1])]]))
        self.sol = _Sol(t, self.values_at(t))

    def minimum(self):
        if self.sol is None:
            raise RuntimeError("Condition profile is missing a solution.")
        return float(np.min(self.sol.u))

    def maximum(self):
        if self.sol is None:
            raise RuntimeError("Condition profile is missing a solution.")
        return float(np.max(self.sol.u))


class AbstractDirectProfile(AbstractVariableProfile):
    def value_at(self, t):
        return self.f(t, self)


class AbstractGradientProfile(AbstractVariableProfile):
    pass


class NullDirectProfile(AbstractDirectProfile):
    """reference src/conditions/direct_variable.jl:49-92"""

    def __init__(self, *, X_start, t_end):
        self.X_start, self.t_end = float(X_start), float(t_end)
        self.tstops = np.array([self.t_end])
        self.f = lambda t, p: p.X_start

    def create_discrete_tstops(self, ts_update):
        self._check_ts(ts_update)
        self.tstops = _float_range(0.0, ts_update, self.t_end)

    def device_desc(self):
        p = np.zeros(16); p[0] = self.X_start
        return KIND_NULL, p


class LinearDirectProfile(AbstractDirectProfile):
    """reference src/conditions/direct_variable.jl:98-155"""

    def __init__(self, *, rate, X_start, X_end):
        rate, X_start, X_end = float(rate), float(X_start), float(X_end)
        if (X_end < X_start and rate > 0) or (X_end > X_start and rate < 0):
            raise RuntimeError("Impossible temperature ramp defined. Check heating rates have the correct signs.")
        self.rate, self.X_start, self.X_end = rate, X_start, X_end
        self.t_end = (X_end - X_start) / rate
        self.tstops = np.array([self.t_end])
        self.f = LinearDirectProfile._f

    @staticmethod
    def _f(t, p):
        if t <= 0.0:
            return p.X_start
        if t <= p.t_end:
            return p.X_start + p.rate * t
        return p.X_end

    def values_at(self, ts):
        ts = np.asarray(ts, dtype=np.float64)
        return np.where(ts <= 0.0, self.X_start, np.where(ts <= self.t_end, self.X_start + self.rate * ts, self.X_end))

    def create_discrete_tstops(self, ts_update):
        self._check_ts(ts_update)
        self.tstops = create_savepoints(0.0, self.t_end, ts_update)

    def device_desc(self):
        p = np.zeros(16); p[:4] = [self.rate, self.X_start, self.X_end, self.t_end]
        return KIND_LINEAR_DIRECT, p


class NullGradientProfile(AbstractGradientProfile):
    """reference src/conditions/gradient_variable.jl:70-114"""

    def __init__(self, *, X_start, t_end):
        self.X_start, self.t_end = float(X_start), float(t_end)
        self.tstops = np.array([self.t_end])
        self.grad = lambda t, p: 0.0

    def value_at(self, t):
        return self.X_start

    def create_discrete_tstops(self, ts_update):
        self._check_ts(ts_update)
        self.tstops = _float_range(0.0, ts_update, self.t_end)

    def device_desc(self):
        p = np.zeros(16); p[0] = self.X_start
        return KIND_NULL, p


class LinearGradientProfile(AbstractGradientProfile):
    """reference src/conditions/gradient_variable.jl:120-175"""

    def __init__(self, *, rate, X_start, X_end):
        rate, X_start, X_end = float(rate), float(X_start), float(X_end)
        if (X_end < X_start and rate > 0) or (X_end > X_start and rate < 0):
            raise RuntimeError("Impossible condition ramp defined. Check heating rates have the correct signs.")
        self.rate, self.X_start, self.X_end = rate, X_start, X_end
        self.t_end = (X_end - X_start) / rate
        self.tstops = np.array([self.t_end])
        # `rate` for every t <= t_end (negative t included), reference :165-170
        self.grad = lambda t, p: p.rate if t <= p.t_end else 0.0

    def value_at(self, t):
        return self.X_start + self.rate * max(min(t, self.t_end), 0.0)

    def create_discrete_tstops(self, ts_update):
        self._check_ts(ts_update)
        self.tstops = create_savepoints(0.0, self.t_end, ts_update)

    def device_desc(self):
        p = np.zeros(16); p[:4] = [self.rate, self.X_start, self.X_end, self.t_end]
        return KIND_LINEAR_GRADIENT, p


class DoubleRampGradientProfile(AbstractGradientProfile):
    """reference src/conditions/gradient_variable.jl:181-310"""

    def __init__(self, *, X_start, t_start_plateau, rate1, X_mid, t_mid_plateau, rate2, X_end,
                 t_end_plateau, t_blend=None):
        X_start, X_mid, X_end, rate1, rate2 = map(float, (X_start, X_mid, X_end, rate1, rate2))
        if ((X_mid > X_start and rate1 < 0) or (X_mid < X_start and rate1 > 0)
                or (X_end > X_mid and rate2 < 0) or (X_end < X_mid and rate2 > 0)):
            raise RuntimeError("Impossible condition ramp defined. Check heating rates have the correct signs.")
        self.rate1, self.rate2 = rate1, rate2
        self.X_start, self.X_mid, self.X_end = X_start, X_mid, X_end
        self.t_start_plateau, self.t_mid_plateau = float(t_start_plateau), float(t_mid_plateau)
        self.t_end_plateau = float(t_end_plateau)
        self.t_startr1 = self.t_start_plateau
        self.t_endr1 = self.t_startr1 + ((X_mid - X_start) / rate1)
        self.t_startr2 = self.t_endr1 + self.t_mid_plateau
        self.t_endr2 = self.t_startr2 + ((X_end - X_mid) / rate2)
        self.t_end = self.t_endr2 + self.t_end_plateau
        if t_blend is None:
            self.t_blend = 0.0
            self.tstops = np.array([self.t_startr1, self.t_endr1, self.t_startr2, self.t_endr2, self.t_end])
            self.grad = DoubleRampGradientProfile._grad
        else:
            tb = self.t_blend = float(t_blend)
            self.tstops = np.array([self.t_startr1 - tb, self.t_startr1 + tb, self.t_endr1 - tb, self.t_endr1 + tb,
                                    self.t_startr2 - tb, self.t_startr2 + tb, self.t_endr2 - tb, self.t_endr2 + tb,
                                    self.t_end])
            self.grad = DoubleRampGradientProfile._grad_blended

    @staticmethod
    def _grad(t, p):
        if p.t_startr1 <= t < p.t_endr1:
            return p.rate1
        if p.t_startr2 <= t < p.t_endr2:
            return p.rate2
        return 0.0

    @staticmethod
    def _grad_blended(t, p):
        tb = p.t_blend
        for ts, te, r in ((p.t_startr1, p.t_endr1, p.rate1), (p.t_startr2, p.t_endr2, p.rate2)):
            if ts - tb <= t < ts + tb:
                return r * (t - ts - tb) / (2 * tb) + r
            if ts + tb <= t < te - tb:
                return r
            if te - tb <= t < te + tb:
                return -r * (t - te - tb) / (2 * tb)
        return 0.0

    def value_at(self, t):
        """closed-form integral of grad from 0"""
        X, tb = self.X_start, self.t_blend
        if t <= 0:
            return X
        for ts, te, r in ((self.t_startr1, self.t_endr1, self.rate1), (self.t_startr2, self.t_endr2, self.rate2)):
            if tb > 0:
                a, b = ts - tb, ts + tb
                lo, hi = max(a, 0.0), min(b, t)
                if hi > lo:
                    X += 0.5 * (r / (2 * tb)) * ((hi - a) ** 2 - (lo - a) ** 2)
                lo, hi = max(b, 0.0), min(te - tb, t)
                if hi > lo:
                    X += r * (hi - lo)
                a, b = te - tb, te + tb
                lo, hi = max(a, 0.0), min(b, t)
                if hi > lo:
                    X += r * (hi - lo) - 0.5 * (r / (2 * tb)) * ((hi - a) ** 2 - (lo - a) ** 2)
            else:
                lo, hi = max(ts, 0.0), min(te, t)
                if hi > lo:
                    X += r * (hi - lo)
        return X

    def create_discrete_tstops(self, ts_update):
        self._check_ts(ts_update)
        tb = self.t_blend
        self.tstops = np.concatenate([[0.0],
                                      create_savepoints(self.t_startr1 - tb, self.t_endr1 + tb, ts_update),
                                      create_savepoints(self.t_startr2 - tb, self.t_endr2 + tb, ts_update),
                                      [self.t_end]])

    def device_desc(self):
        p = np.zeros(16)
        p[:8] = [self.X_start, self.rate1, self.rate2, self.t_startr1, self.t_endr1, self.t_startr2,
                 self.t_endr2, self.t_blend]
        return KIND_DOUBLE_RAMP, p


def isstatic(p):
    return isinstance(p, StaticConditionProfile)


def isvariable(p):
    return isinstance(p, AbstractVariableProfile)


class ConditionSet:
    """reference src/conditions/condition_set.jl:1-58 and accessors :61-191.
    (No Symbolics registration: the device evaluates profiles directly.)"""

    def __init__(self, d, ts_update=None):
        self.symbols, self.profiles = [], []
        for sym, v in d.items():
            if isinstance(v, (int, float)) and not isinstance(v, bool):
                self.profiles.append(StaticConditionProfile(v))
            elif isinstance(v, AbstractConditionProfile):
                if ts_update is not None and isvariable(v):
                    v.create_discrete_tstops(float(ts_update))
                self.profiles.append(v)
            else:
                raise ValueError(f"Condition {sym} does not have a valid profile.")   # ArgumentError
            self.symbols.append(sym)
        self.discrete_updates = ts_update is not None
        self.ts_update = None if ts_update is None else float(ts_update)

    def isstatic(self, sym=None):
        if sym is not None:
            return isstatic(self.get_profile(sym))
        return all(isstatic(p) for p in self.profiles)

    def isvariable(self, sym=None):
        if sym is not None:
            return isvariable(self.get_profile(sym))
        return all(isvariable(p) for p in self.profiles)

    def get_profile(self, sym):
        if sym not in self.symbols:
            raise RuntimeError(f"Condition {sym} does not exist in this ConditionSet")
        return self.profiles[self.symbols.index(sym)]

    def get_initial_conditions(self):
        return {s: (p.value if isstatic(p) else p.X_start) for s, p in zip(self.symbols, self.profiles)}

    def get_static_conditions(self):
        return {s: p.value for s, p in zip(self.symbols, self.profiles) if isstatic(p)}

    def get_variable_conditions(self):
        return {s: p.sol for s, p in zip(self.symbols, self.profiles) if isvariable(p)}

    def get_tstops(self):
        if self.isstatic():
            raise RuntimeError("No tstops available, all conditions in ConditionSet are static.")
        return np.unique(np.concatenate([np.asarray(p.tstops, dtype=np.float64)
                                         for p in self.profiles if isvariable(p)]))

    def get_t_final(self):
        if self.isstatic():
            raise RuntimeError("No t_end available, all conditions in ConditionSet are static.")
        return max(p.t_end for p in self.profiles if isvariable(p))

    def solve_variable_conditions(self, pars, reset=False):
        """solve_variable_conditions! (condition_set.jl:260-268)"""
        for p in self.profiles:
            if isvariable(p):
                p.solve(pars, reset=reset)


get_tstops = ConditionSet.get_tstops
get_t_final = ConditionSet.get_t_final
