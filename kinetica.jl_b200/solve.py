"""`solve_network` front door of the B200 path — host-side mirror of the reference's
src/solving/methods.jl (method structs + dispatch), src/solving/solve_utils.jl (host
pre-processing), src/solving/filters.jl and the output container of src/analysis/io.jl.

The unchanged host steps of the reference run here in the same order (deepcopy, profile
pre-solution, reaction filter, calculator setup, low-k pruning, u0 assembly); everything from
"build the ReactionSystem" down (methods.jl:140-180 / 660-711) is replaced by one call into
libkinetica_b200.so.  There is no CPU fallback.
"""
from __future__ import annotations

import copy
import itertools
from dataclasses import dataclass
from typing import Any, List, Optional, Sequence

import numpy as np

from . import _lib
from .calculator import AbstractKineticCalculator
from .conditions import (ConditionSet, StaticConditionProfile, create_savepoints, isstatic, isvariable)
from .network import RxData, SpeciesData
from .params import ODESimulationParams

STOP_RATE, STOP_SAVE, STOP_CHUNK = 1, 2, 4
RETCODES = {0: "Success", 1: "MaxIters", 2: "DtLessThanMin", 3: "Unstable", 5: "Unfinished"}


# ---------------------------------------------------------------- filters (filters.jl:1-52)
class RxFilter:
    def __init__(self, filters=None, keep_filtered=False):
        self.filters = [lambda sd, rd: [False] * rd.nr] if filters is None else list(filters)
        self.keep_filtered = keep_filtered


def get_filter_mask(rf: RxFilter, sd, rd):
    if len(rf.filters) == 0:
        raise RuntimeError("RxFilter has not filter functions defined.")
    inv = ~np.asarray(rf.filters[0](sd, rd), dtype=bool)
    for f in rf.filters[1:]:
        inv &= ~np.asarray(f(sd, rd), dtype=bool)
    mask = ~inv
    return ~mask if rf.keep_filtered else mask


# ---------------------------------------------------------------- method structs (methods.jl:7-79)
class AbstractODESolveMethod:
    pass


class StaticODESolve(AbstractODESolveMethod):
    def __init__(self, pars, conditions, calculator, filter=None):
        if not conditions.isstatic():
            raise ValueError("All conditions must be static to run a StaticODESolve.")
        if not calculator.has_conditions(conditions.symbols):
            raise ValueError("Calculator does not support all of the provided conditions.")
        self.pars, self.conditions, self.calculator = pars, conditions, calculator
        self.filter = filter if filter is not None else RxFilter()


class VariableODESolve(AbstractODESolveMethod):
    def __init__(self, pars, conditions, calculator, filter=None):
        if not calculator.has_conditions(conditions.symbols):
            raise ValueError("Calculator does not support all of the provided conditions.")
        if not conditions.discrete_updates and not calculator.allows_continuous():
            raise ValueError("Calculator does not support continuous rate updates in simulations.")
        self.pars, self.conditions, self.calculator = pars, conditions, calculator
        self.filter = filter if filter is not None else RxFilter()


class B200EnsembleODESolve(AbstractODESolveMethod):
    """New capability (the reference has no ensembles, docs/src/tutorials/ode-solution.md:190):
    one ConditionSet per member, shared network / calculator / parameters."""

    def __init__(self, pars, conditions: Sequence[ConditionSet], calculator, filter=None):
        if len(conditions) == 0:
            raise ValueError("An ensemble needs at least one ConditionSet.")
        for cs in conditions:
            if not calculator.has_conditions(cs.symbols):
                raise ValueError("Calculator does not support all of the provided conditions.")
            if not cs.isstatic() and not cs.discrete_updates and not calculator.allows_continuous():
                raise ValueError("Calculator does not support continuous rate updates in simulations.")
        self.pars, self.conditions, self.calculator = pars, list(conditions), calculator
        self.filter = filter if filter is not None else RxFilter()


# ---------------------------------------------------------------- host pre-processing (solve_utils.jl)
def get_max_rates(conditions: ConditionSet, calculator):
    """solve_utils.jl:19-54"""
    static = conditions.get_static_conditions()
    variable = [(s, p) for s, p in zip(conditions.symbols, conditions.profiles) if isvariable(p)]
    if not variable:
        return calculator(**static)
    perms = []
    for bits in itertools.product((0, 1), repeat=len(variable)):
        kw = dict(static)
        for (s, p), bit in zip(variable, bits):
            kw[s] = p.maximum() if bit else p.minimum()
        perms.append(calculator(**kw))
    return perms[int(np.argmax([float(np.mean(p)) for p in perms]))]


def get_initial_rates(conditions: ConditionSet, calculator):
    """solve_utils.jl:62-73"""
    return calculator(**conditions.get_initial_conditions())


def calculate_discrete_rates(conditions: ConditionSet, calculator, nr):
    """solve_utils.jl:91-109 -> (tstops, k[Nt, nr])"""
    if not conditions.discrete_updates:
        raise RuntimeError("Cannot calculate discrete rates for a continuous ConditionSet.")
    tstops = conditions.get_tstops()
    static = conditions.get_static_conditions()
    vcs = conditions.get_variable_conditions()
    rows = []
    for ts in tstops:
        kw = dict(static)
        kw.update({s: sol(ts) for s, sol in vcs.items()})
        rows.append(np.asarray(calculator(**kw), dtype=np.float64))
    return tstops, np.array(rows).reshape(len(tstops), nr)


def low_k_removal(max_rates, pars: ODESimulationParams):
    """apply_low_k_cutoff! (solve_utils.jl:213-245) -> indices to remove (0-based)."""
    if pars.low_k_cutoff == "none":
        return np.zeros(0, dtype=np.int64)
    cutoff = pars.reltol / pars.tspan[1] if pars.low_k_cutoff == "auto" else float(pars.low_k_cutoff)
    return np.nonzero(np.asarray(max_rates) * pars.low_k_maxconc ** 2 < cutoff)[0].astype(np.int64)


def make_u0(sd: SpeciesData, pars: ODESimulationParams):
    """solve_utils.jl:262-297"""
    if isinstance(pars.u0, dict):
        u0 = np.zeros(sd.n)
        for spec, conc in pars.u0.items():
            if spec not in sd.toInt:
                raise RuntimeError(f"Species {spec} not in SpeciesData. Check pars.u0 is correct.")
            u0[sd.toInt[spec]] = conc
        return u0
    u0 = np.asarray(pars.u0, dtype=np.float64)
    if len(u0) != sd.n:
        if not pars.allow_short_u0:
            raise RuntimeError("Length of supplied initial concentration vector does not match with number of species in system.")
        full = np.zeros(sd.n)
        full[:len(u0)] = u0
        return full
    return u0.copy()


# ---------------------------------------------------------------- outputs (analysis/io.jl:3-48)
@dataclass
class Solution:
    t: np.ndarray
    u: List[np.ndarray]              # time-major, length-S inner vectors (res.sol.u)
    retcode: str = "Success"
    stats: Any = None

    def __call__(self, tq):
        U = np.array(self.u)
        return np.array([np.interp(tq, self.t, U[:, i]) for i in range(U.shape[1])])


class RateSolution:
    """`res.sol_k` of a discrete-update solve: the precalculated rate constants at the tstops
    (`DiffEqArray(k_precalc, tstops)` in the reference, solve_utils.jl:91-109, analysis/io.jl:36-38):
    `.t` = tstops, `.u[n]` = k(T(tstop_n)) for all reactions."""

    def __init__(self, t, u):
        self.t = np.asarray(t, dtype=np.float64)
        self.u = np.asarray(u, dtype=np.float64)

    def __call__(self, tq):
        """zero-order hold, like the rate update callback (solve_utils.jl:445-450)"""
        i = np.clip(np.searchsorted(self.t, tq, side="right") - 1, 0, len(self.t) - 1)
        return self.u[i]


class _Lazy:
    """a value computed on first access"""

    def __init__(self, fn):
        self.fn = fn


class ODESolveOutput:
    """analysis/io.jl:3-11: sd, rd, sol, sol_k, sol_vcs, pars, conditions.  `sol_k` is the table of
    precalculated rate constants for discrete-update solves (None for static conditions), `sol_vcs`
    the variable-condition trajectories when they are integrated along with the species (continuous
    mode) and None otherwise — the reference's rule (io.jl:36-38).  For ensembles `sol_k` is
    evaluated on first access (Nt x R doubles per member)."""

    def __init__(self, sd, rd, sol, sol_k=None, sol_vcs=None, pars=None, conditions=None, umax=None):
        self.sd, self.rd, self.sol = sd, rd, sol
        self._sol_k = sol_k
        self.sol_vcs, self.pars, self.conditions = sol_vcs, pars, conditions
        self.umax = umax               # per-species max over saved points (identify_next_seeds input)

    @property
    def sol_k(self):
        if isinstance(self._sol_k, _Lazy):
            self._sol_k = self._sol_k.fn()
        return self._sol_k


# ---------------------------------------------------------------- the device-backed solver
def merge_stops(tstops, saveat, t0, tf, chunks=None, plain=None):
    """Merged, sorted stop list of one member with flags.  Times from different sources that differ
    only in the last bits (a chunk boundary `nc * chunkstep` against a tstop of the profile's range
    arithmetic) are one stop; the tstop's value is kept (it is where the profile is evaluated)."""
    ts = np.asarray(tstops, dtype=np.float64) if tstops is not None else np.zeros(0)
    sv = np.asarray(saveat, dtype=np.float64)
    ck = np.asarray(chunks, dtype=np.float64) if chunks is not None else np.zeros(0)
    pl = np.asarray(plain, dtype=np.float64) if plain is not None else np.zeros(0)    # forced step ends without an action
    ts = ts[(ts >= t0) & (ts <= tf)]
    sv = sv[(sv >= t0) & (sv <= tf)]
    ck = ck[(ck > t0) & (ck < tf)]
    pl = pl[(pl > t0) & (pl < tf)]
    t_all = np.concatenate([ts, sv, ck, pl, [tf]])
    f_all = np.concatenate([np.full(len(ts), STOP_RATE), np.full(len(sv), STOP_SAVE), np.full(len(ck), STOP_CHUNK),
                            np.zeros(len(pl)), [0]]).astype(np.int32)
    pri = np.concatenate([np.zeros(len(ts)), np.ones(len(sv)), np.ones(len(ck)), np.ones(len(pl)), [1]])   # tstops first inside a cluster
    order = np.lexsort((pri, t_all))
    t_all, f_all = t_all[order], f_all[order]
    tol = 1e-12 * max(1.0, abs(tf))
    new = np.concatenate([[True], np.diff(t_all) > tol])
    grp = np.cumsum(new) - 1
    allt = np.zeros(grp[-1] + 1)
    flags = np.zeros(grp[-1] + 1, dtype=np.int32)
    # value of a cluster: its tstop if it has one (sorted first among equal times is not guaranteed, so pick explicitly)
    allt[grp[::-1]] = t_all[::-1]                       # first element of every cluster ...
    is_ts = f_all == STOP_RATE
    allt[grp[is_ts]] = t_all[is_ts]                     # ... overridden by the cluster's tstop
    np.bitwise_or.at(flags, grp, f_all)
    allt[-1] = tf if abs(allt[-1] - tf) <= tol else allt[-1]
    return allt, flags


def chunk_grid(pars: ODESimulationParams):
    """Chunk boundaries and save times of a chunkwise solve (methods.jl:214-222, 757-765): chunks of
    `solve_chunkstep`, local save points 0:save_interval:chunkstep (save_interval defaults to the
    chunk step), global time = local time + nc * chunkstep, `(len(saveat_local) - 1) * n_chunks + 1`
    points in all."""
    t0, tf = pars.tspan
    step = pars.solve_chunkstep
    n_chunks = int(tf / step)
    si = step if pars.save_interval is None else pars.save_interval
    nloc = int(np.floor(step / si + 1e-9)) + 1                       # length of 0.0:si:step
    local = np.arange(nloc) * si
    save = [local[i] + nc * step for nc in range(n_chunks) for i in range(nloc - 1)]
    save.append(local[-1] + (n_chunks - 1) * step)
    bounds = np.arange(1, n_chunks) * step
    return np.array(bounds), np.array(save)


class EnsembleSolver:
    """Network-level state on one GPU: symbolic factorisation + calculator tables, reusable
    across solves (what the reference rebuilds through MTK on every solve_network call)."""

    def __init__(self, sd: SpeciesData, rd: RxData, calculator, device=0, ordering=4):
        self.sd, self.rd, self.calculator = sd, rd, calculator
        self.h = _lib.Handle(device)
        self.h.set_network(sd.n, *rd.flatten())
        self.nnzJ, self.nnzLU, self.n_fma = self.h.symbolic(ordering)
        self.dev = calculator.device_arrhenius() if hasattr(calculator, "device_arrhenius") else None
        if self.dev is not None:
            self.h.set_arrhenius(self.dev["A"], self.dev["Ea"], self.dev["n"], self.dev["k_max"], self.dev["t_mult"])

    def close(self):
        self.h.close()

    def _conditions_to_device(self, conds: Sequence[ConditionSet], stop_t, flags):
        """stop_t/flags: [B, ns] per-member tables (or [ns] shared)."""
        B = len(conds)
        shared = stop_t.ndim == 1
        if self.dev is not None:
            kinds = np.zeros(B, dtype=np.int32)
            params = np.zeros((B, 16))
            need_table = False
            ns = stop_t.shape[-1]
            table = np.full((B, ns), np.nan)
            for b, cs in enumerate(conds):
                prof = cs.get_profile("T")
                kinds[b], params[b] = prof.device_desc()
                if isvariable(prof):
                    st = stop_t if shared else stop_t[b]
                    ridx = np.nonzero((flags if shared else flags[b]) & STOP_RATE)[0]
                    if len(ridx) == 0:
                        continue
                    # the reference reads the INTERPOLATED profile solution at each tstop
                    # (solve_utils.jl:101-104); ship a table only where that differs from X(t)
                    ref = np.interp(st[ridx], prof.sol.t, prof.sol.u)
                    exact = prof.values_at(st[ridx])
                    if np.any(np.abs(ref - exact) > 1e-12 * np.maximum(np.abs(exact), 1.0)):
                        need_table = True
                    table[b, ridx] = ref
            return dict(kinds=kinds, params=params, table=table if need_table else None, sol_k=None)
        # calculators without a device kernel: host table (single member)
        if B != 1:
            raise NotImplementedError("host-tabulated calculators are supported for single-member solves only")
        cs = conds[0]
        st = stop_t if shared else stop_t[0]
        fl = flags if shared else flags[0]
        rate_stops = st[(fl & STOP_RATE) != 0]
        k_init = np.asarray(get_initial_rates(cs, self.calculator), dtype=np.float64)
        if cs.isstatic() or len(rate_stops) == 0:
            k_table = np.zeros((0, self.rd.nr))
            sol_k = None
        else:
            ts, k_all = calculate_discrete_rates(cs, self.calculator, self.rd.nr)
            k_table = k_all[np.isin(ts, rate_stops)]
            sol_k = (ts, k_all)
        return dict(k_table=k_table, k_init=k_init, sol_k=sol_k)

    def bind(self, conds: Sequence[ConditionSet], pars: ODESimulationParams):
        """Host-side assembly of the stop / profile / rate tables of an ensemble and their hand-over to
        the library.  Explicit: the tables of the last `bind` stay in force until the next one (bench.py
        binds once and re-runs `prepare`; `solve_network` binds on every call)."""
        if self.calculator is not None and hasattr(self.calculator, "Ea") and len(self.calculator.Ea) != self.rd.nr:
            raise ValueError("calculator and reaction data disagree on the number of reactions")
        self._bind(conds, pars)
        b = self._bound
        if b["shared"]:
            self.h.set_stops(b["stop_t"], b["flags"])
        else:
            self.h.set_member_stops(b["counts"], b["stop_t"], b["flags"])
        if self.dev is not None:
            self.h.set_profiles(b["kinds"], b["params"])
            self.h.set_T_table(b["table"])
        else:
            self.h.set_rate_table(b["k_table"], b["k_init"])
            self.h.set_T_table(None)
        self.h.set_continuous(self.continuous)
        # chunkwise: a failed chunk is repeated on the device (adaptive_solve! per chunk)
        self.h.set_chunking(bool(pars.solve_chunks and pars.adaptive_tols), bool(pars.update_tols))
        self._bound_B, self._bound_pars = len(conds), pars
        return b["sol_k"]

    def _solve_args(self, pars, u0):
        t0, tf = pars.tspan
        # dtmin = eps(tspan[end]) (methods.jl:164), eps(solve_chunkstep) in chunkwise solves (:231)
        dtmin = float(np.spacing(pars.solve_chunkstep if pars.solve_chunks else tf))
        return (self._bound_B, u0, t0, pars.abstol, pars.reltol, dtmin, pars.maxiters, pars.ban_negatives, len(self.save_t))

    def prepare(self, conds: Sequence[ConditionSet], pars: ODESimulationParams, u0):
        """Upload of the bound ensemble (H2D of u0, profiles, stop tables).  `conds` must be the
        ensemble of the last `bind` (bound here on first use)."""
        if getattr(self, "_bound", None) is None or self._bound_B != len(conds):
            self.bind(conds, pars)
        self.h.solve_prepare(*self._solve_args(pars, u0))
        return self._bound["sol_k"]

    def solve(self, conds: Sequence[ConditionSet], pars: ODESimulationParams, u0):
        """bind + one-shot kb2_solve (batch-tiled when the ensemble does not fit the device memory)
        -> (out_u[Ns,S,B], umax[S,B], status[B], stats[B,8], sol_k)"""
        sol_k = self.bind(conds, pars)
        out_u, umax, status, stats = self.h.solve(*self._solve_args(pars, u0))
        return out_u, umax, status, stats, sol_k

    def _bind(self, conds: Sequence[ConditionSet], pars: ODESimulationParams):
        t0, tf = pars.tspan
        if pars.solve_chunks:
            # the reference default (params.jl:65): the integrator is re-initialised every solve_chunkstep
            chunks, saveat = chunk_grid(pars)
        else:
            # complete solve; save_interval = nothing means "every step" in the reference (saveat = []):
            # here a fixed grid of tf/1000 (the device saves at stops only) — documented deviation
            chunks = None
            si = pars.save_interval if pars.save_interval is not None else tf / 1000
            saveat = create_savepoints(t0, tf, si)
        B = len(conds)
        # continuous rate updates (methods.jl:363-458): the profile's tstops (its kinks) are only forced
        # step ends, k follows the profile inside the step on the device
        cont = (not conds[0].isstatic()) and (not conds[0].discrete_updates)
        if any(((not cs.isstatic()) and (not cs.discrete_updates)) != cont for cs in conds):
            raise ValueError("discrete and continuous rate updates cannot be mixed in one ensemble")
        if cont and self.dev is None:
            raise ValueError("Calculator does not support continuous rate updates on the device.")
        self.continuous = cont

        def stops_of(cs):
            ts = None if cs.isstatic() else cs.get_tstops()
            return merge_stops(None, saveat, t0, tf, chunks, plain=ts) if cont else merge_stops(ts, saveat, t0, tf, chunks)

        if conds[0].isstatic():
            tstops0, same = None, all(cs.isstatic() for cs in conds)
            if not same:
                raise ValueError("static and variable ConditionSets cannot be mixed in one ensemble")
        else:
            tstops0 = conds[0].get_tstops()
            same = True
            for cs in conds[1:]:
                ts = cs.get_tstops()
                if len(ts) != len(tstops0) or np.any(ts != tstops0):
                    same = False
                    break
        bound = {"shared": same, "counts": None}
        if same:
            stop_t, flags = stops_of(conds[0])
            self.save_t = stop_t[(flags & STOP_SAVE) != 0]
        else:
            # members with their own tstops grids (e.g. t_end differing in the last bit)
            lists = [stops_of(cs) for cs in conds]
            nmax = max(len(t) for t, _ in lists)
            stop_t = np.zeros((B, nmax))
            flags = np.zeros((B, nmax), dtype=np.int32)
            counts = np.zeros(B, dtype=np.int32)
            for b, (t, f) in enumerate(lists):
                counts[b] = len(t)
                stop_t[b, :len(t)] = t
                stop_t[b, len(t):] = np.inf
                flags[b, :len(t)] = f
            stop_t = np.where(np.isfinite(stop_t), stop_t, 0.0)
            bound["counts"] = counts
            self.save_t = lists[0][0][(lists[0][1] & STOP_SAVE) != 0]
        bound["stop_t"], bound["flags"] = stop_t, flags
        bound.update(self._conditions_to_device(conds, stop_t, flags))
        self._bound = bound

    def run(self):
        return self.h.solve_run()

    def fetch(self, out_u=None, out_umax=None):
        return self.h.solve_fetch(out_u, out_umax)


def _solve_with_retry(solver: EnsembleSolver, conds, pars, u0):
    """adaptive_solve! (solve_utils.jl:376-424) per member: members whose solve failed are solved
    again, as a smaller ensemble, with abstol/reltol tightened x0.1 — at most 5 attempts, never
    below eps; members that succeeded keep their result (each member is its own `solve!`)."""
    p = copy.copy(pars)
    mintol = np.finfo(np.float64).eps
    B = len(conds)
    todo = np.arange(B)
    out_u = umax = status = stats = None
    sol_k = None
    iters = 0
    while True:
        iters += 1
        sub = [conds[b] for b in todo]
        ou, um, st, sx, sk = solver.solve(sub, p, u0)
        if out_u is None:
            out_u, umax, status, stats, sol_k = ou, um, st.copy(), sx, sk
        else:
            out_u[:, :, todo], umax[:, todo], status[todo], stats[todo] = ou, um, st, sx
        bad = st != 0
        if not np.any(bad):
            if pars.update_tols and p.abstol != pars.abstol:
                pars.abstol, pars.reltol = p.abstol, p.reltol
            return out_u, umax, status, stats, sol_k
        # chunkwise solves have already repeated the failed chunk on the device (five attempts)
        if pars.solve_chunks or not pars.adaptive_tols or iters >= 5 or p.abstol / 10 <= mintol or p.reltol / 10 <= mintol:
            raise RuntimeError("ODE solution failed.")       # ErrorException, solve_utils.jl:405-411
        todo = todo[bad]
        p.abstol /= 10
        p.reltol /= 10


def solve_network(method: AbstractODESolveMethod, sd: SpeciesData, rd: RxData, copy_network=True,
                  device=0, solver: Optional[EnsembleSolver] = None, _keep_solver: Optional[dict] = None):
    """methods.jl:105-130 (static) / :330-360 (variable) + the new ensemble method.
    Returns an `ODESolveOutput` (a list of them for `B200EnsembleODESolve`)."""
    if copy_network:
        sd, rd = copy.deepcopy(sd), copy.deepcopy(rd)
    pars, calc = method.pars, method.calculator
    ensemble = isinstance(method, B200EnsembleODESolve)
    conds = method.conditions if ensemble else [method.conditions]
    if copy_network:
        calc = copy.deepcopy(calc)
    for cs in conds:
        cs.solve_variable_conditions(pars)
    mask = get_filter_mask(method.filter, sd, rd)
    rids = np.nonzero(mask)[0]
    rd.splice(rids)
    if len(rids):
        calc.splice(rids)
    calc.setup_network(sd, rd)
    # low-k pruning with ensemble-wide maximum rates (identical to the reference for one member)
    if pars.low_k_cutoff != "none":
        mr = None
        for cs in conds:
            r = np.asarray(get_max_rates(cs, calc))
            mr = r if mr is None else np.maximum(mr, r)
        low = low_k_removal(mr, pars)
        rd.splice(low)
        if len(low):
            calc.splice(low)
    u0 = make_u0(sd, pars)
    own = solver is None
    if own:
        solver = EnsembleSolver(sd, rd, calc, device=device)
    try:
        out_u, umax, status, stats, sol_k = _solve_with_retry(solver, conds, pars, u0)
        save_t = solver.save_t
    except BaseException:
        if own:
            solver.close()
        raise
    if own and _keep_solver is None:
        solver.close()
    if _keep_solver is not None:
        _keep_solver["solver"] = solver          # the caller closes it (multi-GPU gather of its results)
    continuous = getattr(solver, "continuous", False)

    def vc_solution(cs):
        """res.sol_vcs (analysis/io.jl:36-37): the variable conditions at the save times, as the
        continuous solve carries them along with the species"""
        from .conditions import _Sol
        return {s: _Sol(save_t.copy(), np.asarray(p.values_at(save_t), dtype=np.float64))
                for s, p in zip(cs.symbols, cs.profiles) if isvariable(p)}

    def rate_table(cs):
        """res.sol_k: k at the tstops (calculate_discrete_rates); None for static conditions"""
        if cs.isstatic() or not cs.discrete_updates:
            return None
        ts, k_all = calculate_discrete_rates(cs, calc, rd.nr)
        return RateSolution(ts, k_all)

    outs = []
    for b, cs in enumerate(conds):
        sol = Solution(t=save_t.copy(), u=[out_u[s, :, b].copy() for s in range(out_u.shape[0])],
                       retcode=RETCODES.get(int(status[b]), "Failure"), stats=stats[b].copy())
        if sol_k is not None:
            sk = RateSolution(*sol_k)
        else:
            sk = _Lazy(lambda cs=cs: rate_table(cs)) if ensemble else rate_table(cs)
        outs.append(ODESolveOutput(sd=sd, rd=rd, sol=sol, sol_k=None if continuous else sk,
                                   sol_vcs=vc_solution(cs) if continuous else None, pars=pars, conditions=cs,
                                   umax=umax[:, b].copy()))
    return outs if ensemble else outs[0]
