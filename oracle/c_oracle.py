"""ctypes front-end of the plain-C oracle (oracle/crn_oracle.c).

TEST INFRASTRUCTURE ONLY — see the header of oracle/kinetica_oracle.py for who
may import this.  Builds liboracle on demand with `make -C oracle`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import kinetica_oracle as ko

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libcrn_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.ko_solve_rodas4.restype = C.c_int64
    return _LIB


def _p(a, ty):
    return a.ctypes.data_as(C.POINTER(ty))


def _csr(net: ko.Network):
    def flat(ids, nus):
        ptr = np.zeros(net.R + 1, dtype=np.int64)
        for j, r in enumerate(ids):
            ptr[j + 1] = ptr[j] + len(r)
        idx = np.array([s for r in ids for s in r], dtype=np.int64)
        nu = np.array([s for r in nus for s in r], dtype=np.int64)
        if idx.size == 0:
            idx = np.zeros(1, dtype=np.int64); nu = np.zeros(1, dtype=np.int64)
        return ptr, idx, nu
    return flat(net.id_reacs, net.stoic_reacs) + flat(net.id_prods, net.stoic_prods)


def arrhenius(A, Ea, T, k_max=None, t_mult=1.0):
    A = np.ascontiguousarray(A, dtype=np.float64)
    Ea = np.ascontiguousarray(Ea, dtype=np.float64)
    out = np.zeros_like(A)
    lib().ko_arrhenius(C.c_int64(len(A)), _p(A, C.c_double), _p(Ea, C.c_double), C.c_double(T),
                       C.c_double(np.nan if k_max is None else k_max), C.c_double(t_mult),
                       _p(out, C.c_double))
    return out


def rhs(net: ko.Network, u, k):
    rp, ri, rn, pp, pi, pn = _csr(net)
    u = np.ascontiguousarray(u, dtype=np.float64)
    k = np.ascontiguousarray(k, dtype=np.float64)
    du = np.zeros(net.S)
    lib().ko_rhs(C.c_int64(net.S), C.c_int64(net.R), _p(rp, C.c_int64), _p(ri, C.c_int64),
                 _p(rn, C.c_int64), _p(pp, C.c_int64), _p(pi, C.c_int64), _p(pn, C.c_int64),
                 _p(u, C.c_double), _p(k, C.c_double), _p(du, C.c_double))
    return du


def jac_csc(net: ko.Network, u, k):
    rp, ri, rn, pp, pi, pn = _csr(net)
    colptr, rowval = net.pattern_csc()
    u = np.ascontiguousarray(u, dtype=np.float64)
    k = np.ascontiguousarray(k, dtype=np.float64)
    out = np.zeros(max(len(rowval), 1))
    rv = rowval if len(rowval) else np.zeros(1, dtype=np.int64)
    lib().ko_jac_csc(C.c_int64(net.S), C.c_int64(net.R), _p(rp, C.c_int64), _p(ri, C.c_int64),
                     _p(rn, C.c_int64), _p(pp, C.c_int64), _p(pi, C.c_int64), _p(pn, C.c_int64),
                     _p(colptr, C.c_int64), _p(rv, C.c_int64), _p(u, C.c_double), _p(k, C.c_double),
                     _p(out, C.c_double))
    return out[:len(rowval)]


def merge_stops(tstops, saveat, t0, tf):
    """Merged sorted stop list inside [t0, tf] with flags (bit0 rate update, bit1 save);
    tf is always the last stop."""
    ts = np.asarray(tstops, dtype=np.float64) if tstops is not None else np.zeros(0)
    sv = np.asarray(saveat, dtype=np.float64)
    ts = ts[(ts >= t0) & (ts <= tf)]
    sv = sv[(sv >= t0) & (sv <= tf)]
    allt = np.unique(np.concatenate([ts, sv, [tf]]))
    flags = np.zeros(len(allt), dtype=np.int32)
    flags[np.isin(allt, ts)] |= 1
    flags[np.isin(allt, sv)] |= 2
    return allt, flags


def solve_rodas4(net: ko.Network, A, Ea, k_max, t_mult, T_init, tstops, T_stop_fn, u0, tspan, saveat,
                 abstol=1e-10, reltol=1e-8, maxiters=100000, ban_negatives=False, nthreads=0,
                 symbolic=None):
    """Solve B members.  `T_init[b]`; `T_stop_fn(b, t)` gives the member's condition at a stop.
    Returns (out_u[B, Ns, S], status[B], stats[B, 4], save_times)."""
    T_init = np.atleast_1d(np.asarray(T_init, dtype=np.float64))
    B = len(T_init)
    t0, tf = float(tspan[0]), float(tspan[1])
    stop_t, flags = merge_stops(tstops, saveat, t0, tf)
    ns = len(stop_t)
    T_stop = np.zeros((B, ns))
    for b in range(B):
        for s in range(ns):
            T_stop[b, s] = T_stop_fn(b, stop_t[s]) if (flags[s] & 1) else np.nan
    if symbolic is None:
        colptr, rowval = net.pattern_csc()
        perm = ko.min_degree_order(net.S, colptr, rowval)
        rowptr, colidx, diagpos, _ = ko.symbolic_lu(net.S, colptr, rowval, perm)
    else:
        perm, rowptr, colidx, diagpos = symbolic
    rp, ri, rn, pp, pi, pn = _csr(net)
    A = np.ascontiguousarray(A, dtype=np.float64)
    Ea = np.ascontiguousarray(Ea, dtype=np.float64)
    u0 = np.ascontiguousarray(u0, dtype=np.float64)
    stride = 0 if u0.ndim == 1 else net.S
    Ns = int(np.sum((flags & 2) != 0))
    out = np.zeros((B, Ns, net.S))
    status = np.zeros(B, dtype=np.int32)
    stats = np.zeros((B, 4), dtype=np.int64)
    perm = np.ascontiguousarray(perm, dtype=np.int64)
    lib().ko_solve_rodas4(
        C.c_int64(net.S), C.c_int64(net.R), _p(rp, C.c_int64), _p(ri, C.c_int64), _p(rn, C.c_int64),
        _p(pp, C.c_int64), _p(pi, C.c_int64), _p(pn, C.c_int64), _p(perm, C.c_int64),
        _p(rowptr, C.c_int64), _p(colidx, C.c_int64), _p(diagpos, C.c_int64),
        _p(A, C.c_double), _p(Ea, C.c_double), C.c_double(np.nan if k_max is None else k_max),
        C.c_double(t_mult), C.c_int64(B), _p(T_init, C.c_double), C.c_int64(ns),
        _p(stop_t, C.c_double), _p(flags, C.c_int32), _p(T_stop, C.c_double), _p(u0, C.c_double),
        C.c_int64(stride), C.c_double(t0), C.c_double(abstol), C.c_double(reltol),
        C.c_double(np.spacing(tf)), C.c_int64(maxiters), C.c_int32(int(ban_negatives)),
        C.c_int64(Ns), _p(out, C.c_double), _p(status, C.c_int32), _p(stats, C.c_int64),
        C.c_int32(nthreads))
    return out, status, stats, stop_t[(flags & 2) != 0]
