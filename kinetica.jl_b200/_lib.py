"""ctypes binding of libkinetica_b200.so (the C ABI in include/kinetica_b200.h).

The product path has no CPU fallback: if the shared library is missing or no CUDA
device is usable, loading / `Handle()` raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KB2_LIB") or os.path.join(_HERE, "libkinetica_b200.so")   # KB2_LIB: experimental builds only

_i32, _i64, _f64 = C.c_int32, C.c_int64, C.c_double
_pi32, _pi64, _pf64 = C.POINTER(_i32), C.POINTER(_i64), C.POINTER(_f64)
_H = C.c_void_p

# name -> (restype, argtypes): every symbol include/kinetica_b200.h declares
SIGNATURES = {
    "kb2_create": (_i32, [_i32, C.POINTER(_H)]),
    "kb2_destroy": (_i32, [_H]),
    "kb2_last_error": (C.c_char_p, [_H]),
    "kb2_launch_count": (_i64, [_H]),
    "kb2_set_network": (_i32, [_H, _i64, _i64, _pi64, _pi64, _pi64, _pi64, _pi64, _pi64]),
    "kb2_set_ordering": (_i32, [_H, _pi64]),
    "kb2_symbolic": (_i32, [_H, _i32, _pi64, _pi64, _pi64]),
    "kb2_get_pattern": (_i32, [_H, _pi64, _pi64]),
    "kb2_get_ordering": (_i32, [_H, _pi64]),
    "kb2_get_lu_pattern": (_i32, [_H, _pi64, _pi64, _pi64]),
    "kb2_get_plan_stats": (_i32, [_H, _pi64]),
    "kb2_get_plan_array": (_i64, [_H, _i32, _pi32, _i64]),
    "kb2_set_arrhenius": (_i32, [_H, _pf64, _pf64, _pf64, _f64, _f64]),
    "kb2_set_rate_table": (_i32, [_H, _i64, _pf64, _pf64]),
    "kb2_set_profiles": (_i32, [_H, _i64, _pi32, _pf64]),
    "kb2_set_T_table": (_i32, [_H, _i64, _i64, _pf64]),
    "kb2_set_stops": (_i32, [_H, _i64, _pf64, _pi32]),
    "kb2_set_member_stops": (_i32, [_H, _i64, _i64, _pi32, _pf64, _pi32]),
    "kb2_solve": (_i32, [_H, _i64, _pf64, _i64, _f64, _f64, _f64, _f64, _i64, _i32, _i64, _pf64, _pf64, _pi32, _pi64]),
    "kb2_memory_plan": (_i32, [_H, _i64, _pi64, _pi64, _pi64]),
    "kb2_set_batch_tile": (_i32, [_H, _i64]),
    "kb2_last_batch_tiles": (_i64, [_H]),
    "kb2_set_chunking": (_i32, [_H, _i32, _i32]),
    "kb2_set_continuous": (_i32, [_H, _i32]),
    "kb2_solve_prepare": (_i32, [_H, _i64, _pf64, _i64, _f64, _f64, _f64, _f64, _i64, _i32, _i64]),
    "kb2_solve_run": (_i32, [_H, C.POINTER(C.c_float)]),
    "kb2_solve_fetch": (_i32, [_H, _pf64, _pf64, _pi32, _pi64]),
    "kb2_get_phase_times": (_i32, [_H, _pf64, _pi64, _pi64]),
    "kb2_pack_results_device": (_i32, [_H, C.c_void_p, C.c_void_p]),
    "kb2_eval_k": (_i32, [_H, _i64, _pf64, _pf64]),
    "kb2_eval_profile": (_i32, [_H, _i64, _i64, _pf64, _pf64]),
    "kb2_eval_rhs": (_i32, [_H, _i64, _pf64, _pf64, _pf64]),
    "kb2_eval_jac": (_i32, [_H, _i64, _pf64, _pf64, _pf64]),
    "kb2_factor": (_i32, [_H, _i64, _pf64, _pf64, _pf64, _pf64]),
    "kb2_trisolve": (_i32, [_H, _i64, _pf64, _pf64]),
    "kb2_time_kernel": (_i32, [_H, _i32, _i64, _i32, C.POINTER(C.c_float)]),
    "kb2_comm_unique_id": (_i32, [C.POINTER(C.c_uint8)]),
    "kb2_comm_init_rank": (_i32, [_H, _i32, _i32, C.POINTER(C.c_uint8)]),
    "kb2_comm_init_all": (_i32, [_i32, C.POINTER(_H)]),
    "kb2_allgather_results": (_i32, [C.POINTER(_H), _i32, C.POINTER(_pf64), C.POINTER(_pf64)]),
    "kb2_gathered_device": (_i32, [_H, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_float), _pi32, _pi32]),
    "kb2_measure_fp64_peak": (_i32, [_H, _pf64]),
    "kb2_set_tiling": (_i32, [_H, _i32, _i32]),
    "kb2_get_launch_info": (_i32, [_H, _pi32, _pi32]),
}

_lib = None


class Kb2Error(RuntimeError):
    pass


def load():
    """Load the shared library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Kb2Error(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
                           " (make -C kinetica.jl_b200/csrc). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def _f(a):
    return None if a is None else a.ctypes.data_as(_pf64)


def _i(a):
    return None if a is None else a.ctypes.data_as(_pi64)


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Handle:
    """One handle = one GPU = one stream.  `device=-1` gives a host-only handle on which only
    the symbolic analysis works (used by CPU tests); every compute call on it fails."""

    def __init__(self, device: int = 0):
        self._lib = load()
        self._h = _H()
        rc = self._lib.kb2_create(device, C.byref(self._h))
        if rc != 0:
            raise Kb2Error(f"kb2_create(device={device}) failed with status {rc}: no usable CUDA device "
                           "(there is no CPU fallback)")
        self.S = self.R = 0
        self.nnzJ = self.nnzLU = self.n_fma = 0

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.kb2_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise Kb2Error(f"status {rc}: {self._lib.kb2_last_error(self._h).decode()}")

    @property
    def launch_count(self) -> int:
        return int(self._lib.kb2_launch_count(self._h))

    # ---- network + symbolic ----
    def set_network(self, S, reac_ptr, reac_idx, reac_nu, prod_ptr, prod_idx, prod_nu):
        arrs = [np.ascontiguousarray(a, dtype=np.int64) for a in (reac_ptr, reac_idx, reac_nu, prod_ptr, prod_idx, prod_nu)]
        arrs = [a if a.size else np.zeros(1, dtype=np.int64) for a in arrs]
        R = len(reac_ptr) - 1
        self._ck(self._lib.kb2_set_network(self._h, S, R, *[_i(a) for a in arrs]))
        self.S, self.R = int(S), int(R)

    def symbolic(self, ordering=0, perm=None):
        if perm is not None:
            p = np.ascontiguousarray(perm, dtype=np.int64)
            self._ck(self._lib.kb2_set_ordering(self._h, _i(p)))
            ordering = 2
        a, b, c = _i64(), _i64(), _i64()
        self._ck(self._lib.kb2_symbolic(self._h, ordering, C.byref(a), C.byref(b), C.byref(c)))
        self.nnzJ, self.nnzLU, self.n_fma = a.value, b.value, c.value
        return self.nnzJ, self.nnzLU, self.n_fma

    def get_pattern(self):
        colptr = np.zeros(self.S + 1, dtype=np.int64)
        rowval = np.zeros(max(self.nnzJ, 1), dtype=np.int64)
        self._ck(self._lib.kb2_get_pattern(self._h, _i(colptr), _i(rowval)))
        return colptr, rowval[:self.nnzJ]

    def get_ordering(self):
        perm = np.zeros(self.S, dtype=np.int64)
        self._ck(self._lib.kb2_get_ordering(self._h, _i(perm)))
        return perm

    def get_plan_stats(self):
        out = np.zeros(8, dtype=np.int64)
        self._ck(self._lib.kb2_get_plan_stats(self._h, _i(out)))
        return dict(zip(["padded", "panels", "units", "tasks", "fma_padded", "max_width", "map_entries", "ordering"],
                        map(int, out)))

    PLAN_ARRAYS = ["p_row0", "p_nrows", "p_width", "p_next", "p_base", "p_cptr", "cols", "u_info", "t_info", "map",
                   "slot_of", "jslot", "diag_slot"]

    def get_plan(self):
        """The raw block-plan tables (host-side verification of the factorisation schedule)."""
        out = {}
        for which, name in enumerate(self.PLAN_ARRAYS):
            n = int(self._lib.kb2_get_plan_array(self._h, which, None, 0))
            if n < 0:
                raise Kb2Error("plan table %s unavailable" % name)
            a = np.zeros(max(n, 1), dtype=np.int32)
            self._lib.kb2_get_plan_array(self._h, which, a.ctypes.data_as(_pi32), n)
            out[name] = a[:n]
        out["u_info"] = out["u_info"].reshape(-1, 12)
        out["t_info"] = out["t_info"].reshape(-1, 12)
        return out

    GATHER_ARRAYS = ["rhs_ptr", "rhs_rxn", "rhs_coef", "rhs_order", "rate_pos", "ell_ptr", "ell", "jt_ptr", "jt_rxn",
                     "jt_pack", "j_order", "drate_pos", "jell_ptr", "jell", "jt_pk", "meta"]

    def get_gather_tables(self):
        """The gather tables of the right-hand side and the Jacobian (host-side verification):
        CSR rows, work orders, first-touch layouts and the sliced ELLs the device walks."""
        out = {}
        for k, name in enumerate(self.GATHER_ARRAYS):
            which = 13 + k
            n = int(self._lib.kb2_get_plan_array(self._h, which, None, 0))
            if n < 0:
                raise Kb2Error("gather table %s unavailable" % name)
            a = np.zeros(max(n, 1), dtype=np.int32)
            self._lib.kb2_get_plan_array(self._h, which, a.ctypes.data_as(_pi32), n)
            out[name] = a[:n]
        m = out.pop("meta")
        out.update(rhs_nlong=int(m[0]), j_nlong=int(m[1]), jslots=int(m[2]), ell_g=int(m[3]))
        return out

    def get_front_plan(self):
        """The front plan of the window LU (host-side verification)."""
        out = {}
        for which, name in ((29, "f_info"), (30, "lists"), (31, "init"), (32, "meta")):
            n = int(self._lib.kb2_get_plan_array(self._h, which, None, 0))
            if n < 0:
                raise Kb2Error("front plan table %s unavailable" % name)
            a = np.zeros(max(n, 1), dtype=np.int32)
            self._lib.kb2_get_plan_array(self._h, which, a.ctypes.data_as(_pi32), n)
            out[name] = a[:n]
        m = out.pop("meta")
        out.update(NF=int(m[0]), Wr=int(m[1]), Wc=int(m[2]), max_nl=int(m[3]), max_nu=int(m[4]), max_init=int(m[5]))
        out["f_info"] = out["f_info"].reshape(-1, 12)
        out["init"] = out["init"].reshape(-1, 2)
        return out

    def get_launch_info(self):
        a, b = _i32(), _i32()
        self._ck(self._lib.kb2_get_launch_info(self._h, C.byref(a), C.byref(b)))
        return {"members_per_tile": a.value, "ctas_per_sm": b.value}

    def get_lu_pattern(self):
        rowptr = np.zeros(self.S + 1, dtype=np.int64)
        colidx = np.zeros(self.nnzLU, dtype=np.int64)
        diagpos = np.zeros(self.S, dtype=np.int64)
        self._ck(self._lib.kb2_get_lu_pattern(self._h, _i(rowptr), _i(colidx), _i(diagpos)))
        return rowptr, colidx, diagpos

    # ---- calculator / conditions ----
    def set_arrhenius(self, A, Ea, n=None, k_max=None, t_mult=1.0):
        A, Ea = _c64(A), _c64(Ea)
        nn = None if n is None else _c64(n)
        self._ck(self._lib.kb2_set_arrhenius(self._h, _f(A), _f(Ea), _f(nn),
                                             float("nan") if k_max is None else float(k_max), float(t_mult)))

    def set_rate_table(self, k_table, k_init):
        kt, ki = _c64(k_table), _c64(k_init)
        self._ck(self._lib.kb2_set_rate_table(self._h, kt.shape[0], _f(kt), _f(ki)))

    def set_profiles(self, kind, params):
        kind = np.ascontiguousarray(kind, dtype=np.int32)
        params = _c64(params)
        assert params.shape == (len(kind), 16)
        self._ck(self._lib.kb2_set_profiles(self._h, len(kind), kind.ctypes.data_as(_pi32), _f(params)))

    def set_T_table(self, T):
        if T is None:
            self._ck(self._lib.kb2_set_T_table(self._h, 0, 0, None))
            return
        T = _c64(T)
        self._ck(self._lib.kb2_set_T_table(self._h, T.shape[0], T.shape[1], _f(T)))

    def set_stops(self, stop_t, flags):
        st = _c64(stop_t)
        fl = np.ascontiguousarray(flags, dtype=np.int32)
        self._ck(self._lib.kb2_set_stops(self._h, len(st), _f(st), fl.ctypes.data_as(_pi32)))

    def set_member_stops(self, counts, stop_t, flags):
        cnt = np.ascontiguousarray(counts, dtype=np.int32)
        st = _c64(stop_t)
        fl = np.ascontiguousarray(flags, dtype=np.int32)
        assert st.shape == fl.shape == (len(cnt), st.shape[1])
        self._ck(self._lib.kb2_set_member_stops(self._h, st.shape[0], st.shape[1], cnt.ctypes.data_as(_pi32),
                                                _f(st), fl.ctypes.data_as(_pi32)))

    def set_chunking(self, retry_failed_chunks=False, update_tols=False):
        self._ck(self._lib.kb2_set_chunking(self._h, int(bool(retry_failed_chunks)), int(bool(update_tols))))

    def set_continuous(self, continuous=False):
        self._ck(self._lib.kb2_set_continuous(self._h, int(bool(continuous))))

    def set_tiling(self, members_per_tile=0, reserved=0):
        self._ck(self._lib.kb2_set_tiling(self._h, members_per_tile, reserved))

    # ---- solve ----
    def solve_prepare(self, B, u0, t0, abstol, reltol, dtmin, maxiters, ban_negatives, Ns):
        u0 = _c64(u0)
        stride = 0 if u0.ndim == 1 else self.S
        self._u0_keep = u0
        self._B, self._Ns = int(B), int(Ns)
        self._ck(self._lib.kb2_solve_prepare(self._h, B, _f(u0), stride, t0, abstol, reltol, dtmin,
                                             int(maxiters), int(bool(ban_negatives)), Ns))

    def solve(self, B, u0, t0, abstol, reltol, dtmin, maxiters, ban_negatives, Ns):
        """One-shot solve with batch tiling (kb2_solve): returns (out_u[Ns,S,B], umax[S,B], status, stats)."""
        u0 = _c64(u0)
        stride = 0 if u0.ndim == 1 else self.S
        self._B, self._Ns = int(B), int(Ns)
        out_u = np.empty((Ns, self.S, B))
        umax = np.empty((self.S, B))
        status = np.zeros(B, dtype=np.int32)
        stats = np.zeros((B, 8), dtype=np.int64)
        self._ck(self._lib.kb2_solve(self._h, B, _f(u0), stride, t0, abstol, reltol, dtmin, int(maxiters),
                                     int(bool(ban_negatives)), Ns, _f(out_u), _f(umax), status.ctypes.data_as(_pi32), _i(stats)))
        return out_u, umax, status, stats

    def memory_plan(self, Ns):
        a, b, c = _i64(), _i64(), _i64()
        self._ck(self._lib.kb2_memory_plan(self._h, Ns, C.byref(a), C.byref(b), C.byref(c)))
        return {"bytes_per_member": a.value, "b_tile": b.value, "free_bytes": c.value}

    def set_batch_tile(self, b_tile=0):
        self._ck(self._lib.kb2_set_batch_tile(self._h, int(b_tile)))

    @property
    def last_batch_tiles(self) -> int:
        return int(self._lib.kb2_last_batch_tiles(self._h))

    def solve_run(self) -> float:
        ms = C.c_float()
        self._ck(self._lib.kb2_solve_run(self._h, C.byref(ms)))
        return float(ms.value)

    PHASES = ["lu", "stage_rhs", "stage_sweeps", "step_end", "jacobian"]

    def get_phase_times(self):
        """Average duration (ms) of each phase kernel of the last solve, from CUDA events around every
        launch of the sampled rounds; launches sampled per phase; rounds the host loop ran."""
        ms = np.zeros(5)
        n = np.zeros(5, dtype=np.int64)
        r = _i64()
        self._ck(self._lib.kb2_get_phase_times(self._h, _f(ms), _i(n), C.byref(r)))
        return {nm: {"ms": float(ms[q]), "sampled_launches": int(n[q])} for q, nm in enumerate(self.PHASES)}, int(r.value)

    def solve_fetch(self, out_u=None, out_umax=None, want_umax=True):
        B, Ns, S = self._B, self._Ns, self.S
        if out_u is None:
            out_u = np.empty((Ns, S, B))
        if out_umax is None and want_umax:
            out_umax = np.empty((S, B))
        status = np.zeros(B, dtype=np.int32)
        stats = np.zeros((B, 8), dtype=np.int64)
        self._ck(self._lib.kb2_solve_fetch(self._h, _f(out_u), _f(out_umax), status.ctypes.data_as(_pi32), _i(stats)))
        return out_u, out_umax, status, stats

    def pack_results_device(self, final_ptr, umax_ptr):
        self._ck(self._lib.kb2_pack_results_device(self._h, C.c_void_p(final_ptr), C.c_void_p(umax_ptr)))

    # ---- multi-GPU ----
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        if load().kb2_comm_unique_id(buf) != 0:
            raise Kb2Error("kb2_comm_unique_id failed (libnccl.so.2 not loadable?)")
        return bytes(buf)

    def comm_init_rank(self, nranks: int, rank: int, uid: bytes):
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        self._ck(self._lib.kb2_comm_init_rank(self._h, nranks, rank, buf))
        self._nranks = nranks

    def allgather_results(self, to_host=True):
        """All-gather of the last solve's final concentrations and per-species maxima over the
        handle's communicator: ([nranks*B, S], [nranks*B, S]) on the host, or None, None with
        to_host=False (results stay on the device, see gathered_device)."""
        n = getattr(self, "_nranks", 1)
        hs = (_H * 1)(self._h)
        if to_host:
            fin = np.empty((n * self._B, self.S))
            mx = np.empty((n * self._B, self.S))
            pf, pm = (_pf64 * 1)(_f(fin)), (_pf64 * 1)(_f(mx))
            self._ck(self._lib.kb2_allgather_results(hs, 1, pf, pm))
            return fin, mx
        self._ck(self._lib.kb2_allgather_results(hs, 1, None, None))
        return None, None

    def gathered_device(self):
        a, b, ms, r, n = C.c_void_p(), C.c_void_p(), C.c_float(), _i32(), _i32()
        self._ck(self._lib.kb2_gathered_device(self._h, C.byref(a), C.byref(b), C.byref(ms), C.byref(r), C.byref(n)))
        return {"final_ptr": a.value, "umax_ptr": b.value, "gather_ms": float(ms.value), "rank": r.value, "nranks": n.value}

    # ---- kernel-level ----
    def eval_k(self, T):
        T = _c64(np.atleast_1d(T))
        out = np.empty((self.R, len(T)))
        self._ck(self._lib.kb2_eval_k(self._h, len(T), _f(T), _f(out)))
        return out

    def eval_profile(self, B, t):
        t = _c64(np.atleast_1d(t))
        out = np.empty((B, len(t)))
        self._ck(self._lib.kb2_eval_profile(self._h, B, len(t), _f(t), _f(out)))
        return out

    def eval_rhs(self, u, k):
        u, k = _c64(u), _c64(k)
        out = np.empty_like(u)
        self._ck(self._lib.kb2_eval_rhs(self._h, u.shape[1], _f(u), _f(k), _f(out)))
        return out

    def eval_jac(self, u, k):
        u, k = _c64(u), _c64(k)
        out = np.empty((max(self.nnzJ, 1), u.shape[1]))
        self._ck(self._lib.kb2_eval_jac(self._h, u.shape[1], _f(u), _f(k), _f(out)))
        return out[:self.nnzJ]

    def factor(self, u, k, hg_inv, want_lu=True):
        u, k, hg = _c64(u), _c64(k), _c64(hg_inv)
        out = np.empty((self.nnzLU, u.shape[1])) if want_lu else None
        self._ck(self._lib.kb2_factor(self._h, u.shape[1], _f(u), _f(k), _f(hg), _f(out)))
        return out

    def trisolve(self, rhs):
        rhs = _c64(rhs)
        out = np.empty_like(rhs)
        self._ck(self._lib.kb2_trisolve(self._h, rhs.shape[1], _f(rhs), _f(out)))
        return out

    def measure_fp64_peak(self) -> float:
        v = _f64()
        self._ck(self._lib.kb2_measure_fp64_peak(self._h, C.byref(v)))
        return float(v.value)

    def time_kernel(self, which, B, iters=10) -> float:
        ms = C.c_float()
        self._ck(self._lib.kb2_time_kernel(self._h, which, B, iters, C.byref(ms)))
        return float(ms.value)
