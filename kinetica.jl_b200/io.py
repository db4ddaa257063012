"""`save_output` / `load_output` — the reference's on-disk format for solve results
(src/analysis/io.jl:50-255): the `ODESolveOutput` is broken down into a dictionary tree of base
arrays and values and written as BSON with BSON.jl's Julia tagging, so that results of B200 solves
can be read by `Kinetica.load_output` / `BSON.load` and the reference's files can be read here.

Tagging follows BSON.jl's lowering: a `Dict{Symbol,Any}` is a plain document; `Symbol` ->
`{tag: "symbol", name}`; arrays of bits types -> `{tag: "array", type: {tag: "datatype", name:
["Core", "Float64"], params: []}, size, data: <binary>}` (the form pinned by the reference's shipped
fixture examples/getting_started/arrhenius_params.bson); arrays of anything else -> the same with
`data` a BSON array of the lowered elements; tuples -> `{tag: "tuple", data}`; other dictionaries
-> `{tag: "dict", data: [keys, values]}`; `nothing` -> null.  Parity status: the round trip and the
bits-array form are tested here; files written by this module have NOT been read back by Julia
(no julia binary in this image).
"""
from __future__ import annotations

import struct
from collections import OrderedDict

import numpy as np

from .conditions import (ConditionSet, DoubleRampGradientProfile, LinearDirectProfile, LinearGradientProfile,
                         NullDirectProfile, NullGradientProfile, StaticConditionProfile, _Sol, isstatic)
from .network import RxData, SpeciesData
from .params import ODESimulationParams

KINETICA_VERSION = (0, 7, 2)


class Symbol(str):
    """a Julia Symbol (keys of the tree are written as plain document keys, values as tagged symbols)"""


# ---------------------------------------------------------------- BSON container
def _cstr(s):
    return s.encode() + b"\x00"


def _enc_value(key, v):
    k = _cstr(key)
    if v is None:
        return b"\x0a" + k
    if isinstance(v, bool):
        return b"\x08" + k + (b"\x01" if v else b"\x00")
    if isinstance(v, (int, np.integer)):
        return b"\x12" + k + struct.pack("<q", int(v))
    if isinstance(v, (float, np.floating)):
        return b"\x01" + k + struct.pack("<d", float(v))
    if isinstance(v, str):
        b = v.encode() + b"\x00"
        return b"\x02" + k + struct.pack("<i", len(b)) + b
    if isinstance(v, (bytes, bytearray)):
        return b"\x05" + k + struct.pack("<i", len(v)) + b"\x00" + bytes(v)
    if isinstance(v, dict):
        return b"\x03" + k + _enc_doc(v)
    if isinstance(v, (list, tuple)):
        return b"\x04" + k + _enc_doc({str(i): x for i, x in enumerate(v)})
    raise TypeError(f"cannot encode {type(v)} as BSON")


def _enc_doc(d):
    body = b"".join(_enc_value(str(k), v) for k, v in d.items())
    return struct.pack("<i", len(body) + 5) + body + b"\x00"


def _dec_doc(buf, off=0):
    (size,) = struct.unpack_from("<i", buf, off)
    end = off + size - 1
    off += 4
    out = OrderedDict()
    while off < end:
        ty = buf[off]; off += 1
        z = buf.index(b"\x00", off)
        key = buf[off:z].decode(); off = z + 1
        if ty in (3, 4):
            val, off = _dec_doc(buf, off)
            if ty == 4:
                val = [val[str(i)] for i in range(len(val))]
        elif ty == 2:
            (n,) = struct.unpack_from("<i", buf, off)
            val = buf[off + 4:off + 4 + n - 1].decode(); off += 4 + n
        elif ty == 5:
            (n,) = struct.unpack_from("<i", buf, off)
            val = bytes(buf[off + 5:off + 5 + n]); off += 5 + n
        elif ty == 18:
            (val,) = struct.unpack_from("<q", buf, off); off += 8
        elif ty == 16:
            (val,) = struct.unpack_from("<i", buf, off); off += 4
        elif ty == 1:
            (val,) = struct.unpack_from("<d", buf, off); off += 8
        elif ty == 8:
            val = buf[off] != 0; off += 1
        elif ty == 10:
            val = None
        else:
            raise ValueError(f"unsupported BSON type {ty}")
        out[key] = val
    return out, end + 1


# ---------------------------------------------------------------- Julia tagging (BSON.jl lowering)
_BITS = {np.dtype("float64"): "Float64", np.dtype("int64"): "Int64", np.dtype("uint8"): "UInt8", np.dtype("int32"): "Int32"}
_BITS_INV = {v: k for k, v in _BITS.items()}


def _dtype_tag(name, params=()):
    return {"tag": "datatype", "name": ["Core", name], "params": list(params)}


def lower(x):
    """Python value -> BSON.jl's tagged form."""
    if x is None or isinstance(x, (bool, float, np.floating, Raw)):
        return x
    if isinstance(x, Symbol):
        return {"tag": "symbol", "name": str(x)}
    if isinstance(x, (int, np.integer, str)):
        return x
    if isinstance(x, np.ndarray) and x.dtype in _BITS:
        return {"tag": "array", "type": _dtype_tag(_BITS[x.dtype]), "size": [int(n) for n in x.shape],
                "data": np.asfortranarray(x).tobytes(order="F")}
    if isinstance(x, tuple):
        return {"tag": "tuple", "data": [lower(v) for v in x]}
    if isinstance(x, (list, np.ndarray)):
        return {"tag": "array", "type": _dtype_tag("Any"), "size": [len(x)], "data": [lower(v) for v in x]}
    if isinstance(x, dict):
        if all(isinstance(k, str) and not isinstance(k, Symbol) for k in x) and getattr(x, "julia_symbol_keys", True) \
                and not getattr(x, "plain_dict", False):
            return {k: lower(v) for k, v in x.items()}                         # Dict{Symbol,Any}: a plain document
        return {"tag": "dict", "data": [[lower(k) for k in x.keys()], [lower(v) for v in x.values()]]}
    raise TypeError(f"cannot lower {type(x)}")


class Raw(dict):
    """an already tagged BSON node: written as it is"""


class JuliaDict(dict):
    """a dictionary whose keys are NOT symbols (Dict{String,Int} ...): written with the `dict` tag"""
    plain_dict = True


def raise_(d):
    """BSON.jl's tagged form -> Python value."""
    if isinstance(d, list):
        return [raise_(v) for v in d]
    if not isinstance(d, dict):
        return d
    tag = d.get("tag")
    if tag == "symbol":
        return Symbol(d["name"])
    if tag == "tuple":
        return tuple(raise_(v) for v in d["data"])
    if tag == "array":
        name = d["type"].get("name", ["Core", "Any"])[-1]
        if isinstance(d["data"], (bytes, bytearray)):
            return np.frombuffer(d["data"], dtype=_BITS_INV[name]).reshape([int(n) for n in d["size"]], order="F").copy()
        return [raise_(v) for v in d["data"]]
    if tag == "dict":
        keys, vals = d["data"]
        return JuliaDict((raise_(k), raise_(v)) for k, v in zip(keys, vals))
    if tag == "struct" and d.get("type", {}).get("name", [""])[-1] == "VersionNumber":
        return tuple(raise_(v) for v in d["data"][:3])
    return OrderedDict((k, raise_(v)) for k, v in d.items())


def bson_dump(path, tree):
    with open(path, "wb") as f:
        f.write(_enc_doc(lower(tree)))


def bson_load(path):
    doc, _ = _dec_doc(open(path, "rb").read())
    return raise_(doc)


# ---------------------------------------------------------------- save_output / load_output
_PROFILE_TYPES = {c.__name__: c for c in (StaticConditionProfile, NullDirectProfile, LinearDirectProfile, NullGradientProfile,
                                          LinearGradientProfile, DoubleRampGradientProfile)}


def _vecs(rows):
    """Vector{Vector{T}}: a list of bits arrays"""
    return [np.asarray(r) for r in rows]


def save_output(out, saveto: str):
    """io.jl:70-168.  1-based indices on disk (the reference's), 0-based in memory."""
    sol_vcs = None if out.sol_vcs is None else {str(s): _vecs([[float(x)] for x in v.u]) for s, v in out.sol_vcs.items()}
    sol_k = None if out.sol_k is None else {"u": _vecs(out.sol_k.u), "t": np.asarray(out.sol_k.t, dtype=np.float64)}
    profiles = []
    for prof in out.conditions.profiles:
        name = type(prof).__name__
        if isstatic(prof):
            profiles.append({"pType": Symbol(name), "value": float(prof.value)})
            continue
        pd = OrderedDict()
        for k, v in vars(prof).items():
            if k == "sol":
                continue
            pd[k] = np.asarray(v, dtype=np.float64) if isinstance(v, np.ndarray) else v
        if getattr(prof, "sol", None) is not None:
            pd["sol"] = {"u": _vecs([[float(x)] for x in prof.sol.u]), "t": np.asarray(prof.sol.t, dtype=np.float64)}
        pd["pType"] = Symbol(name)
        pd["grad" if "Gradient" in name else "f"] = None
        profiles.append(pd)
    sd, rd, pars = out.sd, out.rd, out.pars
    tree = {
        "KineticaCoreVersion": Raw({"tag": "struct", "type": {"tag": "datatype", "name": ["Base", "VersionNumber"], "params": []},
                                    "data": list(KINETICA_VERSION) + [{"tag": "tuple", "data": []}, {"tag": "tuple", "data": []}]}),
        "sd": {"toInt": JuliaDict((k, int(v) + 1) for k, v in sd.toInt.items()), "n": int(sd.n),
               "xyz": JuliaDict(), "level_found": JuliaDict((i + 1, 1) for i in range(sd.n))},
        "rd": {"nr": int(rd.nr), "mapped_rxns": [],
               "id_reacs": _vecs([np.asarray(r, dtype=np.int64) + 1 for r in rd.id_reacs]),
               "id_prods": _vecs([np.asarray(r, dtype=np.int64) + 1 for r in rd.id_prods]),
               "stoic_reacs": _vecs([np.asarray(r, dtype=np.int64) for r in rd.stoic_reacs]),
               "stoic_prods": _vecs([np.asarray(r, dtype=np.int64) for r in rd.stoic_prods]),
               "dH": np.zeros(rd.nr), "rhash": [], "level_found": np.ones(rd.nr, dtype=np.int64)},
        "pars": {"tspan": tuple(float(x) for x in pars.tspan),
                 "u0": JuliaDict(pars.u0) if isinstance(pars.u0, dict) else np.asarray(pars.u0, dtype=np.float64),
                 "solver": Symbol(type(pars.solver).__name__), "jac": bool(pars.jac), "sparse": bool(pars.sparse),
                 "adaptive_tols": bool(pars.adaptive_tols), "update_tols": bool(pars.update_tols),
                 "solve_chunks": bool(pars.solve_chunks), "solve_chunkstep": float(pars.solve_chunkstep),
                 "maxiters": int(pars.maxiters), "ban_negatives": bool(pars.ban_negatives), "progress": bool(pars.progress),
                 "save_interval": None if pars.save_interval is None else float(pars.save_interval),
                 "low_k_cutoff": Symbol(pars.low_k_cutoff) if isinstance(pars.low_k_cutoff, str) else float(pars.low_k_cutoff),
                 "allow_short_u0": bool(pars.allow_short_u0)},
        "sol": {"u": _vecs(out.sol.u), "t": np.asarray(out.sol.t, dtype=np.float64), "vcs": sol_vcs, "k": sol_k},
        "conditions": {"symbols": [Symbol(s) for s in out.conditions.symbols], "profiles": profiles,
                       "discrete_updates": bool(out.conditions.discrete_updates),
                       "ts_update": None if out.conditions.ts_update is None else float(out.conditions.ts_update)},
    }
    bson_dump(saveto, tree)


def load_output(outfile: str):
    """io.jl:180-250 -> ODESolveOutput (profile functions are not stored, like in the reference: the
    profiles are rebuilt from their parameters, so here they work again)."""
    from .solve import ODESolveOutput, RateSolution, Solution
    d = bson_load(outfile)
    toInt = {k: int(v) - 1 for k, v in d["sd"]["toInt"].items()}
    sd = SpeciesData(sorted(toInt, key=toInt.get))
    rdd = d["rd"]
    rd = RxData([list(np.asarray(r) - 1) for r in rdd["id_reacs"]], [list(np.asarray(r) - 1) for r in rdd["id_prods"]],
                [list(np.asarray(r)) for r in rdd["stoic_reacs"]], [list(np.asarray(r)) for r in rdd["stoic_prods"]])
    pd = dict(d["pars"])
    pd.pop("solver", None)
    u0 = pd.pop("u0")
    lk = pd.pop("low_k_cutoff")
    pars = ODESimulationParams(u0=dict(u0) if isinstance(u0, dict) else np.asarray(u0), tspan=tuple(pd.pop("tspan")),
                               low_k_cutoff=str(lk) if isinstance(lk, str) else lk, **pd)
    profs = []
    for p in d["conditions"]["profiles"]:
        p = dict(p)
        cls = _PROFILE_TYPES[str(p.pop("pType")).split(".")[-1]]
        if cls is StaticConditionProfile:
            profs.append(StaticConditionProfile(p["value"]))
            continue
        sol = p.pop("sol", None)
        obj = cls.__new__(cls)
        for k, v in p.items():
            if k not in ("f", "grad"):
                setattr(obj, k, v)
        obj.sol = None if sol is None else _Sol(sol["t"], np.array([float(np.asarray(x)[0]) for x in sol["u"]]))
        profs.append(obj)
    cs = ConditionSet.__new__(ConditionSet)
    cs.symbols = [str(s) for s in d["conditions"]["symbols"]]
    cs.profiles = profs
    cs.discrete_updates = bool(d["conditions"]["discrete_updates"])
    cs.ts_update = d["conditions"]["ts_update"]
    sol = Solution(t=np.asarray(d["sol"]["t"]), u=[np.asarray(x) for x in d["sol"]["u"]])
    k = d["sol"]["k"]
    sol_k = None if k is None else RateSolution(k["t"], np.array([np.asarray(x) for x in k["u"]]))
    vcs = d["sol"]["vcs"]
    sol_vcs = None if vcs is None else {str(s): _Sol(sol.t, np.array([float(np.asarray(x)[0]) for x in v])) for s, v in vcs.items()}
    U = np.array(sol.u)
    return ODESolveOutput(sd=sd, rd=rd, sol=sol, sol_k=sol_k, sol_vcs=sol_vcs, pars=pars, conditions=cs, umax=U.max(axis=0))
