"""The Julia shim (julia/KineticaB200.jl) and the ccall snippets of INTEGRATION.md cannot be executed here
(no julia binary), so their `ccall` signatures are checked statically against the prototypes of
include/kinetica_b200.h: symbol exists, return type, number of arguments and every argument's type class
(handle, int32, int64, double, pointer to one of them, out-handle, C string)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _c_prototypes():
    text = open(os.path.join(ROOT, "include", "kinetica_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = "\n".join(l for l in text.split("\n") if not l.lstrip().startswith("#") and "extern" not in l)
    protos = {}
    for stmt in text.split(";"):
        m = re.search(r"([\w\s\*]+?)\b(kb2_\w+)\s*\(([^)]*)\)\s*$", stmt.strip(), flags=re.S)
        if not m or "typedef" in m.group(1):
            continue
        ret, name, args = " ".join(m.group(1).split()), m.group(2), m.group(3)
        protos[name] = (_c_class(ret + " f"), [_c_class(" ".join(a.split())) for a in args.split(",") if a.strip() and a.strip() != "void"])
    return protos


def _c_class(decl):
    """type class of a C parameter declaration `type [*]name`"""
    d = decl.replace("const", " ").strip()
    ptr = d.count("*")
    d = d.replace("*", " ")
    base = " ".join(d.split()[:-1]) if len(d.split()) > 1 else d.strip()      # drop the parameter name
    if base == "kb2_handle":
        return "handle*" if ptr else "handle"
    if base == "char":
        return "cstring"
    if base == "void":
        return "voidptr" if ptr else "void"
    name = {"int32_t": "i32", "int64_t": "i64", "double": "f64", "float": "f32", "long long": "i64", "uint8_t": "u8", "size_t": "usize"}[base]
    return name + "*" * ptr


def _jl_class(t):
    t = t.strip()
    if t == "Ptr{Cvoid}":
        return "handle"
    if t in ("Ref{Ptr{Cvoid}}", "Ptr{Ptr{Cvoid}}"):
        return "handle*"
    if t == "Cstring":
        return "cstring"
    m = re.fullmatch(r"(?:Ptr|Ref)\{(\w+)\}", t)
    base = {"Int32": "i32", "Int64": "i64", "Float64": "f64", "Float32": "f32", "Cint": "i32", "Cdouble": "f64"}
    if m:
        return base[m.group(1)] + "*"
    return base[t]


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "{([":
            depth += 1
        if ch in "})]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


def _ccalls(path):
    text = open(path).read()
    calls = []
    for m in re.finditer(r"ccall\(\(:(kb2_\w+),\s*LIB\),\s*([\w\{\}]+),\s*\(", text):
        i, depth = m.end(), 1
        while depth:                      # the argument-type tuple
            depth += {"(": 1, ")": -1}.get(text[i], 0)
            i += 1
        types = [t for t in _split_top(text[m.end():i - 1]) if t.strip()]
        j, depth = i, 1                   # the rest of the call: the actual arguments
        while depth:
            depth += {"(": 1, ")": -1}.get(text[j], 0)
            j += 1
        nargs = len([a for a in _split_top(text[i:j - 1].lstrip(", \n")) if a.strip()])
        calls.append((m.group(1), m.group(2), types, nargs, text.count("\n", 0, m.start()) + 1))
    return calls


@pytest.mark.parametrize("path", ["julia/KineticaB200.jl", "INTEGRATION.md"])
def test_ccall_signatures_match_the_header(path):
    protos = _c_prototypes()
    assert len(protos) > 40 and protos["kb2_create"] == ("i32", ["i32", "handle*"])
    calls = _ccalls(os.path.join(ROOT, path))
    assert calls, path
    for name, ret, types, nargs, line in calls:
        assert name in protos, f"{path}:{line}: {name} is not declared in include/kinetica_b200.h"
        cret, cargs = protos[name]
        assert _jl_class(ret) == cret, f"{path}:{line}: {name} returns {cret}, ccall says {ret}"
        assert [_jl_class(t) for t in types] == cargs, f"{path}:{line}: {name}{cargs} vs ccall {types}"
        assert nargs == len(types), f"{path}:{line}: {name}: {len(types)} argument types but {nargs} arguments"


def test_shim_binds_the_entry_points_of_the_path():
    names = {c[0] for c in _ccalls(os.path.join(ROOT, "julia", "KineticaB200.jl"))}
    assert {"kb2_create", "kb2_destroy", "kb2_last_error", "kb2_set_network", "kb2_symbolic", "kb2_set_arrhenius",
            "kb2_set_rate_table", "kb2_set_profiles", "kb2_set_member_stops", "kb2_set_chunking", "kb2_solve"} <= names
