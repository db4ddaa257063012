"""Regenerates tests/golden/arrhenius_params.json from the reference's shipped fixture
(reference examples/getting_started/arrhenius_params.bson: two Julia-tagged Float64[30]
arrays, `Ea` in J/mol and `A`).  Needs /root/reference; run in the authoring container only:

    python tests/golden/make_golden.py
"""
import json
import os
import struct

SRC = "/root/reference/examples/getting_started/arrhenius_params.bson"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "arrhenius_params.json")


def parse_doc(buf, off=0):
    """Minimal BSON reader: documents (0x03/0x04), strings (0x02), binary (0x05), int64 (0x12)."""
    (size,) = struct.unpack_from("<i", buf, off)
    end = off + size - 1
    off += 4
    out = {}
    while off < end:
        ty = buf[off]; off += 1
        z = buf.index(b"\x00", off)
        key = buf[off:z].decode(); off = z + 1
        if ty in (3, 4):
            val, off = parse_doc(buf, off)
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
        elif ty == 1:
            (val,) = struct.unpack_from("<d", buf, off); off += 8
        else:
            raise ValueError(f"unsupported BSON type {ty}")
        out[key] = val
    return out, end + 1


def main():
    doc, _ = parse_doc(open(SRC, "rb").read())
    res = {}
    for name in ("Ea", "A"):
        arr = doc[name]
        assert arr["tag"] == "array" and arr["type"]["name"] == ["Core", "Float64"]
        n = arr["size"][0]
        res[name] = [float.hex(x) for x in struct.unpack(f"<{n}d", arr["data"])]
    res["_source"] = "reference examples/getting_started/arrhenius_params.bson (hex floats, bit-exact)"
    with open(DST, "w") as f:
        json.dump(res, f, indent=1)
    print("wrote", DST)


if __name__ == "__main__":
    main()
