"""Turn an ncu metrics CSV of a short C3 solve into profiles/r02_traffic.json: average DRAM bytes per
launch (dram__bytes_read.sum + dram__bytes_write.sum) and average duration of every phase kernel,
keyed to the hashes of the kernel sources and to the block plan it was captured on (bench.py drops a
phase's number when a source it depends on, or the plan, changes).

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        --csv --log-file gpurun_out/r2_traffic.csv -k regex:'k_step|k_stage|k_lu' -c 450 \
        python scripts/profile_solve.py 4096 0.02
    python scripts/ncu_traffic.py gpurun_out/r2_traffic.csv c3 4096 <padded slots of the plan in use>
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_source_hashes         # noqa: E402

PHASE = {"k_step_jac": "jacobian", "k_lu_window": "lu", "k_step_lu": "lu", "k_stage_rhs": "stage_rhs",
         "k_stage_sweep": "stage_sweeps", "k_step_end": "step_end"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "second": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}


def main(path, workload, members, padded):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    acc = {}
    for r in rows[1:]:
        name = r[ix["Kernel Name"]]
        ph = next((v for k, v in PHASE.items() if k in name), None)
        if ph is None:
            continue
        val = float(r[ix["Metric Value"]].replace(",", "")) * UNIT.get(r[ix["Metric Unit"]], 1.0)
        a = acc.setdefault(ph, {"launches": set(), "read": 0.0, "write": 0.0, "ms": 0.0})
        a["launches"].add(r[ix["ID"]])
        m = r[ix["Metric Name"]]
        if m == "dram__bytes_read.sum":
            a["read"] += val
        elif m == "dram__bytes_write.sum":
            a["write"] += val
        elif m == "gpu__time_duration.sum":
            a["ms"] += val
    out = {"source_hashes": kernel_source_hashes(), "plan": {"padded": int(padded)}, "workload": workload, "members": int(members),
           "from": os.path.basename(path),
           "dram_bytes_per_launch": {}, "detail": {}}
    for ph, a in acc.items():
        n = len(a["launches"])
        out["dram_bytes_per_launch"][ph] = (a["read"] + a["write"]) / n
        out["detail"][ph] = {"launches": n, "read_per_launch": a["read"] / n, "write_per_launch": a["write"] / n,
                             "ms_per_launch_under_ncu": a["ms"] / n}
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "c3", sys.argv[3] if len(sys.argv) > 3 else 4096,
         sys.argv[4] if len(sys.argv) > 4 else 0)
