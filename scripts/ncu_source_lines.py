"""Per-source-line view of an .ncu-rep captured with `--set full --import-source on` (kernels built with
-lineinfo): warp-sampling counts and executed instructions summed over the SASS of every CUDA line, the
hottest lines first, plus the sums over named line ranges (phases).
Usage: python scripts/ncu_source_lines.py gpurun_out/x.ncu-rep kinetica.jl_b200/csrc/kb2_front.cuh [name=first-last ...]"""
import collections
import csv
import subprocess
import sys


def main():
    rep, src = sys.argv[1], sys.argv[2]
    ranges = []
    for a in sys.argv[3:]:
        name, r = a.split("=")
        lo, hi = r.split("-")
        ranges.append((name, int(lo), int(hi)))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    text = open(src).read().split("\n")
    cur = None
    agg = collections.defaultdict(lambda: [0, 0])
    for r in rows[3:]:
        if len(r) < 8:
            continue
        if r[0] != "":
            cur = int(r[0]) if r[0].isdigit() else cur
            continue
        if r[2].startswith("0x"):
            try:
                agg[cur][0] += int(r[4]); agg[cur][1] += int(r[7])
            except ValueError:
                pass
    tot = sum(a[0] for a in agg.values()) or 1
    toti = sum(a[1] for a in agg.values()) or 1
    print("# %s: %d warp samples, %d warp instructions; lines of %s" % (rep, tot, toti, src))
    if ranges:
        print("# phases (line ranges of the source):")
        for name, lo, hi in ranges:
            sm = sum(a[0] for k, a in agg.items() if k is not None and lo <= k <= hi)
            im = sum(a[1] for k, a in agg.items() if k is not None and lo <= k <= hi)
            print("#   %-44s lines %4d-%4d  samples %5.1f %%  instructions %5.1f %%" % (name, lo, hi, 100.0 * sm / tot, 100.0 * im / toti))
    print("# line  samples%  instr%  source")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
        line = text[k - 1].strip()[:110] if k and k <= len(text) else ""
        print("%5s  %6.1f  %6.1f  %s" % (k, 100.0 * a[0] / tot, 100.0 * a[1] / toti, line))


if __name__ == "__main__":
    main()
