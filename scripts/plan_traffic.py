"""Where the factorisation's U' traffic goes, from the block plan alone (no GPU): doubles of U' loaded per
LU (sum over tasks of targets x source rows), against the storage, and how much of it two consecutive
panels would share if they were updated together from one load (the 16-row super-panel idea)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kinetica_b200 import _lib
from kinetica_b200.synthetic import synthetic_crn, SEED_BASE

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
cid = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sd, rd, Ea, A = synthetic_crn(S, 5 * S, SEED_BASE + cid)
h = _lib.Handle(-1)
h.set_network(S, *rd.flatten())
print("symbolic", h.symbolic(4), h.get_plan_stats())
P = h.get_plan()
cols, cptr, width, nrows, nxt = P["cols"], P["p_cptr"], P["p_width"], P["p_nrows"], P["p_next"]
npan = len(nrows)
padded = int(sum(int(width[p]) * int(nrows[p]) for p in range(npan)))
# per target panel: set of (source panel, global column) pairs it loads U' for
loads = [dict() for _ in range(npan)]
tot_u = tot_l = 0
for (Pn, x0, x1, task0, ntask, dmode, *_rest) in P["u_info"]:
    gcols = cols[cptr[Pn]: cptr[Pn] + width[Pn]]
    for tk in range(task0, task0 + ntask):
        Q, lp, ntg, map0, nq = P["t_info"][tk][:5]
        inch = lp >> 30
        if not inch:
            tot_l += int(nq) * int(nrows[Pn])
        for t in range(ntg):
            e = P["map"][map0 + t]
            gc = int(gcols[x0 + (e >> 16)])
            loads[Pn][(int(Q), gc)] = int(nq)
            tot_u += int(nq)
shared = 0
for p in range(0, npan - 1, 2):
    a, b = loads[p], loads[p + 1]
    shared += sum(nq for key, nq in a.items() if key in b)
def saved(group):
    sv = 0
    for p in range(0, npan, group):
        seen = {}
        for q in range(p, min(p + group, npan)):
            for key, nq in loads[q].items():
                if key in seen:
                    sv += nq
                seen[key] = nq
    return sv
print(f"padded storage        {padded:10d} doubles per member")
print(f"U' values loaded      {tot_u:10d}  ({tot_u / padded:.2f}x the storage)")
print(f"L' blocks re-staged   {tot_l:10d}  ({tot_l / padded:.2f}x)")
for grp in (2, 4, 8):
    sv = saved(grp)
    print(f"loaded once per group of {grp} panels ({8 * grp:2d} rows): {sv:10d} saved ({100.0 * sv / max(tot_u, 1):.0f} % of the U' loads)")
uniq = len({key for d in loads for key in d})
print(f"distinct (source, column) pairs: {uniq} -> a perfect cache would load {sum(nq for d in [dict((k, v) for dd in loads for k, v in dd.items())] for nq in d.values())} doubles")
