"""CPU-only exploration of elimination orderings for the panel / front plans: padded storage (what the
triangular sweeps stream), padded FMAs and the front statistics of the window LU (what its update
and strip phases cost) per candidate ordering.  Needs no GPU (host-only handle).
Usage: python scripts/explore_orderings.py [S] [config id]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp
from scipy.sparse.csgraph import reverse_cuthill_mckee
from kinetica_b200 import _lib
from kinetica_b200.synthetic import synthetic_crn, SEED_BASE

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
cid = int(sys.argv[2]) if len(sys.argv) > 2 else 3
R = 5 * S
sd, rd, Ea, A = synthetic_crn(S, R, SEED_BASE + cid)


def stats(ordering=0, perm=None, label=""):
    h = _lib.Handle(-1)
    h.set_network(S, *rd.flatten())
    nnzJ, nnzLU, nfma = h.symbolic(ordering, perm=perm)
    ps = h.get_plan_stats()
    fp = h.get_front_plan()
    f = fp["f_info"]
    nr, nu, nl = f[:, 0], f[:, 2], f[:, 3]
    dwork = int(np.sum(nr.astype(np.int64) * nu * nl))
    dblocks = int(np.sum(((nl + 7) // 8) * ((nu + 3) // 4)))
    strips = int(np.sum(nu + nl))
    print("%-28s nnzLU %7d fma %8d | padded %7d (x%.3f) fma_pad %8d | NF %4d win %3dx%3d D-fma %8d blocks %6d strip tasks %6d hot %3d la %3d"
          % (label, nnzLU, nfma, ps["padded"], ps["padded"] / nnzLU, ps["fma_padded"], fp["NF"], fp["Wr"], fp["Wc"], dwork, dblocks, strips,
             int(f[:, 9].sum()), int(f[:, 10].sum())), flush=True)
    perm_out = h.get_ordering()
    colptr, rowval = h.get_pattern()
    h.close()
    return perm_out, colptr, rowval


p_md, colptr, rowval = stats(0, label="min degree")
p_nat, _, _ = stats(1, label="natural")
p_dl, _, _ = stats(3, label="dense-last natural")

# symmetrised graph
rows = rowval
cols = np.repeat(np.arange(S), np.diff(colptr))
G = sp.coo_matrix((np.ones(len(rows)), (rows, cols)), shape=(S, S)).tocsr()
G = ((G + G.T) > 0).astype(np.int8).tocsr()
deg = np.asarray(G.sum(axis=1)).ravel()
thr = max(32, 8 * int(np.sort(deg)[S // 2]))
dense = np.where(deg > thr)[0]
sparse_ = np.where(deg <= thr)[0]
print("dense species", len(dense), "threshold", thr)
dense_sorted = dense[np.argsort(deg[dense], kind="stable")]

# RCM on the non-dense part, dense last
Gs = G[sparse_][:, sparse_]
rcm = reverse_cuthill_mckee(Gs.tocsr(), symmetric_mode=True)
stats(perm=np.concatenate([sparse_[rcm], dense_sorted]), label="dense-last RCM")
stats(perm=np.concatenate([sparse_[rcm[::-1]], dense_sorted]), label="dense-last CM")

# Fiedler (spectral) order of the non-dense part
try:
    from scipy.sparse.linalg import eigsh
    L = sp.diags(np.asarray(Gs.sum(axis=1)).ravel()) - Gs
    vals, vecs = eigsh(L.astype(float), k=2, sigma=-1e-3, which="LM")
    fied = np.argsort(vecs[:, 1])
    stats(perm=np.concatenate([sparse_[fied], dense_sorted]), label="dense-last spectral")
except Exception as e:
    print("spectral failed", e)

# different dense thresholds with natural order
for t in (16, 24, 48, 64, 100):
    d = np.where(deg > t)[0]
    s_ = np.where(deg <= t)[0]
    stats(perm=np.concatenate([s_, d[np.argsort(deg[d], kind="stable")]]), label="dense-last natural thr %d (%d dense)" % (t, len(d)))
