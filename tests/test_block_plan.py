"""Host-side check of the block plan (kb2_panel.cpp): a numpy interpreter executes the plan
tables exactly the way the CUDA warp does (units, source-block tasks, target maps, diagonal
modes) on one random member, and the result must be the Crout LU of the permuted matrix.
Needs no GPU: the plan comes from a host-only handle."""
import numpy as np
import pytest

from kinetica_b200 import _lib
from kinetica_b200.synthetic import synthetic_crn, SEED_BASE
from oracle import kinetica_oracle as ko


def run_plan(plan, lu):
    """In-place block Crout LU over the padded storage `lu` (one member), following
    tile_lu in kinetica.jl_b200/csrc/kb2_kernels.cuh step for step."""
    P_ = plan
    invd = {}
    for (P, x0, x1, task0, ntask, dmode, fi_off, fi_nq, nr_, nxt_, p0_, base_) in P_["u_info"]:
        nr, nxt, base, p0 = P_["p_nrows"][P], P_["p_next"][P], P_["p_base"][P], P_["p_row0"][P]
        assert (nr_, nxt_, p0_, base_) == (nr, nxt, p0, base)
        cw = x1 - x0
        inch_seen = []
        Wp = lu[base + x0 * nr: base + x1 * nr].reshape(cw, nr).copy()          # [c][r]
        for tk in range(task0, task0 + ntask):
            Q, lp, ntg, map0, nq_, bq_, uqq_, nx_off, nx_nq, pf_off, pf_len, _ = P_["t_info"][tk]
            lpos, inch = lp & 0x3fffffff, lp >> 30
            nq, nxq, bq = P_["p_nrows"][Q], P_["p_next"][Q], P_["p_base"][Q]
            assert (nq_, bq_, uqq_) == (nq, bq, bq + nxq * nq)
            # staging links: the next in-chunk source after this task
            later = [P_["t_info"][z] for z in range(tk + 1, task0 + ntask) if P_["t_info"][z][1] >> 30]
            assert (nx_off, nx_nq) == ((later[0][6], later[0][4]) if later else (-1, 0))
            if inch:
                inch_seen.append((uqq_, nq))
            if ntg:
                pqs = [P_["map"][map0 + t] & 0xffff for t in range(ntg)]
                assert pf_off == bq + min(pqs) * nq and pf_len == (max(pqs) + 1 - min(pqs)) * nq
            if inch:
                X = Wp[lpos - x0: lpos - x0 + nq].T.copy()                        # [r][q]
                Uqq = lu[bq + nxq * nq: bq + (nxq + nq) * nq].reshape(nq, nq).T     # [a][q]
                for a in range(nq - 1):
                    for q in range(a + 1, nq):
                        X[:, q] -= X[:, a] * Uqq[a, q]
                Wp[lpos - x0: lpos - x0 + nq] = X.T
                L = X
            else:
                L = lu[base + lpos * nr: base + (lpos + nq) * nr].reshape(nq, nr).T   # [r][q]
            for t in range(ntg):
                e = P_["map"][map0 + t]
                pq, pp = e & 0xffff, e >> 16
                assert 0 <= pp < cw
                u = lu[bq + pq * nq: bq + (pq + 1) * nq]
                Wp[pp] -= L @ u
        assert (fi_off, fi_nq) == (inch_seen[0] if inch_seen else (-1, 0))
        if dmode:
            dpos = nxt - x0
            if dmode == 1:
                D = Wp[dpos: dpos + nr].T.copy()                                   # [r][j]
                for j in range(nr):
                    inv = 1.0 / D[j, j]
                    invd[p0 + j] = inv
                    D[j, j + 1:] *= inv
                    for r in range(j + 1, nr):
                        D[r, j + 1:] -= D[r, j] * D[j, j + 1:]
                Wp[dpos: dpos + nr] = D.T
                Lpp = D
                c0 = dpos + nr
            else:
                Lpp = lu[base + nxt * nr: base + (nxt + nr) * nr].reshape(nr, nr).T
                c0 = 0
            for t in range(c0, cw):
                w = Wp[t]
                for r in range(nr):
                    for a in range(r):
                        w[r] -= Lpp[r, a] * w[a]
                    w[r] *= invd[p0 + r]
        lu[base + x0 * nr: base + x1 * nr] = Wp.reshape(-1)
    return invd


@pytest.mark.parametrize("S,R,ordering", [(96, 400, 0), (200, 1000, 3), (200, 1000, 0), (420, 2100, 3), (30, 60, 1), (64, 256, 4)])
def test_block_plan_reproduces_crout_lu(S, R, ordering):
    sd, rd, Ea, A = synthetic_crn(S, R, SEED_BASE + 50 + S)
    h = _lib.Handle(-1)
    h.set_network(S, *rd.flatten())
    h.symbolic(ordering)
    plan = h.get_plan()
    st = h.get_plan_stats()
    rowptr, colidx, diagpos = h.get_lu_pattern()
    perm = h.get_ordering()
    net = ko.Network(S, rd.id_reacs, rd.id_prods, rd.stoic_reacs, rd.stoic_prods)
    rng = np.random.default_rng(S)
    u = rng.uniform(0, 1, S)
    k = 10 ** rng.uniform(-3, 3, R)
    W = np.eye(S) * 50.0 - net.jac_dense(u, k)
    Wp = W[np.ix_(perm, perm)]
    # structure of the plan
    assert st["padded"] == int(np.sum(plan["p_nrows"] * plan["p_width"]))
    assert len(set(plan["slot_of"].tolist())) == len(plan["slot_of"])          # exact entries get distinct slots
    if S == 420:
        assert st["max_width"] > 96 and st["units"] > st["panels"]               # wide panels are chunked
    # assemble the padded storage from the exact pattern
    lu = np.zeros(st["padded"])
    for i in range(S):
        for p in range(rowptr[i], rowptr[i + 1]):
            lu[plan["slot_of"][p]] = Wp[i, colidx[p]]
    # every nonzero of the permuted matrix lies inside the pattern
    mask = np.zeros((S, S), bool)
    for i in range(S):
        mask[i, colidx[rowptr[i]:rowptr[i + 1]]] = True
    assert np.all(Wp[~mask] == 0.0)
    invd = run_plan(plan, lu)
    L = np.zeros((S, S)); U = np.eye(S)
    for i in range(S):
        for p in range(rowptr[i], rowptr[i + 1]):
            j = colidx[p]
            if j <= i:
                L[i, j] = lu[plan["slot_of"][p]]
            else:
                U[i, j] = lu[plan["slot_of"][p]]
    assert np.max(np.abs(L @ U - Wp)) <= 1e-11 * np.max(np.abs(Wp))
    assert np.allclose([invd[i] for i in range(S)], 1.0 / np.diag(L), rtol=1e-14)
    # padding slots stay exactly zero
    exact = np.zeros(st["padded"], bool)
    exact[plan["slot_of"]] = True
    assert np.all(lu[~exact] == 0.0)
    # diagonal and Jacobian slot tables agree with slot_of
    for i in range(S):
        assert plan["diag_slot"][i] == plan["slot_of"][diagpos[i]]
    h.close()
