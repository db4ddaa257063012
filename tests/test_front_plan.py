"""Host-side check of the front plan of the window LU (kb2_front.cpp): a numpy interpreter executes
the plan tables the way k_lu_window does (window slots, init entries, pivot block, strips, rank-nr
update) on one random member, starting from the COMPACT Jacobian values, and must reproduce the
result of the block plan interpreter (tests/test_block_plan.py): both schedules apply the same
updates to every entry in the same order.  Needs no GPU."""
import numpy as np
import pytest

from kinetica_b200 import _lib
from kinetica_b200.synthetic import synthetic_crn, SEED_BASE
from oracle import kinetica_oracle as ko
from test_block_plan import run_plan



def run_fronts(fp, plan, jv, hg_inv, padded):
    """Window LU of one member: returns (lu storage, invd) like run_plan."""
    Wr, Wc = fp["Wr"], fp["Wc"]
    win = np.zeros(Wr * Wc)                        # inactive entries are zero
    lu = np.full(padded, np.nan)
    invd = {}
    lists, init = fp["lists"], fp["init"]
    dead_prev, dead_prev_r, dead_prev_c = False, set(), set()
    finfo = fp["f_info"]

    def orig(src):
        v = -jv[(src >> 1) - 1] if (src >> 1) else 0.0
        return v + hg_inv if src & 1 else v

    Dn = None                                      # look-ahead copy of this front's pivot block (made during the previous front)
    for P, (nr, p0, nu, nl, base, nxt, loff, ioff, icnt, hot, la, _c) in enumerate(finfo):
        # window entries that become live at this front and have an original value
        touches_prev = False
        for pos, src in init[ioff: ioff + icnt]:
            assert win[pos] == 0.0                 # nothing was there before
            win[pos] = orig(src)
            touches_prev |= (pos // Wc in dead_prev_r) or (pos % Wc in dead_prev_c) if dead_prev else False
        assert hot or not touches_prev            # values land in a slot of the previous front only in flagged fronts
        prs, pcs = lists[loff: loff + 8], lists[loff + 8: loff + 16]
        if la:      # look-ahead fronts: no entry of the pivot block is new at this front
            blk = {int(r) * Wc + int(c) for r in prs[:nr] for c in pcs[:nr]}
            assert P >= 1 and not any(int(pos) in blk for pos, _ in init[ioff: ioff + icnt])
        ucs = lists[loff + 16: loff + 16 + nu]
        ujj = lists[loff + 16 + nu: loff + 16 + 2 * nu]
        lrs = lists[loff + 16 + 2 * nu: loff + 16 + 2 * nu + nl]
        lgs = lists[loff + 16 + 2 * nu + nl: loff + 16 + 2 * nu + 2 * nl]
        assert np.all(np.diff(ucs) > 0) and sorted(ujj.tolist()) == list(range(nu))
        assert np.all(prs[:nr] >= 0) and np.all(pcs[:nr] >= 0) and np.all(prs[nr:] < 0)
        # pivot block: Crout, pivots on L, unit diagonal on U (same operation order as tile_lu)
        D = np.array([[win[prs[r] * Wc + pcs[j]] for j in range(nr)] for r in range(nr)])
        if la:      # the copy the kernel factorises was made ahead: it must be the block as it stands now, bit for bit
            assert Dn is not None and Dn.shape == D.shape and np.array_equal(Dn, D)
        for j in range(nr):
            inv = 1.0 / D[j, j]
            invd[p0 + j] = inv
            D[j, j + 1:] *= inv
            for r in range(j + 1, nr):
                D[r, j + 1:] -= D[r, j] * D[j, j + 1:]
        for r in range(nr):
            for j in range(nr):
                lu[base + (nxt + j) * nr + r] = D[r, j]
        # U strip: U'[:, j] = inv(L') w
        U12 = np.zeros((nr, nu))
        for jj in range(nu):
            w = np.array([win[prs[r] * Wc + ucs[jj]] for r in range(nr)])
            for r in range(nr):
                for a in range(r):
                    w[r] -= D[r, a] * w[a]
                w[r] *= invd[p0 + r]
            U12[:, jj] = w
            for r in range(nr):
                win[prs[r] * Wc + ucs[jj]] = w[r]          # strips are computed in place
            lu[base + (nxt + nr + ujj[jj]) * nr: base + (nxt + nr + ujj[jj] + 1) * nr] = w
        # L strip: L'[i, :] = x inv(U'_PP)
        L21 = np.zeros((nl, nr))
        for ii in range(nl):
            X = np.array([win[lrs[ii] * Wc + pcs[q]] for q in range(nr)])
            for a in range(nr - 1):
                for q in range(a + 1, nr):
                    X[q] -= X[a] * D[a, q]
            L21[ii] = X
            for q in range(nr):
                win[lrs[ii] * Wc + pcs[q]] = X[q]
            slot0, stride = lgs[ii] & 0x0fffffff, (lgs[ii] >> 28) + 1
            for q in range(nr):
                lu[slot0 + q * stride] = X[q]
        # look-ahead (k_lu_window, warp 0): the next front's pivot block, copied from the window BEFORE
        # this front's update and brought up to date with this front's strips
        Dn = None
        if P + 1 < len(finfo) and finfo[P + 1][10]:
            nr1, lo1 = int(finfo[P + 1][0]), int(finfo[P + 1][6])
            r1, c1 = lists[lo1: lo1 + 8], lists[lo1 + 8: lo1 + 16]
            Dn = np.array([[win[r1[a] * Wc + c1[b]] for b in range(nr1)] for a in range(nr1)])
            for a in range(nr1):
                for q in range(nr):
                    l = win[r1[a] * Wc + pcs[q]]
                    for b in range(nr1):
                        Dn[a, b] -= l * win[prs[q] * Wc + c1[b]]
        # rank-nr update, pivots in ascending order
        for ii in range(nl):
            for jj in range(nu):
                pos = lrs[ii] * Wc + ucs[jj]
                w = win[pos]
                for q in range(nr):
                    w -= L21[ii, q] * U12[q, jj]
                win[pos] = w
        # the pivot rows and columns are dead from here on: their slots are cleared
        for r in range(nr):
            win[prs[r] * Wc: (prs[r] + 1) * Wc] = 0.0
            win[pcs[r]::Wc] = 0.0
        dead_prev = True
        dead_prev_r, dead_prev_c = set(prs[:nr].tolist()), set(pcs[:nr].tolist())
    return lu, invd


@pytest.mark.parametrize("S,R,ordering", [(96, 400, 0), (200, 1000, 3), (200, 1000, 0), (420, 2100, 3), (30, 60, 1), (64, 256, 4),
                                          (200, 1000, 5), (420, 2100, 6), (200, 1000, 7), (96, 400, 6)])
def test_front_plan_matches_block_plan(S, R, ordering):
    sd, rd, Ea, A = synthetic_crn(S, R, SEED_BASE + 50 + S)
    h = _lib.Handle(-1)
    h.set_network(S, *rd.flatten())
    h.symbolic(ordering)
    plan, fp, st = h.get_plan(), h.get_front_plan(), h.get_plan_stats()
    rowptr, colidx, diagpos = h.get_lu_pattern()
    colptr, rowval = h.get_pattern()
    perm = h.get_ordering()
    net = ko.Network(S, rd.id_reacs, rd.id_prods, rd.stoic_reacs, rd.stoic_prods)
    rng = np.random.default_rng(S)
    u = rng.uniform(0, 1, S)
    k = 10 ** rng.uniform(-3, 3, R)
    J = net.jac_dense(u, k)
    hg = 50.0
    jv = np.array([J[rowval[p], l] for l in range(S) for p in range(colptr[l], colptr[l + 1])])
    # reference: the block plan interpreter on the assembled padded storage
    Wp = (np.eye(S) * hg - J)[np.ix_(perm, perm)]
    ref = np.zeros(st["padded"])
    for i in range(S):
        for p in range(rowptr[i], rowptr[i + 1]):
            ref[plan["slot_of"][p]] = Wp[i, colidx[p]]
    ref_invd = run_plan(plan, ref)
    assert fp["NF"] == st["panels"] and fp["Wr"] > 0 and fp["Wc"] > 0
    assert fp["f_info"][0, 10] == 0 and set(np.unique(fp["f_info"][:, 10])) <= {0, 1}
    lu, invd = run_fronts(fp, plan, jv, hg, st["padded"])
    assert not np.any(np.isnan(lu))                # every storage slot is written (exactly once)
    # same updates in the same order (the interpreters differ in how they round a block product;
    # the CUDA kernels are compared bit for bit in tests/test_gpu_kernels.py)
    scale = np.max(np.abs(ref))
    assert np.max(np.abs(lu - ref)) <= 1e-12 * scale
    exact = np.zeros(st["padded"], bool)
    exact[plan["slot_of"]] = True
    assert np.all(lu[~exact] == 0.0)              # padding stays exactly zero
    assert np.allclose([invd[i] for i in range(S)], [ref_invd[i] for i in range(S)], rtol=1e-12)
    h.close()


def test_front_window_of_the_bench_network_fits_shared_memory():
    """C3 (1k species / 5k reactions): the window of four members plus the strip buffers must fit
    the 227 KB of shared memory a CTA can have on sm_100."""
    sd, rd, Ea, A = synthetic_crn(1000, 5000, SEED_BASE + 3)
    h = _lib.Handle(-1)
    h.set_network(1000, *rd.flatten())
    h.symbolic(4)
    fp = h.get_front_plan()
    MB = 4
    need = 8 * MB * (fp["Wr"] * fp["Wc"] + 64 + 8) + 8 * (16 + 2 * fp["max_nu"] + 2 * fp["max_nl"]) + 8 * 4 * 256
    assert need <= 227 * 1024, (fp["Wr"], fp["Wc"], fp["max_nl"], fp["max_nu"], need)
    h.close()


@pytest.mark.parametrize("S,seed", [(1000, 3), (5000, 5)])
def test_plans_of_the_baseline_networks_factorise_correctly(S, seed):
    """C3 and C5 (BASELINE configs 3 and 5) with the ordering `auto` chooses for them (Sloan 1:2 / 2:1): the
    front plan and the block plan, interpreted in numpy, give the same factors, L U = P W P^T and a linear
    system solved through the factors agrees with numpy's — at the networks' full size, on the host."""
    import scipy.linalg as sl
    R = 5 * S
    sd, rd, Ea, A = synthetic_crn(S, R, SEED_BASE + seed)
    h = _lib.Handle(-1)
    h.set_network(S, *rd.flatten())
    h.symbolic(4)
    plan, fp, st = h.get_plan(), h.get_front_plan(), h.get_plan_stats()
    assert st["ordering"] in (6, 7) and fp["Wr"] * fp["Wc"] * 4 * 8 < 200 * 1024
    rowptr, colidx, diagpos = h.get_lu_pattern()
    colptr, rowval = h.get_pattern()
    perm = h.get_ordering()
    h.close()
    net = ko.Network(S, rd.id_reacs, rd.id_prods, rd.stoic_reacs, rd.stoic_prods)
    rng = np.random.default_rng(S)
    u = rng.uniform(0, 1, S)
    k = 10 ** rng.uniform(-3, 3, R)
    J = net.jac_sparse(u, k).toarray()
    hg = 5000.0
    jv = np.array([J[rowval[p], l] for l in range(S) for p in range(colptr[l], colptr[l + 1])])
    Wp = (np.eye(S) * hg - J)[np.ix_(perm, perm)]
    ref = np.zeros(st["padded"])
    for i in range(S):
        for p in range(rowptr[i], rowptr[i + 1]):
            ref[plan["slot_of"][p]] = Wp[i, colidx[p]]
    run_plan(plan, ref)
    lu, invd = run_fronts(fp, plan, jv, hg, st["padded"])
    scale = np.max(np.abs(ref))
    assert not np.any(np.isnan(lu)) and np.max(np.abs(lu - ref)) <= 1e-12 * scale
    L, U = np.zeros((S, S)), np.eye(S)
    for i in range(S):
        for p in range(rowptr[i], rowptr[i + 1]):
            j, v = colidx[p], lu[plan["slot_of"][p]]
            if j <= i:
                L[i, j] = v
            else:
                U[i, j] = v
    assert np.max(np.abs(L @ U - Wp)) <= 1e-12 * scale
    b = rng.normal(size=S)
    x = sl.solve_triangular(U, sl.solve_triangular(L, b, lower=True), lower=False)
    x_ref = np.linalg.solve(Wp, b)
    assert np.max(np.abs(x - x_ref)) <= 1e-10 * np.max(np.abs(x_ref))
