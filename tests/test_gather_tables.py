"""Host-side check of the gather tables of the right-hand side and the Jacobian
(kb2_symbolic.cpp: work orders, first-touch layouts, sliced ELLs).  A numpy interpreter walks the
tables exactly the way the CUDA warp does (tile_rhs / tile_jac_entries / ell_gather in
kinetica.jl_b200/csrc/kb2_kernels.cuh) on one random member; the result must be the mass-action
right-hand side and the analytic Jacobian of the oracle (reference semantics: SURVEY.md section 8a
R5, src/solving/solve_utils.jl:318-349).  Needs no GPU: the tables come from a host-only handle."""
import numpy as np
import pytest

from kinetica_b200 import _lib
from kinetica_b200.synthetic import synthetic_crn, SEED_BASE
from oracle import kinetica_oracle as ko


def unpack(v):
    v = np.asarray(v, dtype=np.int64)
    coef = ((v >> 24) & 0xff).astype(np.int64)
    coef = np.where(coef >= 128, coef - 256, coef)
    return coef, v & 0xffffff


def ell_gather(order, n, nlong, ell_ptr, ell, G, src, put):
    """ell_gather of kb2_kernels.cuh: item order[nlong + g*G + slot], terms in ascending step."""
    ng = len(ell_ptr) - 1
    assert ng == -(-(n - nlong) // G)
    for g in range(ng):
        base, ln = int(ell_ptr[g]), (int(ell_ptr[g + 1]) - int(ell_ptr[g])) // G
        assert (int(ell_ptr[g + 1]) - base) % G == 0 and base % G == 0      # 16-byte aligned lane slices
        for slot in range(G):
            z = nlong + g * G + slot
            acc = 0.0
            for t in range(ln):
                c, ix = unpack(ell[base + t * G + slot])
                acc += float(c) * src[int(ix)]
            if z < n:
                put(int(order[z]), acc)


def handmade():
    """Edge cases the synthetic generator never produces: three distinct reactants (third slot of the
    derivative table), a collision partner that nets out, a stoichiometry above 3 (general power
    path), a species that takes part in nothing."""
    import kinetica_b200 as kb
    #        A+B+C -> D      A+M -> B+M    4A -> E     D -> A+B     2E -> 3A + C
    reacs = [[0, 1, 2],      [0, 6],       [0],        [3],         [4]]
    prods = [[3],            [1, 6],       [4],        [0, 1],      [0, 2]]
    sr    = [[1, 1, 1],      [1, 1],       [4],        [1],         [2]]
    sp    = [[1],            [1, 1],       [1],        [1, 1],      [3, 1]]
    return kb.RxData(reacs, prods, sr, sp)           # species 5 is inert, species 6 = M


def walk(S, R, seed):
    if seed is None:
        rd = handmade()
    else:
        sd, rd, Ea, A = synthetic_crn(S, R, seed)
    h = _lib.Handle(-1)
    h.set_network(S, *rd.flatten())
    h.symbolic(4)
    return rd, h, h.get_gather_tables()


@pytest.mark.parametrize("S,R,seed", [(64, 256, SEED_BASE + 100), (300, 1500, SEED_BASE + 3), (40, 90, SEED_BASE + 7),
                                      (200, 3000, SEED_BASE + 9), (7, 5, None)])
def test_rhs_tables_reproduce_mass_action(S, R, seed):
    rd, h, T = walk(S, R, seed)
    net = ko.Network(S, rd.id_reacs, rd.id_prods, rd.stoic_reacs, rd.stoic_prods)
    rng = np.random.default_rng(1)
    u = rng.uniform(0.01, 1.0, S); k = 10 ** rng.uniform(-2, 3, R)
    rate = k.copy()
    for j in range(R):
        for i, nu in zip(rd.id_reacs[j], rd.stoic_reacs[j]):
            rate[j] *= u[i] ** nu
    ref = net.rhs(u, k)
    # pass 1: the rate table in first-touch order
    pos = T["rate_pos"]
    assert sorted(pos.tolist()) == list(range(R))                  # a permutation
    table = np.empty(R); table[pos] = rate
    # the CSR by species is exactly the net stoichiometry, ascending reactions
    ptr, rxn, coef, order, nlong, G = T["rhs_ptr"], T["rhs_rxn"], T["rhs_coef"], T["rhs_order"], T["rhs_nlong"], T["ell_g"]
    assert sorted(order.tolist()) == list(range(S))
    lens = np.diff(ptr)[order]
    assert np.all(lens[:-1] >= lens[1:]) and np.all(lens[:nlong] > 64) and (nlong == S or lens[nlong] <= 64)
    out = np.full(S, np.nan)
    for z in range(nlong):                                          # hub rows: lanes stride over the terms
        i = order[z]
        e = np.arange(ptr[i], ptr[i + 1])
        assert np.all(np.diff(rxn[e]) > 0)
        out[i] = np.sum(coef[e] * table[pos[rxn[e]]])
    def put(i, a):
        assert np.isnan(out[i]); out[i] = a
    ell_gather(order, S, nlong, T["ell_ptr"], T["ell"], G, table, put)
    assert not np.any(np.isnan(out))
    np.testing.assert_allclose(out, ref, rtol=1e-12, atol=1e-12 * np.max(np.abs(ref)))
    # first-touch layout: in the device's traversal order (hub rows, then group by group, step by
    # step, slot by slot) every position that has not been seen before is the next unused one
    seen = 0
    def touch(p):
        nonlocal seen
        assert p <= seen
        seen = max(seen, p + 1)
    for z in range(nlong):
        for e in range(ptr[order[z]], ptr[order[z] + 1]):
            touch(int(pos[rxn[e]]))
    for z0 in range(nlong, S, G):
        rows = order[z0:z0 + G]
        for t in range(int(max(ptr[r + 1] - ptr[r] for r in rows))):
            for r in rows:
                if t < ptr[r + 1] - ptr[r]:
                    touch(int(pos[rxn[ptr[r] + t]]))


@pytest.mark.parametrize("S,R,seed", [(64, 256, SEED_BASE + 100), (300, 1500, SEED_BASE + 3), (200, 3000, SEED_BASE + 9),
                                      (7, 5, None)])
def test_jacobian_tables_reproduce_analytic_jacobian(S, R, seed):
    rd, h, T = walk(S, R, seed)
    net = ko.Network(S, rd.id_reacs, rd.id_prods, rd.stoic_reacs, rd.stoic_prods)
    colptr, rowval = h.get_pattern()
    rng = np.random.default_rng(2)
    u = rng.uniform(0.01, 1.0, S); k = 10 ** rng.uniform(-2, 3, R)
    Jref = np.asarray(net.jac_dense(u, k))
    if Jref is None:
        # finite structure: J[i,l] = sum_j net[i,j] k_j nu_lj u_l^(nu_lj - 1) prod_{m != l} u_m^nu_mj
        Jref = np.zeros((S, S))
        for j in range(R):
            sub = {}
            for i, nu in zip(rd.id_reacs[j], rd.stoic_reacs[j]):
                sub[i] = sub.get(i, 0) + nu
            netc = {}
            for i, nu in zip(rd.id_reacs[j], rd.stoic_reacs[j]):
                netc[i] = netc.get(i, 0) - nu
            for i, nu in zip(rd.id_prods[j], rd.stoic_prods[j]):
                netc[i] = netc.get(i, 0) + nu
            for l, nl in sub.items():
                d = k[j] * nl * u[l] ** (nl - 1)
                for mm, nm in sub.items():
                    if mm != l:
                        d *= u[mm] ** nm
                for i, c in netc.items():
                    if c:
                        Jref[i, l] += c * d
    ns, nnzJ = T["jslots"], len(rowval)
    # pass 1: derivative table d[j][s] = k_j * d(prod)/du_(slot s) / nu_s, slots = distinct reactants in
    # ascending species order (rdesc), stored in first-touch order
    dpos = T["drate_pos"]
    assert sorted(dpos.tolist()) == list(range(R * ns))
    dtab = np.zeros(R * ns)
    for j in range(R):
        sub = {}
        for i, nu in zip(rd.id_reacs[j], rd.stoic_reacs[j]):
            sub[i] = sub.get(i, 0) + nu
        assert len(sub) <= ns
        for s, (l, nl) in enumerate(sorted(sub.items())):
            d = k[j] * u[l] ** (nl - 1)
            for mm, nm in sub.items():
                if mm != l:
                    d *= u[mm] ** nm
            dtab[dpos[j * ns + s]] = d
    ptr, order, nlong, G = T["jt_ptr"], T["j_order"], T["j_nlong"], T["ell_g"]
    assert len(ptr) == nnzJ + 1 and sorted(order.tolist()) == list(range(nnzJ))
    # packed terms agree with the (reaction, coefficient*4 + slot) pairs
    c_pk, ix_pk = unpack(T["jt_pk"][:ptr[-1]])
    assert np.array_equal(c_pk, T["jt_pack"] >> 2)
    assert np.array_equal(ix_pk, dpos[T["jt_rxn"].astype(np.int64) * ns + (T["jt_pack"] & 3)])
    Jval = np.full(nnzJ, np.nan)
    for z in range(nlong):
        p = order[z]
        t = np.arange(ptr[p], ptr[p + 1])
        Jval[p] = np.sum(c_pk[t] * dtab[ix_pk[t]])
    def put(p, a):
        assert np.isnan(Jval[p]); Jval[p] = a
    ell_gather(order, nnzJ, nlong, T["jell_ptr"], T["jell"], G, dtab, put)
    assert not np.any(np.isnan(Jval))
    J = np.zeros((S, S))
    for l in range(S):
        for q in range(colptr[l], colptr[l + 1]):
            J[rowval[q], l] = Jval[q]
    np.testing.assert_allclose(J, Jref, rtol=1e-12, atol=1e-12 * np.max(np.abs(Jref)))
