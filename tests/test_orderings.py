"""Host-side tests of the elimination orderings of kb2_symbolic (kb2_symbolic.cpp `banded_order`,
`auto` in kb2_api.cu) against the restatement in oracle/orderings.py and scipy's Cuthill-McKee, and of the
cost model `auto` chooses by.  Needs no GPU (host-only handles)."""
import numpy as np
import pytest
import scipy.sparse as sp
from scipy.sparse.csgraph import reverse_cuthill_mckee

from kinetica_b200 import _lib
from kinetica_b200.synthetic import synthetic_crn, SEED_BASE
from oracle import orderings as oo

NETS = [(1000, 5000, 3), (200, 1000, 250), (420, 2100, 470), (64, 256, 100), (30, 60, 80)]


def _handle(S, R, seed):
    sd, rd, _, _ = synthetic_crn(S, R, SEED_BASE + seed)
    h = _lib.Handle(-1)
    h.set_network(S, *rd.flatten())
    return h


@pytest.mark.parametrize("S,R,seed", NETS)
def test_banded_orderings_match_the_restatement(S, R, seed):
    h = _handle(S, R, seed)
    h.symbolic(1)
    colptr, rowval = h.get_pattern()
    for code, kind, w in ((3, "natural", None), (5, "rcm", None), (6, "sloan", (1, 2)), (7, "sloan", (2, 1))):
        h.symbolic(code)
        perm = h.get_ordering()
        assert sorted(perm.tolist()) == list(range(S))
        ref = oo.banded_order(S, colptr, rowval, kind, w) if w else oo.banded_order(S, colptr, rowval, kind)
        assert np.array_equal(perm, ref), (code, kind)
        assert h.get_plan_stats()["ordering"] == code
    h.close()


def test_rcm_is_scipys_on_the_graph_without_the_hubs():
    S, R, seed = 1000, 5000, 3
    h = _handle(S, R, seed)
    h.symbolic(5)
    perm = h.get_ordering()
    colptr, rowval = h.get_pattern()
    h.close()
    adj, sdeg, hub, hubs = oo.hubless_graph(S, colptr, rowval)
    assert len(hubs) == 8 and sorted(hubs) == list(range(8))      # the generator's hub species
    keep = np.where(~hub)[0]
    pos = -np.ones(S, dtype=int); pos[keep] = np.arange(len(keep))
    rows = [pos[v] for v in keep for _ in adj[v]]
    cols = [pos[w] for v in keep for w in adj[v]]
    G = sp.coo_matrix((np.ones(len(rows)), (rows, cols)), shape=(len(keep), len(keep))).tocsr()
    r = reverse_cuthill_mckee(G, symmetric_mode=True)
    assert np.array_equal(perm[:len(keep)], keep[r]) and perm[len(keep):].tolist() == hubs


def _cost(h, cap=227 * 1024):
    ps, fp = h.get_plan_stats(), h.get_front_plan()
    f = fp["f_info"]
    blocks = float(np.sum(((f[:, 3] + 7) // 8) * ((f[:, 2] + 3) // 4)))
    lu = 66.3 * blocks + 38.9 * float(np.sum(f[:, 2] + f[:, 3])) + 4523.0 * fp["NF"] + 2000.0 * float(np.sum(f[:, 10] == 0))
    lcap = 16 + 2 * fp["max_nu"] + 2 * fp["max_nl"]
    fits = [mw for mw in (4, 2, 1) if 8 * mw * (fp["Wr"] * fp["Wc"] + 144) + 8 * 4 * 256 + 3 * lcap * 4 + 16 <= cap]
    lu *= 4.0 / fits[0] if fits else 6.0
    return lu + 5.25 * ps["padded"]


@pytest.mark.parametrize("S,R,seed", NETS)
def test_auto_keeps_the_candidate_with_the_smallest_modelled_cost(S, R, seed):
    """... unless that is less than 3 % ahead of the natural order with hub species last (the baseline every
    full-size run was first made with; 3 % is within the model's error)."""
    h = _handle(S, R, seed)
    costs = {}
    for code in (3, 5, 6, 7, 0):
        h.symbolic(code)
        costs[code] = _cost(h)
    h.symbolic(4)
    chosen = h.get_plan_stats()["ordering"]
    others = {c: v for c, v in costs.items() if c != 3}
    best = min(others, key=lambda c: (others[c], (0, 5, 6, 7).index(c)))
    assert chosen == (best if others[best] < 0.97 * costs[3] else 3)
    assert abs(_cost(h) - costs[chosen]) < 1e-6 * costs[chosen]
    h.close()


def test_auto_choices_on_the_baseline_networks():
    """C3 and C5 leave the natural order (modelled 12 % / 6 % ahead), C4 keeps it (2 %)."""
    for S, seed, want in ((1000, 3, 6), (5000, 5, 7), (10000, 4, 3)):
        h = _handle(S, 5 * S, seed)
        h.symbolic(4)
        assert h.get_plan_stats()["ordering"] == want, (S, h.get_plan_stats())
        h.close()


def test_bench_network_gets_the_profile_ordering():
    """C3: Sloan's ordering (1:2) — a tenth less padded storage and a fifth fewer padded FMAs than the
    natural order with hubs last, and a window that fits shared memory with four members per CTA."""
    h = _handle(1000, 5000, 3)
    h.symbolic(3)
    nat = h.get_plan_stats()
    h.symbolic(4)
    st, fp = h.get_plan_stats(), h.get_front_plan()
    assert st["ordering"] == 6
    assert st["padded"] < 0.92 * nat["padded"] and st["fma_padded"] < 0.82 * nat["fma_padded"]
    assert 8 * 4 * (fp["Wr"] * fp["Wc"] + 144) < 200 * 1024
    h.close()


def test_random_small_networks_every_ordering():
    """Forty random networks (4 to 70 species, with and without locality or hubs): every ordering is a
    permutation, equals the restatement, and its block and front plans factorise W correctly (numpy
    interpreters of tests/test_block_plan.py and tests/test_front_plan.py)."""
    from oracle import kinetica_oracle as ko
    from test_block_plan import run_plan
    from test_front_plan import run_fronts
    rng = np.random.default_rng(7)
    done = 0
    for trial in range(40):
        S = int(rng.integers(4, 70))
        R = 2 * int(rng.integers(1, 2 * S))
        try:
            sd, rd, _, _ = synthetic_crn(S, R, 1000 + trial, w=float(rng.choice([1.5, 4, 16, 60])),
                                         n_hubs=int(min(S, rng.integers(1, 9))), p_hub=float(rng.choice([0.0, 0.1, 0.5])))
        except ValueError:          # more reactions asked for than the tiny network has
            continue
        net = ko.Network(S, rd.id_reacs, rd.id_prods, rd.stoic_reacs, rd.stoic_prods)
        u = rng.uniform(0, 1, S)
        k = 10 ** rng.uniform(-2, 2, R)
        J = net.jac_dense(u, k)
        for code in (3, 5, 6, 7, 4, 0):
            h = _lib.Handle(-1)
            h.set_network(S, *rd.flatten())
            h.symbolic(code)
            perm = h.get_ordering()
            assert sorted(perm.tolist()) == list(range(S))
            plan, fp, st = h.get_plan(), h.get_front_plan(), h.get_plan_stats()
            rowptr, colidx, _ = h.get_lu_pattern()
            colptr, rowval = h.get_pattern()
            h.close()
            if code in (5, 6, 7):
                ref = oo.banded_order(S, colptr, rowval, "rcm" if code == 5 else "sloan", (1, 2) if code == 6 else (2, 1))
                assert np.array_equal(ref, perm), (trial, code)
            jv = np.array([J[rowval[p], l] for l in range(S) for p in range(colptr[l], colptr[l + 1])])
            Wp = (np.eye(S) * 1e4 - J)[np.ix_(perm, perm)]
            ref = np.zeros(st["padded"])
            for i in range(S):
                for p in range(rowptr[i], rowptr[i + 1]):
                    ref[plan["slot_of"][p]] = Wp[i, colidx[p]]
            run_plan(plan, ref)
            lu, _ = run_fronts(fp, plan, jv, 1e4, st["padded"])
            assert not np.any(np.isnan(lu)) and np.max(np.abs(lu - ref)) <= 1e-11 * np.max(np.abs(ref)), (trial, code)
            done += 1
    assert done >= 150


def test_auto_on_a_network_without_locality_falls_back_to_minimum_degree():
    """Reactions between species drawn from the whole network (window = S / 3): the banded candidates' fill is
    several times the minimum-degree fill; `auto` drops them by their envelope before running their symbolic LU."""
    import time
    S = 1500
    sd, rd, _, _ = synthetic_crn(S, 5 * S, SEED_BASE + 77, w=S / 3.0)
    h = _lib.Handle(-1)
    h.set_network(S, *rd.flatten())
    t = time.perf_counter()
    h.symbolic(0)
    t_md = time.perf_counter() - t
    nnz_md = h.nnzLU
    h.symbolic(3)
    assert h.nnzLU > 2 * nnz_md
    t = time.perf_counter()
    h.symbolic(4)
    t_auto = time.perf_counter() - t
    assert h.get_plan_stats()["ordering"] == 0 and h.nnzLU == nnz_md
    assert t_auto < 3 * t_md + 1.0          # one full analysis plus four cheap rejections
    h.close()
