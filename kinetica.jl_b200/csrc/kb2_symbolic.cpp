// Host-side symbolic analysis, computed once per network and shared by every ensemble member.
//
// Replaces, for the B200 path, what the reference obtains from upstream packages when it builds
// `ODEProblem(osys, u0map, tspan, pmap; jac=true, sparse=true)` (reference
// src/solving/methods.jl:157-158, 686-687): the analytic-Jacobian sparsity pattern, plus the
// symbolic phase of the sparse LU the chosen solver would run (docs use KLU,
// docs/src/getting-started.md:66-70).  Mass-action semantics follow `make_rs`
// (src/solving/solve_utils.jl:318-334) with Catalyst's non-combinatoric rate law.
#include "kb2_internal.h"

#include <algorithm>
#include <set>
#include <utility>

namespace kb2 {

std::string build_network(Network &net)
{
    const int64_t S = net.S, R = net.R;
    if (S <= 0) return "network has no species";
    if ((int64_t)net.rp.size() != R + 1 || (int64_t)net.pp.size() != R + 1) return "bad CSR pointer length";
    net.sub_ptr.assign(1, 0); net.sub_idx.clear(); net.sub_exp.clear();
    net.net_ptr.assign(1, 0); net.net_idx.clear(); net.net_coef.clear();
    for (int64_t j = 0; j < R; ++j) {
        std::vector<std::pair<int64_t, int64_t>> sub, nt;
        auto add = [](std::vector<std::pair<int64_t, int64_t>> &v, int64_t s, int64_t n) {
            for (auto &e : v) if (e.first == s) { e.second += n; return; }
            v.emplace_back(s, n);
        };
        for (int64_t e = net.rp[j]; e < net.rp[j + 1]; ++e) {
            if (net.ri[e] < 0 || net.ri[e] >= S) return "reactant species id out of range";
            if (net.rn[e] <= 0) return "reactant stoichiometry must be positive";
            add(sub, net.ri[e], net.rn[e]);
            add(nt, net.ri[e], -net.rn[e]);
        }
        for (int64_t e = net.pp[j]; e < net.pp[j + 1]; ++e) {
            if (net.pi[e] < 0 || net.pi[e] >= S) return "product species id out of range";
            if (net.pn[e] <= 0) return "product stoichiometry must be positive";
            add(nt, net.pi[e], net.pn[e]);
        }
        std::sort(sub.begin(), sub.end());
        std::sort(nt.begin(), nt.end());
        if (sub.size() > 3) return "more than 3 distinct reactant species in one reaction is not supported";
        for (auto &e : sub) {
            if (e.second > 255) return "reactant stoichiometry above 255 is not supported";
            net.sub_idx.push_back((int32_t)e.first);
            net.sub_exp.push_back((int32_t)e.second);
        }
        for (auto &e : nt)
            if (e.second != 0) {
                net.net_idx.push_back((int32_t)e.first);
                net.net_coef.push_back((int32_t)e.second);
            }
        net.sub_ptr.push_back((int32_t)net.sub_idx.size());
        net.net_ptr.push_back((int32_t)net.net_idx.size());
    }
    return "";
}

// Minimum degree on the symmetrised pattern with an explicit elimination graph.  Selection rule
// (the contract the oracle mirrors): alive vertex with the smallest (current degree, index).
static void min_degree(int64_t S, const std::vector<int64_t> &colptr, const std::vector<int64_t> &rowval,
                       std::vector<int64_t> &perm)
{
    std::vector<std::vector<int32_t>> adj(S);
    for (int64_t l = 0; l < S; ++l)
        for (int64_t p = colptr[l]; p < colptr[l + 1]; ++p) {
            int64_t i = rowval[p];
            if (i != l) { adj[i].push_back((int32_t)l); adj[l].push_back((int32_t)i); }
        }
    for (auto &a : adj) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); }
    std::set<std::pair<int32_t, int32_t>> heap;
    for (int64_t v = 0; v < S; ++v) heap.emplace((int32_t)adj[v].size(), (int32_t)v);
    perm.clear();
    std::vector<int32_t> merged;
    while (!heap.empty()) {
        auto it = heap.begin();
        int32_t v = it->second;
        heap.erase(it);
        perm.push_back(v);
        std::vector<int32_t> nb;
        nb.swap(adj[v]);
        for (int32_t a : nb) {
            heap.erase({(int32_t)adj[a].size(), a});
            merged.clear();
            // adj[a] = (adj[a] U nb) \ {a, v}
            size_t x = 0, y = 0;
            const auto &A = adj[a];
            while (x < A.size() || y < nb.size()) {
                int32_t c;
                if (y >= nb.size() || (x < A.size() && A[x] < nb[y])) c = A[x++];
                else if (x >= A.size() || nb[y] < A[x]) c = nb[y++];
                else { c = A[x]; ++x; ++y; }
                if (c != a && c != v) merged.push_back(c);
            }
            adj[a] = merged;
            heap.emplace((int32_t)adj[a].size(), a);
        }
    }
}

// Orderings for networks with locality ("banded" networks: species react with species near them,
// plus a few hub species that react with everything).  Hub species — symmetrised degree above
// max(32, 8 * median) — go last, by ascending degree; the others are ordered on the graph without
// the hubs by
//   3  their natural order,
//   5  reverse Cuthill-McKee (George-Liu pseudo-peripheral start per component),
//   6  Sloan's profile reduction with weights (W1, W2) = (1, 2),
//   7  Sloan's with (2, 1).
// The panel factorisation pads least, and the window of the right-looking LU stays smallest, when
// the envelope of the permuted pattern is small and smooth: natural order is as good as the
// generator's locality, RCM minimises the bandwidth, Sloan's the profile (= the fill).
// All choices are deterministic: ties go to the smaller (degree, index).
static void banded_order(int64_t S, const std::vector<int64_t> &colptr, const std::vector<int64_t> &rowval, int kind,
                         std::vector<int64_t> &perm)
{
    std::vector<std::vector<int32_t>> adj(S);
    for (int64_t l = 0; l < S; ++l)
        for (int64_t p = colptr[l]; p < colptr[l + 1]; ++p)
            if (rowval[p] != l) { adj[l].push_back((int32_t)rowval[p]); adj[rowval[p]].push_back((int32_t)l); }
    std::vector<int64_t> deg(S);
    for (int64_t v = 0; v < S; ++v) {
        std::sort(adj[v].begin(), adj[v].end());
        adj[v].erase(std::unique(adj[v].begin(), adj[v].end()), adj[v].end());
        deg[v] = (int64_t)adj[v].size();
    }
    std::vector<int64_t> srt(deg);
    std::sort(srt.begin(), srt.end());
    const int64_t thr = std::max<int64_t>(32, 8 * srt[S / 2]);
    std::vector<char> hub(S, 0);
    std::vector<std::pair<int64_t, int64_t>> dense;
    for (int64_t v = 0; v < S; ++v) if (deg[v] > thr) { hub[v] = 1; dense.emplace_back(deg[v], v); }
    std::sort(dense.begin(), dense.end());
    perm.clear();
    if (kind == 3) {
        for (int64_t v = 0; v < S; ++v) if (!hub[v]) perm.push_back(v);
    } else {
        // the graph without the hubs; neighbours sorted by (degree in that graph, index)
        std::vector<int32_t> sdeg(S, 0);
        for (int64_t v = 0; v < S; ++v) {
            if (hub[v]) { adj[v].clear(); continue; }
            adj[v].erase(std::remove_if(adj[v].begin(), adj[v].end(), [&](int32_t w) { return hub[w] != 0; }), adj[v].end());
            sdeg[v] = (int32_t)adj[v].size();
        }
        auto by_deg = [&](int32_t a, int32_t b) { return sdeg[a] != sdeg[b] ? sdeg[a] < sdeg[b] : a < b; };
        for (int64_t v = 0; v < S; ++v) std::sort(adj[v].begin(), adj[v].end(), by_deg);
        std::vector<char> done(S, 0);
        std::vector<int32_t> lev(S, -1), order, touched;
        // breadth-first levels from `start` over the vertices that are not done; neighbours in
        // (degree, index) order: this IS the Cuthill-McKee numbering of the component
        auto bfs = [&](int32_t start) {
            for (int32_t v : touched) lev[v] = -1;
            touched.clear(); order.clear();
            lev[start] = 0; touched.push_back(start); order.push_back(start);
            for (size_t q = 0; q < order.size(); ++q) {
                const int32_t v = order[q];
                for (int32_t w : adj[v])
                    if (lev[w] < 0 && !done[w]) { lev[w] = lev[v] + 1; touched.push_back(w); order.push_back(w); }
            }
            return lev[order.back()];      // eccentricity of start
        };
        // vertex of the last level with the smallest (degree, index)
        auto last_level_min = [&](int ecc) {
            int32_t best = -1;
            for (int32_t v : order) if (lev[v] == ecc && (best < 0 || by_deg(v, best))) best = v;
            return best;
        };
        auto pseudo_peripheral = [&](int32_t start) {
            int ecc = bfs(start);
            for (;;) {
                const int32_t cand = last_level_min(ecc);
                const int e2 = bfs(cand);
                if (e2 > ecc) { start = cand; ecc = e2; }
                else { bfs(start); return start; }
            }
        };
        std::vector<int32_t> roots;
        for (int64_t v = 0; v < S; ++v) if (!hub[v]) roots.push_back((int32_t)v);
        std::sort(roots.begin(), roots.end(), by_deg);
        std::vector<int64_t> seq;
        const int W1 = kind == 7 ? 2 : 1, W2 = kind == 7 ? 1 : 2;
        std::vector<int32_t> status(S, 0), dist(S, 0);      // Sloan: 0 inactive, 1 preactive, 2 active, 3 numbered
        std::vector<int64_t> prio(S, 0);
        for (int32_t v0 : roots) {
            if (done[v0]) continue;
            const int32_t s = pseudo_peripheral(v0);       // leaves the levels / order of the search from s behind
            if (kind == 5) {
                for (int32_t v : order) { done[v] = 1; seq.push_back(v); }
                continue;
            }
            // Sloan: number from s towards the far end e; priority = W2 * distance to e - W1 * (degree + 1),
            // raised by W1 whenever a neighbour is numbered or becomes active (the vertex would add
            // fewer new vertices to the front).  Highest priority first, ties to the smaller index.
            const int ecc = lev[order.back()];
            const int32_t e = last_level_min(ecc);
            std::vector<int32_t> comp(order);
            bfs(e);
            for (int32_t v : comp) { dist[v] = lev[v]; status[v] = 0; prio[v] = (int64_t)W2 * dist[v] - (int64_t)W1 * (sdeg[v] + 1); }
            std::set<std::pair<int64_t, int32_t>> heap;      // (-priority, vertex)
            auto raise = [&](int32_t w) {
                if (status[w] == 1 || status[w] == 2) heap.erase({-prio[w], w});
                prio[w] += W1;
                if (status[w] == 0) status[w] = 1;
                heap.emplace(-prio[w], w);
            };
            status[s] = 1;
            heap.emplace(-prio[s], s);
            while (!heap.empty()) {
                const int32_t v = heap.begin()->second;
                heap.erase(heap.begin());
                if (status[v] == 1)
                    for (int32_t w : adj[v]) if (!done[w] && status[w] != 3) raise(w);
                status[v] = 3; done[v] = 1; seq.push_back(v);
                for (int32_t w : adj[v])
                    if (status[w] == 1 && !done[w]) {
                        heap.erase({-prio[w], w});
                        status[w] = 2; prio[w] += W1;
                        heap.emplace(-prio[w], w);
                        for (int32_t x : adj[w]) if (!done[x] && status[x] != 3) raise(x);
                    }
            }
        }
        if (kind == 5) std::reverse(seq.begin(), seq.end());
        perm = seq;
    }
    for (auto &d : dense) perm.push_back(d.second);
}

std::string build_symbolic(const Network &net, int ordering, Symbolic &sym, int64_t fma_limit, int64_t env_limit)
{
    const int64_t S = net.S, R = net.R;
    // ---- Jacobian pattern P_J = {(i,l): exists j, net[i,j] != 0 and nu_lj > 0}, CSC ----
    {
        std::vector<std::vector<int32_t>> cols(S);
        for (int64_t j = 0; j < R; ++j)
            for (int32_t a = net.sub_ptr[j]; a < net.sub_ptr[j + 1]; ++a)
                for (int32_t e = net.net_ptr[j]; e < net.net_ptr[j + 1]; ++e)
                    cols[net.sub_idx[a]].push_back(net.net_idx[e]);
        sym.colptr.assign(S + 1, 0);
        sym.rowval.clear();
        for (int64_t l = 0; l < S; ++l) {
            auto &c = cols[l];
            std::sort(c.begin(), c.end());
            c.erase(std::unique(c.begin(), c.end()), c.end());
            for (int32_t i : c) sym.rowval.push_back(i);
            sym.colptr[l + 1] = (int64_t)sym.rowval.size();
        }
        sym.nnzJ = (int64_t)sym.rowval.size();
    }
    // ---- ordering ----
    if (ordering == 0) min_degree(S, sym.colptr, sym.rowval, sym.perm);
    else if (ordering == 1) { sym.perm.resize(S); for (int64_t a = 0; a < S; ++a) sym.perm[a] = a; }
    else if (ordering == 3 || ordering == 5 || ordering == 6 || ordering == 7)
        banded_order(S, sym.colptr, sym.rowval, ordering, sym.perm);
    else {
        if ((int64_t)sym.perm.size() != S) return "ordering 2 requested but kb2_set_ordering was not called";
        std::vector<char> seen(S, 0);
        for (int64_t a = 0; a < S; ++a) {
            if (sym.perm[a] < 0 || sym.perm[a] >= S || seen[sym.perm[a]]) return "supplied ordering is not a permutation";
            seen[sym.perm[a]] = 1;
        }
    }
    sym.iperm.assign(S, 0);
    for (int64_t a = 0; a < S; ++a) sym.iperm[sym.perm[a]] = a;
    if (env_limit != INT64_MAX) {
        // envelope of the permuted symmetrised pattern: sum over rows of (row - first column) on both sides
        std::vector<int64_t> first(S);
        for (int64_t a = 0; a < S; ++a) first[a] = a;
        for (int64_t l = 0; l < S; ++l)
            for (int64_t p = sym.colptr[l]; p < sym.colptr[l + 1]; ++p) {
                const int64_t a = sym.iperm[sym.rowval[p]], b = sym.iperm[l];
                first[std::max(a, b)] = std::min(first[std::max(a, b)], std::min(a, b));
            }
        int64_t env = 0;
        for (int64_t a = 0; a < S; ++a) env += 2 * (a - first[a]);
        if (env > env_limit) return "profile beyond the limit";
    }
    // ---- row-wise symbolic LU of P (P_J U diag) P^T without pivoting ----
    {
        std::vector<std::vector<int32_t>> rows(S);
        for (int64_t a = 0; a < S; ++a) rows[a].push_back((int32_t)a);
        for (int64_t l = 0; l < S; ++l)
            for (int64_t p = sym.colptr[l]; p < sym.colptr[l + 1]; ++p)
                rows[sym.iperm[sym.rowval[p]]].push_back((int32_t)sym.iperm[l]);
        std::vector<std::vector<int32_t>> upper(S);
        std::vector<int32_t> mark(S, -1);
        sym.rowptr.assign(S + 1, 0);
        sym.colidx.clear();
        sym.diagpos.assign(S, 0);
        sym.n_fma = 0;
        std::vector<int32_t> pat;
        for (int64_t i = 0; i < S; ++i) {
            pat.clear();
            std::set<int32_t> lower;           // pending pivots < i, ascending
            for (int32_t c : rows[i])
                if (mark[c] != (int32_t)i) { mark[c] = (int32_t)i; pat.push_back(c); if (c < i) lower.insert(c); }
            while (!lower.empty()) {
                int32_t k = *lower.begin();
                lower.erase(lower.begin());
                sym.n_fma += (int64_t)upper[k].size();
                if (sym.n_fma > fma_limit) return "fill beyond the limit";
                for (int32_t j : upper[k])
                    if (mark[j] != (int32_t)i) { mark[j] = (int32_t)i; pat.push_back(j); if (j < i) lower.insert(j); }
            }
            std::sort(pat.begin(), pat.end());
            for (size_t q = 0; q < pat.size(); ++q) {
                if (pat[q] == (int32_t)i) sym.diagpos[i] = sym.rowptr[i] + (int64_t)q;
                if (pat[q] > (int32_t)i) upper[i].push_back(pat[q]);
                sym.colidx.push_back(pat[q]);
            }
            sym.rowptr[i + 1] = (int64_t)sym.colidx.size();
        }
        sym.nnzLU = (int64_t)sym.colidx.size();
    }
    if (sym.nnzLU >= (int64_t)1 << 31 || sym.n_fma >= (int64_t)1 << 32) return "factorisation too large for 32-bit tables";
    // ---- RHS gather CSR (rows = species, entries in ascending reaction order) ----
    {
        std::vector<int32_t> cnt(S + 1, 0);
        for (size_t e = 0; e < net.net_idx.size(); ++e) cnt[net.net_idx[e] + 1]++;
        sym.rhs_ptr.assign(S + 1, 0);
        for (int64_t i = 0; i < S; ++i) sym.rhs_ptr[i + 1] = sym.rhs_ptr[i] + cnt[i + 1];
        sym.rhs_rxn.assign(net.net_idx.size(), 0);
        sym.rhs_coef.assign(net.net_idx.size(), 0);
        std::vector<int32_t> fill(sym.rhs_ptr.begin(), sym.rhs_ptr.end() - 1);
        for (int64_t j = 0; j < R; ++j)
            for (int32_t e = net.net_ptr[j]; e < net.net_ptr[j + 1]; ++e) {
                int32_t pos = fill[net.net_idx[e]]++;
                sym.rhs_rxn[pos] = (int32_t)j;
                sym.rhs_coef[pos] = net.net_coef[e];
            }
    }
    // ---- reaction descriptors ----
    sym.rdesc.assign(4 * (size_t)std::max<int64_t>(R, 1), -1);
    for (int64_t j = 0; j < R; ++j) {
        int32_t ex = 0;
        int s = 0;
        for (int32_t a = net.sub_ptr[j]; a < net.sub_ptr[j + 1]; ++a, ++s) {
            sym.rdesc[4 * j + s] = net.sub_idx[a];
            ex |= net.sub_exp[a] << (8 * s);
        }
        sym.rdesc[4 * j + 3] = ex;
    }
    // ---- Jacobian terms per CSC entry ----
    {
        std::vector<std::vector<std::pair<int32_t, int32_t>>> terms(sym.nnzJ);
        for (int64_t j = 0; j < R; ++j) {
            int s = 0;
            for (int32_t a = net.sub_ptr[j]; a < net.sub_ptr[j + 1]; ++a, ++s) {
                int64_t l = net.sub_idx[a];
                for (int32_t e = net.net_ptr[j]; e < net.net_ptr[j + 1]; ++e) {
                    int64_t i = net.net_idx[e];
                    auto b = sym.rowval.begin() + sym.colptr[l], en = sym.rowval.begin() + sym.colptr[l + 1];
                    int64_t p = std::lower_bound(b, en, i) - sym.rowval.begin();
                    int32_t c = net.net_coef[e] * net.sub_exp[a];
                    terms[p].emplace_back((int32_t)j, (int32_t)(c * 4 + s));
                }
            }
        }
        sym.jt_ptr.assign(sym.nnzJ + 1, 0);
        sym.jt_rxn.clear(); sym.jt_pack.clear();
        for (int64_t p = 0; p < sym.nnzJ; ++p) {
            for (auto &t : terms[p]) { sym.jt_rxn.push_back(t.first); sym.jt_pack.push_back(t.second); }
            sym.jt_ptr[p + 1] = (int32_t)sym.jt_rxn.size();
        }
    }
    // ---- LU slot tables ----
    sym.lu_rowptr.assign(sym.rowptr.begin(), sym.rowptr.end());
    sym.lu_colidx.assign(sym.colidx.begin(), sym.colidx.end());
    sym.lu_diagpos.assign(sym.diagpos.begin(), sym.diagpos.end());
    sym.slot_src.assign(sym.nnzLU, 0);
    sym.max_rowlen = 0;
    for (int64_t a = 0; a < S; ++a) {
        sym.max_rowlen = std::max<int32_t>(sym.max_rowlen, (int32_t)(sym.rowptr[a + 1] - sym.rowptr[a]));
        int64_t i = sym.perm[a];
        for (int64_t q = sym.rowptr[a]; q < sym.rowptr[a + 1]; ++q) {
            int64_t l = sym.perm[sym.colidx[q]];
            auto b = sym.rowval.begin() + sym.colptr[l], en = sym.rowval.begin() + sym.colptr[l + 1];
            auto it = std::lower_bound(b, en, i);
            int32_t p1 = (it != en && *it == i) ? (int32_t)(it - sym.rowval.begin()) + 1 : 0;
            sym.slot_src[q] = (p1 << 1) | (sym.colidx[q] == a ? 1 : 0);
        }
    }
    // ---- work orders of the gather loops (load balance: hub species have rows of hundreds of
    // terms, which one lane must not walk alone) ----
    {
        auto order_by_len = [](const std::vector<int32_t> &ptr, int64_t n, int thr, std::vector<int32_t> &ord, int32_t &nlong) {
            ord.resize(n);
            for (int64_t i = 0; i < n; ++i) ord[i] = (int32_t)i;
            std::stable_sort(ord.begin(), ord.end(), [&](int32_t a, int32_t b) { return ptr[a + 1] - ptr[a] > ptr[b + 1] - ptr[b]; });
            nlong = 0;
            while (nlong < n && ptr[ord[nlong] + 1] - ptr[ord[nlong]] > thr) ++nlong;
        };
        order_by_len(sym.rhs_ptr, S, Symbolic::RHS_LONG, sym.rhs_order, sym.rhs_nlong);
        order_by_len(sym.jt_ptr, sym.nnzJ, Symbolic::JAC_LONG, sym.j_order, sym.j_nlong);
    }
    // ---- sliced ELL of the gather rows that go one per lane (RHS rows, Jacobian entries) ----
    // The items after the first `nlong` of `order` are cut into groups of ELL_G consecutive items
    // (about equally long: the order is by decreasing length).  Group g stores, for every step
    // t < len_g and every slot rho < ELL_G, one int  coef << 24 | index  at
    // ell[ell_ptr[g] + t*ELL_G + rho]; an item that has run out repeats its last index with
    // coefficient 0 (index 0 for an empty item).  A lane owns ELL_G/LN consecutive slots, so the
    // indices of one step are one or two 16-byte loads that do not depend on any data.
    auto build_ell = [&](const std::vector<int32_t> &ptr, const std::vector<int32_t> &order, int64_t n, int32_t nlong,
                         const std::vector<int32_t> &idx, const std::vector<int32_t> &coef,
                         std::vector<int32_t> &ell_ptr, std::vector<int32_t> &ell) -> std::string {
        const int G = Symbolic::ELL_G;
        const int64_t nrows = n - nlong;
        const int64_t ng = (nrows + G - 1) / G;
        ell_ptr.assign(ng + 1, 0);
        ell.clear();
        for (int64_t g = 0; g < ng; ++g) {
            int32_t len = 0;
            for (int rho = 0; rho < G; ++rho) {
                const int64_t z = nlong + g * G + rho;
                if (z < n) len = std::max(len, ptr[order[z] + 1] - ptr[order[z]]);
            }
            const size_t base = ell.size();
            ell.resize(base + (size_t)len * G, 0);
            for (int rho = 0; rho < G; ++rho) {
                const int64_t z = nlong + g * G + rho;
                if (z >= n) continue;
                const int32_t i = order[z], e0 = ptr[i], cnt = ptr[i + 1] - e0;
                for (int32_t t = 0; t < len; ++t) {
                    int32_t v = 0;
                    if (cnt > 0) {
                        const int32_t e = e0 + std::min(t, cnt - 1);
                        const int32_t c = t < cnt ? coef[e] : 0;
                        if (c < -128 || c > 127 || idx[e] < 0 || idx[e] >= (1 << 24)) return "coefficient or index out of range for the packed gather table";
                        v = (int32_t)(((uint32_t)(c & 0xff) << 24) | (uint32_t)idx[e]);
                    }
                    ell[base + (size_t)t * G + rho] = v;
                }
            }
            ell_ptr[g + 1] = (int32_t)ell.size();
        }
        if (ell.empty()) ell.assign(4, 0);
        return "";
    };
    // First-touch layout of a gathered table: position of every source index in the order in which the
    // device's traversal (long items first, lanes striding over their terms; then group by group,
    // step by step, slot by slot) touches it first.  What the 64 slots of a warp gather in one
    // step is then mostly consecutive sectors of new data instead of 64 random ones, and re-touches
    // go to sectors that were read a few steps earlier.
    auto first_touch = [&](const std::vector<int32_t> &ptr, const std::vector<int32_t> &order, int64_t n, int32_t nlong,
                           const std::vector<int32_t> &idx, int64_t nsrc, std::vector<int32_t> &pos) {
        const int G = Symbolic::ELL_G;
        pos.assign(nsrc, -1);
        int32_t next = 0;
        auto touch = [&](int32_t j) { if (pos[j] < 0) pos[j] = next++; };
        for (int32_t z = 0; z < nlong; ++z)
            for (int32_t e = ptr[order[z]]; e < ptr[order[z] + 1]; ++e) touch(idx[e]);
        for (int64_t z0 = nlong; z0 < n; z0 += G) {
            int32_t len = 0;
            for (int64_t z = z0; z < std::min<int64_t>(z0 + G, n); ++z) len = std::max(len, ptr[order[z] + 1] - ptr[order[z]]);
            for (int32_t t = 0; t < len; ++t)
                for (int64_t z = z0; z < std::min<int64_t>(z0 + G, n); ++z) {
                    const int32_t e0 = ptr[order[z]], cnt = ptr[order[z] + 1] - e0;
                    if (t < cnt) touch(idx[e0 + t]);
                }
        }
        for (int64_t j = 0; j < nsrc; ++j) touch((int32_t)j);
    };
    {
        // the rate table is stored in first-touch order: rate_pos[j] = where reaction j's rate lives
        first_touch(sym.rhs_ptr, sym.rhs_order, S, sym.rhs_nlong, sym.rhs_rxn, R, sym.rate_pos);
        sym.rhs_src.resize(sym.rhs_rxn.size());
        for (size_t e = 0; e < sym.rhs_rxn.size(); ++e) sym.rhs_src[e] = sym.rate_pos[sym.rhs_rxn[e]];
        std::string err = build_ell(sym.rhs_ptr, sym.rhs_order, S, sym.rhs_nlong, sym.rhs_src, sym.rhs_coef, sym.ell_ptr, sym.ell);
        if (!err.empty()) return err;
        if (sym.rate_pos.empty()) sym.rate_pos.assign(1, 0);
    }
    // Jacobian terms address the derivative table  d[j*jslots + s] = d(rate_j)/du_(slot s) / nu_s
    // that the device fills per reaction; coefficient = net coefficient * nu_s
    {
        int32_t ns = 1;
        for (int64_t j = 0; j < R; ++j) ns = std::max(ns, net.sub_ptr[j + 1] - net.sub_ptr[j]);
        if (ns > 3) return "more than three distinct reactants in one reaction";
        sym.jslots = ns;
        const size_t nt = sym.jt_rxn.size();
        sym.jt_idx.resize(nt); sym.jt_coef.resize(nt); sym.jt_pk.resize(std::max<size_t>(nt, 1), 0);
        for (size_t t = 0; t < nt; ++t) {
            sym.jt_idx[t] = sym.jt_rxn[t] * ns + (sym.jt_pack[t] & 3);
            sym.jt_coef[t] = sym.jt_pack[t] >> 2;
            if (sym.jt_coef[t] < -128 || sym.jt_coef[t] > 127 || sym.jt_idx[t] >= (1 << 24)) return "Jacobian term out of range for the packed gather table";
            sym.jt_pk[t] = (int32_t)(((uint32_t)(sym.jt_coef[t] & 0xff) << 24) | (uint32_t)sym.jt_idx[t]);
        }
        // same first-touch layout for the derivative table: drate_pos[j*jslots + s]
        first_touch(sym.jt_ptr, sym.j_order, sym.nnzJ, sym.j_nlong, sym.jt_idx, R * ns, sym.drate_pos);
        for (size_t t = 0; t < nt; ++t) {
            sym.jt_idx[t] = sym.drate_pos[sym.jt_idx[t]];
            sym.jt_pk[t] = (int32_t)(((uint32_t)(sym.jt_coef[t] & 0xff) << 24) | (uint32_t)sym.jt_idx[t]);
        }
        if (sym.drate_pos.empty()) sym.drate_pos.assign(1, 0);
        std::string err = build_ell(sym.jt_ptr, sym.j_order, sym.nnzJ, sym.j_nlong, sym.jt_idx, sym.jt_coef, sym.jell_ptr, sym.jell);
        if (!err.empty()) return err;
    }
    sym.ready = true;
    return "";
}

}  // namespace kb2
