// Panel plan for the register-blocked sparse LU / panel triangular solves.
//
// Rows of the (permuted) L\U pattern are grouped into panels of up to PR consecutive rows that
// share one column pattern C_P (the union of their patterns; padding entries are explicit zeros
// and stay exactly zero through the elimination).  The numeric factorisation walks the panels in
// order; a panel is processed in column chunks of at most CW columns so that a thread holds its
// PR x (CW/32) targets in registers while the pivots stream past.  Everything here is computed
// once on the host and shared by all ensemble members.
#include "kb2_internal.h"

#include <algorithm>

namespace kb2 {

std::string build_panels(Symbolic &sym, int64_t S)
{
    PanelPlan &pp = sym.panels;
    pp = PanelPlan();
    const int PR = PanelPlan::PR, CW = PanelPlan::CW;
    // ---- panels: greedy runs of consecutive rows; a run stops when the union pattern would
    // outgrow one chunk (unless the row alone is wider than a chunk) ----
    std::vector<int32_t> mark(S, -1), uni;
    pp.row_panel.assign(S, 0);
    pp.row_r.assign(S, 0);
    pp.p_cptr.push_back(0);
    int64_t base = 0;
    for (int64_t p0 = 0; p0 < S;) {
        const int32_t P = (int32_t)pp.p_row0.size();
        uni.clear();
        int nr = 0;
        while (nr < PR && p0 + nr < S) {
            const int64_t i = p0 + nr;
            size_t before = uni.size();
            for (int64_t q = sym.rowptr[i]; q < sym.rowptr[i + 1]; ++q) {
                int32_t c = (int32_t)sym.colidx[q];
                if (mark[c] != P) { mark[c] = P; uni.push_back(c); }
            }
            // rows of the panel are pivots of each other: make sure every diagonal of the run is a column
            if (nr > 0 && (int)uni.size() > CW && (int)before <= CW) {
                for (size_t z = before; z < uni.size(); ++z) mark[uni[z]] = -1;
                uni.resize(before);
                break;
            }
            ++nr;
        }
        std::sort(uni.begin(), uni.end());
        int next = 0;
        while (next < (int)uni.size() && uni[next] < p0) ++next;
        // all diagonals p0..p0+nr-1 are present and contiguous right after the external columns
        for (int r = 0; r < nr; ++r)
            if (next + r >= (int)uni.size() || uni[next + r] != p0 + r) return "internal error: panel diagonal block is not contiguous";
        pp.p_row0.push_back((int32_t)p0);
        pp.p_nrows.push_back(nr);
        pp.p_width.push_back((int32_t)uni.size());
        pp.p_next.push_back(next);
        if (base + (int64_t)nr * (int64_t)uni.size() >= ((int64_t)1 << 31)) return "panel storage exceeds 32-bit slot indices";
        pp.p_base.push_back((int32_t)base);
        base += (int64_t)nr * (int64_t)uni.size();
        for (int32_t c : uni) pp.cols.push_back(c);
        pp.p_cptr.push_back((int32_t)pp.cols.size());
        for (int r = 0; r < nr; ++r) { pp.row_panel[p0 + r] = P; pp.row_r[p0 + r] = r; }
        pp.max_width = std::max(pp.max_width, (int32_t)uni.size());
        p0 += nr;
    }
    pp.padded = base;
    const int32_t NP = (int32_t)pp.p_row0.size();
    // ---- storage slot of every exact-pattern entry; assembly sources over the padded storage ----
    pp.slot_of.assign(sym.nnzLU, 0);
    pp.slot_src.assign(pp.padded, 0);
    pp.diag_slot.assign(S, 0);
    std::vector<int32_t> where(S, -1);
    for (int32_t P = 0; P < NP; ++P) {
        const int32_t *C = pp.cols.data() + pp.p_cptr[P];
        const int W = pp.p_width[P];
        for (int c = 0; c < W; ++c) where[C[c]] = c;
        for (int r = 0; r < pp.p_nrows[P]; ++r) {
            const int64_t i = pp.p_row0[P] + r;
            for (int64_t q = sym.rowptr[i]; q < sym.rowptr[i + 1]; ++q) {
                const int32_t slot = pp.p_base[P] + r * W + where[sym.colidx[q]];
                pp.slot_of[q] = slot;
                pp.slot_src[slot] = sym.slot_src[q];
                if (q == sym.diagpos[i]) pp.diag_slot[i] = slot;
            }
        }
    }
    // ---- units (panel x column chunk), their pivot steps and column maps ----
    // per-panel lookup of column -> position, rebuilt on demand for source panels
    std::vector<int32_t> pos_in_Q(S, -1);
    pp.n_fma_padded = 0;
    for (int32_t P = 0; P < NP; ++P) {
        const int32_t *C = pp.cols.data() + pp.p_cptr[P];
        const int W = pp.p_width[P], next = pp.p_next[P], nr = pp.p_nrows[P], p0 = pp.p_row0[P];
        // chunk boundaries: multiples of CW, never splitting the diagonal block [next, next+nr)
        std::vector<int> cuts{0};
        while (cuts.back() < W) {
            int x0 = cuts.back(), x1 = std::min(W, x0 + CW);
            if (x0 < next && x1 > next && x1 < next + nr) x1 = next;            // block would straddle: cut before it
            if (x0 < next + nr && x0 >= next && x1 < next + nr) return "internal error: diagonal block wider than a chunk";
            cuts.push_back(x1);
        }
        for (size_t ci = 0; ci + 1 < cuts.size(); ++ci) {
            const int x0 = cuts[ci], x1 = cuts[ci + 1];
            PanelPlan::Unit u;
            u.panel = P; u.x0 = x0; u.x1 = x1;
            u.step0 = (int32_t)pp.s_e.size();
            u.map0 = (int32_t)(pp.maps.size() / CW);
            u.diag_here = (next >= x0 && next < x1) ? 1 : 0;
            u.diag_before = (next + nr <= x0) ? 1 : 0;
            // candidate pivots: external columns left of the chunk (PRE) and inside it (INCHUNK)
            std::vector<int32_t> unit_maps;     // source panel ids, in first-use order
            auto map_for = [&](int32_t Q) -> int32_t {
                for (size_t z = 0; z < unit_maps.size(); ++z) if (unit_maps[z] == Q) return (int32_t)z;
                // build the map: chunk column c -> position in C_Q (or -1)
                const int32_t *CQ = pp.cols.data() + pp.p_cptr[Q];
                for (int c = 0; c < pp.p_width[Q]; ++c) pos_in_Q[CQ[c]] = c;
                for (int c = 0; c < CW; ++c) pp.maps.push_back(x0 + c < x1 ? pos_in_Q[C[x0 + c]] : -1);
                for (int c = 0; c < pp.p_width[Q]; ++c) pos_in_Q[CQ[c]] = -1;
                unit_maps.push_back(Q);
                return (int32_t)unit_maps.size() - 1;
            };
            const int ext_end = std::min(next, x1);
            u.n_pre = 0; u.n_ext = 0;
            for (int e = 0; e < ext_end; ++e) {
                const int32_t k = C[e];
                const int32_t Q = pp.row_panel[k];
                const int32_t *CQ = pp.cols.data() + pp.p_cptr[Q];
                const int WQ = pp.p_width[Q];
                // does row k (its panel pattern) reach into this chunk beyond column k?
                int cnt = 0;
                {
                    const int lo = std::max(x0, e + 1);
                    // two-pointer intersection of C[lo..x1) with CQ
                    int a = lo, b2 = (int)(std::upper_bound(CQ, CQ + WQ, k) - CQ);
                    while (a < x1 && b2 < WQ) {
                        if (C[a] == CQ[b2]) { ++cnt; ++a; ++b2; }
                        else if (C[a] < CQ[b2]) ++a; else ++b2;
                    }
                }
                const bool inchunk = e >= x0;
                if (cnt == 0 && !inchunk) continue;      // in-chunk pivots are always finalised (their L value must be stored)
                pp.s_e.push_back(e);
                pp.s_k.push_back(k);
                pp.s_src.push_back(pp.p_base[Q] + pp.row_r[k] * WQ);
                pp.s_map.push_back(cnt ? map_for(Q) : -1);
                if (inchunk) ++u.n_ext; else ++u.n_pre;
                pp.n_fma_padded += (int64_t)cnt * nr;
            }
            // PRE steps must precede in-chunk steps: they already do (e ascending, x0 splits them)
            u.n_maps = (int32_t)unit_maps.size();
            if (u.diag_here || u.diag_before) pp.n_fma_padded += (int64_t)nr * (nr - 1) / 2 * std::max(0, x1 - std::max(x0, next));
            pp.units.push_back(u);
        }
    }
    // ---- per-step staging tables.  The part of the pivot row that a unit needs is one contiguous
    // slot range of the source panel's storage; the CTA copies that range into shared memory
    // (SEG slots per stage) and every lane looks its targets up with a byte index relative to the
    // range start.  A pivot whose targets span more than SEG source slots is split into several
    // steps; the follow-up steps reuse the already published multipliers.
    {
        const int NQ = CW / 32, SEG = PanelPlan::SEG;
        std::vector<int32_t> ne, nk, nsrc, nmap, nflag;
        pp.s_meta.clear(); pp.s_idx.clear();
        for (auto &u : pp.units) {
            const int nsteps = u.n_pre + u.n_ext;
            const int old0 = u.step0;
            u.step0 = (int32_t)ne.size();
            int n_pre_new = 0, n_in_new = 0;
            for (int z = 0; z < nsteps; ++z) {
                const size_t st = (size_t)old0 + z;
                const bool pre = z < u.n_pre;
                // absolute source slot per chunk column (or -1)
                std::vector<int32_t> off(CW, -1);
                if (pp.s_map[st] >= 0) {
                    const int32_t *mp = pp.maps.data() + ((size_t)u.map0 + pp.s_map[st]) * CW;
                    const int ce = pp.s_e[st] - u.x0;
                    for (int cl = 0; cl < CW; ++cl)
                        if (mp[cl] >= 0 && cl > ce) off[cl] = pp.s_src[st] + mp[cl];
                }
                bool first = true;
                for (;;) {
                    int32_t lo = -1;
                    for (int cl = 0; cl < CW; ++cl) if (off[cl] >= 0 && (lo < 0 || off[cl] < lo)) lo = off[cl];
                    if (lo < 0 && !first) break;
                    int32_t hi = lo;
                    std::vector<int8_t> idx(CW, (int8_t)SEG);   // SEG = the stage's constant zero row
                    if (lo >= 0)
                        for (int cl = 0; cl < CW; ++cl)
                            if (off[cl] >= 0 && off[cl] - lo < SEG) { idx[cl] = (int8_t)(off[cl] - lo); hi = std::max(hi, off[cl]); off[cl] = -1; }
                    ne.push_back(pp.s_e[st]); nk.push_back(pp.s_k[st]);
                    // meta: {first source slot, number of slots, pivot column, flags (1 = reuse multipliers)}
                    pp.s_meta.push_back(lo < 0 ? 0 : lo);
                    pp.s_meta.push_back(lo < 0 ? 0 : hi - lo + 1);
                    pp.s_meta.push_back(pp.s_e[st]);
                    pp.s_meta.push_back((first || pre) ? 0 : 1);   // in-chunk follow-ups reuse the published multipliers
                    for (int cs = 0; cs < 32; ++cs) {
                        uint32_t wd = 0;
                        for (int q = 0; q < NQ; ++q) wd |= (uint32_t)(uint8_t)idx[cs * NQ + q] << (8 * q);
                        pp.s_idx.push_back((int32_t)wd);
                    }
                    if (pre) ++n_pre_new; else ++n_in_new;
                    first = false;
                    if (lo < 0) break;
                }
            }
            u.n_pre = n_pre_new; u.n_ext = n_in_new;
        }
        pp.s_e.swap(ne); pp.s_k.swap(nk);
        pp.s_src.clear(); pp.s_map.clear(); pp.maps.clear();
    }
    // ---- blocks: up to NB consecutive steps share one barrier round.  In-chunk steps of one block
    // all belong to the same owner lane (NQ consecutive pivot columns of one thread), which resolves
    // the dependencies among its pivots before publishing their multipliers together.
    {
        const int NB = PanelPlan::NB, NQ = CW / 32;
        pp.b_info.clear(); pp.b_idx.clear();
        for (auto &u : pp.units) {
            u.block0 = (int32_t)(pp.b_info.size() / 4);
            const int n = u.n_pre + u.n_ext;
            int z = 0;
            while (z < n) {
                const bool pre = z < u.n_pre;
                const int lim = pre ? u.n_pre : n;
                int cnt = 1;
                const int owner = pre ? -1 : (pp.s_meta[4 * ((size_t)u.step0 + z) + 2] - u.x0) / NQ;
                // multiplier buffer index of each step: new pivot -> next buffer, follow-up -> same buffer
                int lj = 0;
                std::vector<int> ljs{0};
                while (z + cnt < lim && cnt < NB) {
                    const size_t st = (size_t)u.step0 + z + cnt;
                    const bool follow = (pp.s_meta[4 * st + 3] & 1) != 0;
                    if (!pre && !follow && (pp.s_meta[4 * st + 2] - u.x0) / NQ != owner) break;
                    if (!follow) ++lj;
                    ljs.push_back(lj);
                    ++cnt;
                }
                // a follow-up must stay in the block of its pivot
                while (cnt > 1 && z + cnt < lim && (pp.s_meta[4 * ((size_t)u.step0 + z + cnt) + 3] & 1)) { --cnt; ljs.pop_back(); }
                if (z + cnt < lim && (pp.s_meta[4 * ((size_t)u.step0 + z + cnt) + 3] & 1))
                    return "a pivot row needs more staged segments than one block holds (unsupported pattern)";
                for (int j = 0; j < cnt; ++j) pp.s_meta[4 * ((size_t)u.step0 + z + j) + 3] |= ljs[j] << 8;
                int ljcode = 0;
                for (int j = 0; j < cnt; ++j) ljcode |= ljs[j] << (8 + 2 * j);
                pp.b_info.push_back(z); pp.b_info.push_back(cnt); pp.b_info.push_back((pre ? 0 : 1) | ljcode); pp.b_info.push_back(owner);
                for (int cs = 0; cs < 32; ++cs)
                    for (int j = 0; j < NB; ++j)
                        pp.b_idx.push_back(j < cnt ? pp.s_idx[((size_t)u.step0 + z + j) * 32 + cs] : (int32_t)0x40404040);
                z += cnt;
            }
            u.n_blocks = (int32_t)(pp.b_info.size() / 4) - u.block0;
        }
    }
    pp.ready = true;
    return "";
}

}  // namespace kb2
