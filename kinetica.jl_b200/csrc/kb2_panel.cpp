// Block plan for the supernodal sparse LU / panel triangular solves.
//
// Rows of the (permuted) L\U pattern are grouped into panels of up to PR consecutive rows that
// share one column pattern C_P: the union of their exact patterns, closed so that whenever one
// column of an earlier panel Q appears left of the diagonal, all of Q's columns do.  Padding
// entries are explicit zeros and stay exactly zero through the elimination (every update that
// reaches one has a structurally zero factor).  With complete source blocks the numeric
// factorisation of a panel is a sequence of small dense products
//     W_P[:, targets] -= L_PQ (nr x nq) * U_Q[:, targets]        for Q in the L part of P,
// with L_PQ = W_P[:, Q] * inv(U_QQ) (block Crout: pivots on L, unit diagonal on U).
// A panel wider than CW columns is processed in column chunks (units); chunk boundaries never
// split a source block or the diagonal block.  Everything here is computed once on the host and
// shared by all ensemble members.
#include "kb2_internal.h"

#include <algorithm>

namespace kb2 {

std::string build_panels(Symbolic &sym, int64_t S)
{
    PanelPlan &pp = sym.panels;
    pp = PanelPlan();
    const int PR = PanelPlan::PR, CW = PanelPlan::CW;
    if (S >= 65536) return "block plan packs column positions into 16 bits: S must be < 65536";
    // ---- panels: greedy runs of consecutive rows.  A run stops when the padded block would hold
    // much more than the exact entries of its rows. ----
    std::vector<int32_t> mark(S, -1), uni;
    pp.row_panel.assign(S, 0);
    pp.row_r.assign(S, 0);
    pp.p_cptr.push_back(0);
    int64_t base = 0;
    for (int64_t p0 = 0; p0 < S;) {
        const int32_t P = (int32_t)pp.p_row0.size();
        uni.clear();
        int nr = 0;
        int64_t exact = 0;
        while (nr < PR && p0 + nr < S) {
            const int64_t i = p0 + nr;
            const size_t before = uni.size();
            auto add = [&](int32_t c) { if (mark[c] != P) { mark[c] = P; uni.push_back(c); } };
            for (int64_t q = sym.rowptr[i]; q < sym.rowptr[i + 1]; ++q) {
                const int32_t c = (int32_t)sym.colidx[q];
                if (c < p0) {       // close over the source panel of an external column
                    const int32_t Q = pp.row_panel[c];
                    for (int r = 0; r < pp.p_nrows[Q]; ++r) add(pp.p_row0[Q] + r);
                } else add(c);
            }
            const int64_t ex2 = exact + (sym.rowptr[i + 1] - sym.rowptr[i]);
            if (nr > 0 && (int64_t)(nr + 1) * (int64_t)uni.size() > 2 * ex2 + 64) {
                for (size_t z = before; z < uni.size(); ++z) mark[uni[z]] = -1;
                uni.resize(before);
                break;
            }
            exact = ex2;
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
    // ---- storage slot of every exact-pattern entry (column-major inside a panel: c*nr + r), of
    // every Jacobian entry and of every diagonal ----
    pp.slot_of.assign(sym.nnzLU, 0);
    pp.jslot.assign(sym.nnzJ, -1);
    pp.diag_slot.assign(S, 0);
    std::vector<int32_t> where(S, -1);
    for (int32_t P = 0; P < NP; ++P) {
        const int32_t *C = pp.cols.data() + pp.p_cptr[P];
        const int W = pp.p_width[P], nr = pp.p_nrows[P];
        for (int c = 0; c < W; ++c) where[C[c]] = c;
        for (int r = 0; r < nr; ++r) {
            const int64_t i = pp.p_row0[P] + r;
            for (int64_t q = sym.rowptr[i]; q < sym.rowptr[i + 1]; ++q) {
                const int32_t slot = pp.p_base[P] + where[sym.colidx[q]] * nr + r;
                pp.slot_of[q] = slot;
                const int32_t src = sym.slot_src[q];
                if ((src >> 1) > 0) pp.jslot[(src >> 1) - 1] = slot;
                if (q == sym.diagpos[i]) pp.diag_slot[i] = slot;
            }
        }
    }
    for (int64_t p = 0; p < sym.nnzJ; ++p) if (pp.jslot[p] < 0) return "internal error: Jacobian entry without a storage slot";
    // ---- units, tasks and target maps ----
    std::vector<int32_t> pos_in_P(S, -1);
    pp.n_fma_padded = 0;
    for (int32_t P = 0; P < NP; ++P) {
        const int32_t *C = pp.cols.data() + pp.p_cptr[P];
        const int W = pp.p_width[P], next = pp.p_next[P], nr = pp.p_nrows[P];
        for (int c = 0; c < W; ++c) pos_in_P[C[c]] = c;
        // admissible cut positions: source-block starts in the L part, the diagonal block start,
        // and every column position from the end of the diagonal block on
        std::vector<char> cut_ok(W + 1, 0);
        for (int c = 0; c < next; ++c) if (pp.row_r[C[c]] == 0) cut_ok[c] = 1;
        cut_ok[next] = 1;
        for (int c = next + nr; c <= W; ++c) cut_ok[c] = 1;
        std::vector<int> cuts{0};
        while (cuts.back() < W) {
            const int x0 = cuts.back();
            int x1 = std::min(W, x0 + CW);
            while (x1 > x0 && !cut_ok[x1]) --x1;
            if (x1 == x0) return "internal error: no admissible chunk boundary";
            cuts.push_back(x1);
        }
        for (size_t ci = 0; ci + 1 < cuts.size(); ++ci) {
            const int x0 = cuts[ci], x1 = cuts[ci + 1];
            const int32_t task0 = (int32_t)pp.n_tasks();
            const int dmode = (next >= x0 && next < x1) ? 1 : (next + nr <= x0 ? 2 : 0);
            const int lend = std::min(next, x1);
            for (int e = 0; e < lend;) {
                const int32_t Q = pp.row_panel[C[e]];
                const int nq = pp.p_nrows[Q], WQ = pp.p_width[Q], nextQ = pp.p_next[Q];
                if (pp.row_r[C[e]] != 0) return "internal error: source block is not complete";
                const int32_t *CQ = pp.cols.data() + pp.p_cptr[Q];
                const bool inchunk = e >= x0;
                const int32_t map0 = (int32_t)pp.map.size();
                int ntg = 0, cq_min = 0, cq_max = 0;
                for (int cq = nextQ + nq; cq < WQ; ++cq) {
                    const int ppos = pos_in_P[CQ[cq]];
                    if (ppos >= x0 && ppos < x1) {
                        pp.map.push_back(cq | ((ppos - x0) << 16));
                        if (ntg == 0) cq_min = cq;
                        cq_max = cq;
                        ++ntg;
                    }
                }
                if (inchunk || ntg > 0) {
                    // {Q, lpos | inchunk << 30, targets, first map entry, nq, base of Q, slot of U'_QQ,
                    //  [next in-chunk source: slot of its U'_QQ, its nq], [span of U' columns the targets touch: first slot, slots], 0}
                    const int32_t rec[PanelPlan::TREC] = {Q, e | (inchunk ? (1 << 30) : 0), ntg, map0, nq, pp.p_base[Q],
                                                          pp.p_base[Q] + nextQ * nq, -1, 0,
                                                          pp.p_base[Q] + cq_min * nq, ntg ? (cq_max + 1 - cq_min) * nq : 0, 0};
                    pp.t_info.insert(pp.t_info.end(), rec, rec + PanelPlan::TREC);
                    pp.n_fma_padded += (int64_t)ntg * nr * nq + (inchunk ? (int64_t)nr * nq * (nq - 1) / 2 : 0);
                }
                e += nq;
            }
            if (dmode) pp.n_fma_padded += (int64_t)nr * (nr - 1) / 2 * std::max(0, x1 - std::max(x0, next));
            const int32_t ntask = (int32_t)pp.n_tasks() - task0;
            // link every task (and the unit) to the next in-chunk source, whose U'_QQ is staged ahead
            int32_t nx_off = -1, nx_nq = 0;
            for (int32_t tk = ntask - 1; tk >= 0; --tk) {
                int32_t *rec = pp.t_info.data() + (size_t)(task0 + tk) * PanelPlan::TREC;
                rec[7] = nx_off; rec[8] = nx_nq;
                if (rec[1] >> 30) { nx_off = rec[6]; nx_nq = rec[4]; }
            }
            const int32_t ui[PanelPlan::UREC] = {P, x0, x1, task0, ntask, dmode, nx_off, nx_nq, nr, next, pp.p_row0[P], pp.p_base[P]};
            pp.u_info.insert(pp.u_info.end(), ui, ui + PanelPlan::UREC);
        }
        for (int c = 0; c < W; ++c) pos_in_P[C[c]] = -1;
    }
    pp.ready = true;
    return "";
}

}  // namespace kb2
