// Front plan of the window LU (see FrontPlan in kb2_internal.h and k_lu_window in kb2_front.cuh).
//
// The block plan (kb2_panel.cpp) is left-looking: a panel gathers the updates of all its source
// blocks, and every U' block is fetched again by each later panel that has it as a source.  The
// front plan schedules the SAME arithmetic right-looking: after panel P is finished, its update
//     W[i, j] -= L'[i, P] * U'[P, j]      i in Lrows(P), j in Ucols(P)
// is applied at once to the active submatrix, which lives in shared memory (the "window").  A row
// (column) enters the window at the first front that touches it and leaves when it is eliminated,
// so every factor value is written to HBM exactly once and never read again by the factorisation,
// and W itself is never materialised: window entries receive their original value from the compact
// Jacobian values the moment their row and column are both active.  Every entry receives its updates in
// ascending source order and ascending pivot order inside a source, exactly like the block plan,
// so the two factorisations agree bit for bit.
// Everything here is computed once on the host and shared by all ensemble members.
#include "kb2_internal.h"

#include <algorithm>

namespace kb2 {

std::string build_fronts(Symbolic &sym, int64_t S)
{
    const PanelPlan &pp = sym.panels;
    FrontPlan &fp = sym.fronts;
    fp = FrontPlan();
    if (!pp.ready) return "block plan missing";
    const int PR = PanelPlan::PR;
    const int32_t NP = (int32_t)pp.p_row0.size();
    fp.NF = NP;
    // ---- Lrows: for every panel Q the later panels P that have Q as a (complete) source block,
    // with the position of Q's first column in P's pattern ----
    struct Tgt { int32_t P, lpos; };
    std::vector<std::vector<Tgt>> targets(NP);
    std::vector<int32_t> act_rowpanel(NP), act_col(S);
    for (int32_t P = 0; P < NP; ++P) {
        act_rowpanel[P] = P;
        const int32_t *C = pp.cols.data() + pp.p_cptr[P];
        const int next = pp.p_next[P];
        for (int e = 0; e < next;) {
            const int32_t Q = pp.row_panel[C[e]];
            if (pp.row_r[C[e]] != 0) return "internal error: source block is not complete";
            targets[Q].push_back({P, e});
            act_rowpanel[P] = std::min(act_rowpanel[P], Q);
            e += pp.p_nrows[Q];
        }
    }
    for (int64_t j = 0; j < S; ++j) act_col[j] = pp.row_panel[j];
    for (int32_t P = 0; P < NP; ++P) {
        const int32_t *C = pp.cols.data() + pp.p_cptr[P];
        for (int c = pp.p_next[P] + pp.p_nrows[P]; c < pp.p_width[P]; ++c) act_col[C[c]] = std::min(act_col[C[c]], P);
    }
    // activation lists per front
    std::vector<std::vector<int32_t>> new_rowpanels(NP), new_cols(NP);
    for (int32_t P = 0; P < NP; ++P) new_rowpanels[act_rowpanel[P]].push_back(P);
    for (int64_t j = 0; j < S; ++j) new_cols[act_col[j]].push_back((int32_t)j);
    // original entries by row (exact LU pattern rows carry the J entry of every slot) and by column
    std::vector<int32_t> ct_ptr(S + 1, 0), ct_row, ct_src;
    for (int64_t q = 0; q < sym.nnzLU; ++q) ct_ptr[sym.colidx[q] + 1]++;
    for (int64_t j = 0; j < S; ++j) ct_ptr[j + 1] += ct_ptr[j];
    ct_row.assign(sym.nnzLU, 0); ct_src.assign(sym.nnzLU, 0);
    {
        std::vector<int32_t> fill(ct_ptr.begin(), ct_ptr.end() - 1);
        for (int64_t i = 0; i < S; ++i)
            for (int64_t q = sym.rowptr[i]; q < sym.rowptr[i + 1]; ++q) {
                const int32_t z = fill[sym.colidx[q]]++;
                ct_row[z] = (int32_t)i; ct_src[z] = sym.slot_src[q];
            }
    }
    // ---- slot simulation ----
    // Inactive window entries are kept at zero (the kernel clears the slots of the pivot rows and
    // columns when a front is done), so only the original values have to be written when an
    // entry becomes live: at the later of its row's and its column's activation.  Free slots are
    // handed out oldest first, so that a slot given up by front P-1 is not taken again at front P
    // while anything else is free; a front that does take one is flagged (its values may only be
    // written after the slot has been cleared: one more barrier in the kernel).
    std::vector<int32_t> rslot(S, -1), cslot(S, -1), rfreed(S + 8, -2), cfreed(S + 8, -2);   // front at which a slot was given up
    std::vector<int32_t> free_r, free_c;       // FIFO queues (index of the head kept separately)
    size_t head_r = 0, head_c = 0;
    int32_t nrs = 0, ncs = 0;
    auto take = [](std::vector<int32_t> &fr, size_t &head, int32_t &n) {
        if (head == fr.size()) return n++;
        return fr[head++];
    };
    struct InitE { int32_t r, c, src; };      // window row slot, column slot, source word
    std::vector<std::vector<InitE>> init_of(NP);
    std::vector<int32_t> hot(NP, 0);
    std::vector<int32_t> active_rows, active_cols;
    for (int32_t P = 0; P < NP; ++P) {
        std::vector<int32_t> nrows_new, ncols_new;
        for (int32_t Pn : new_rowpanels[P])
            for (int r = 0; r < pp.p_nrows[Pn]; ++r) nrows_new.push_back(pp.p_row0[Pn] + r);
        ncols_new = new_cols[P];
        std::sort(nrows_new.begin(), nrows_new.end());
        std::sort(ncols_new.begin(), ncols_new.end());
        for (int32_t i : nrows_new) {
            rslot[i] = take(free_r, head_r, nrs);
            if ((int32_t)rfreed.size() <= rslot[i]) rfreed.resize(rslot[i] + 1, -2);
            if (rfreed[rslot[i]] == P - 1) hot[P] = 1;
            active_rows.push_back(i);
        }
        for (int32_t j : ncols_new) {
            cslot[j] = take(free_c, head_c, ncs);
            if ((int32_t)cfreed.size() <= cslot[j]) cfreed.resize(cslot[j] + 1, -2);
            if (cfreed[cslot[j]] == P - 1) hot[P] = 1;
            active_cols.push_back(j);
        }
        // original values of (new row, active column) ...
        for (int32_t i : nrows_new)
            for (int64_t q = sym.rowptr[i]; q < sym.rowptr[i + 1]; ++q) {
                const int32_t j = (int32_t)sym.colidx[q], src = sym.slot_src[q];
                if (cslot[j] >= 0 && src != 0) init_of[P].push_back({rslot[i], cslot[j], src});
            }
        // ... and of (row that was active before, new column)
        for (int32_t j : ncols_new)
            for (int32_t z = ct_ptr[j]; z < ct_ptr[j + 1]; ++z) {
                const int32_t i = ct_row[z], src = ct_src[z];
                if (rslot[i] < 0 || src == 0) continue;
                if (!std::binary_search(nrows_new.begin(), nrows_new.end(), i)) init_of[P].push_back({rslot[i], cslot[j], src});
            }
        fp.max_init = std::max<int32_t>(fp.max_init, (int32_t)init_of[P].size());
        // ---- the front's lists (slots are final from here on) ----
        const int nr = pp.p_nrows[P], next = pp.p_next[P], W = pp.p_width[P], p0 = pp.p_row0[P];
        const int32_t *C = pp.cols.data() + pp.p_cptr[P];
        const int nu = W - next - nr;
        int nl = 0;
        for (const Tgt &t : targets[P]) nl += pp.p_nrows[t.P];
        const int32_t loff = (int32_t)fp.lists.size();
        for (int r = 0; r < PR; ++r) fp.lists.push_back(r < nr ? rslot[p0 + r] : -1);
        for (int r = 0; r < PR; ++r) fp.lists.push_back(r < nr ? cslot[p0 + r] : -1);
        {
            // U columns in ascending SLOT order (neighbouring lanes of the update then touch
            // neighbouring shared-memory words), each with its position in the panel's U part
            std::vector<std::pair<int32_t, int32_t>> us;
            for (int c = next + nr; c < W; ++c) {
                if (cslot[C[c]] < 0) return "internal error: U column without a window slot";
                us.emplace_back(cslot[C[c]], c - next - nr);
            }
            std::sort(us.begin(), us.end());
            for (auto &e : us) fp.lists.push_back(e.first);
            for (auto &e : us) fp.lists.push_back(e.second);
        }
        for (const Tgt &t : targets[P])
            for (int r = 0; r < pp.p_nrows[t.P]; ++r) {
                if (rslot[pp.p_row0[t.P] + r] < 0) return "internal error: L row without a window slot";
                fp.lists.push_back(rslot[pp.p_row0[t.P] + r]);
            }
        for (const Tgt &t : targets[P]) {
            const int nrt = pp.p_nrows[t.P];
            for (int r = 0; r < nrt; ++r) {
                const int64_t slot0 = (int64_t)pp.p_base[t.P] + (int64_t)t.lpos * nrt + r;
                if (slot0 >= ((int64_t)1 << 28)) return "window LU: panel storage exceeds 2^28 slots";
                fp.lists.push_back((int32_t)slot0 | ((nrt - 1) << 28));
            }
        }
        fp.max_nl = std::max(fp.max_nl, nl);
        fp.max_nu = std::max(fp.max_nu, nu);
        // look-ahead: the pivot block of P can be factorised while front P-1 is still updating the
        // window if all of its rows and columns were active before P (none of its entries is new).
        // (Extending it to every front — new entries taken from their original values — was measured
        // slower: warp 0 is missing from the update of a look-ahead front)
        int32_t la = P >= 1 && act_rowpanel[P] < P;
        for (int r = 0; r < nr && la; ++r) la = act_col[p0 + r] < P;
        const int32_t rec[FrontPlan::FREC] = {nr, p0, nu, nl, pp.p_base[P], next, loff, 0, (int32_t)init_of[P].size(), hot[P], la, 0};
        fp.f_info.insert(fp.f_info.end(), rec, rec + FrontPlan::FREC);
        // ---- eliminate: the pivot rows and columns leave the window ----
        for (int r = 0; r < nr; ++r) {
            free_r.push_back(rslot[p0 + r]); rfreed[rslot[p0 + r]] = P; rslot[p0 + r] = -1;
            free_c.push_back(cslot[p0 + r]); cfreed[cslot[p0 + r]] = P; cslot[p0 + r] = -1;
        }
        active_rows.erase(std::remove_if(active_rows.begin(), active_rows.end(), [&](int32_t i) { return rslot[i] < 0; }), active_rows.end());
        active_cols.erase(std::remove_if(active_cols.begin(), active_cols.end(), [&](int32_t j) { return cslot[j] < 0; }), active_cols.end());
    }
    fp.Wr = nrs;
    fp.Wc = ncs | 1;        // odd row pitch: consecutive row slots fall into different shared-memory bank groups
    for (int32_t P = 0; P < NP; ++P) {
        fp.f_info[(size_t)P * FrontPlan::FREC + 7] = (int32_t)(fp.init.size() / 2);
        for (const InitE &e : init_of[P]) {
            fp.init.push_back(e.r * fp.Wc + e.c);
            fp.init.push_back(e.src);
        }
    }
    fp.ready = true;
    return "";
}

}  // namespace kb2
