// Host-side data model shared by kb2_symbolic.cpp and kb2_api.cu.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace kb2 {

// Reaction table as handed over the C ABI (RxData flattened; reference
// src/exploration/network.jl:193-203).
struct Network {
    int64_t S = 0, R = 0;
    std::vector<int64_t> rp, ri, rn;   // reactants CSR
    std::vector<int64_t> pp, pi, pn;   // products  CSR
    // derived
    std::vector<int32_t> sub_ptr, sub_idx, sub_exp;   // merged substrate exponents per reaction
    std::vector<int32_t> net_ptr, net_idx, net_coef;  // net stoichiometry per reaction (zeros dropped)
};

// Block plan of the supernodal (block Crout) LU and the panel triangular solves (kb2_panel.cpp).
// Rows of the permuted L\U pattern are grouped into panels of <= PR consecutive rows sharing one
// padded column pattern; the L part of a panel is a union of COMPLETE earlier panels (source
// blocks), so every panel-panel update is a small dense  nr x nq  times  nq x ntargets  product.
struct PanelPlan {
    static constexpr int PR = 8;      // rows per panel
    static constexpr int CW = 96;     // columns of a panel held in shared memory at a time (one chunk)
    bool ready = false;
    int64_t padded = 0, n_fma_padded = 0;
    int32_t max_width = 0;
    // per panel: first row, rows, pattern width, position of the diagonal block in the pattern,
    // first storage slot, offset of the pattern in `cols`
    std::vector<int32_t> p_row0, p_nrows, p_width, p_next, p_base, p_cptr, cols;
    std::vector<int32_t> row_panel, row_r;
    std::vector<int32_t> slot_of;     // exact-pattern slot -> storage slot  p_base + c*nr + r
    std::vector<int32_t> jslot;       // Jacobian entry (CSC order) -> storage slot
    std::vector<int32_t> diag_slot;   // per pivot row
    // units = (panel, column chunk [x0,x1)); UREC ints each:
    // {panel, x0, x1, first task, tasks, diagonal mode (0 later chunk, 1 in this chunk, 2 earlier chunk),
    //  slot of U'_QQ of the first in-chunk source (-1 none), its nq, nr, next, first row, base slot}
    static constexpr int UREC = 12, TREC = 12;
    std::vector<int32_t> u_info;
    // tasks = source blocks applied to a unit; TREC ints each:
    // {source panel Q, position of Q's first column in the target pattern | in-chunk << 30, targets, first map entry,
    //  nq, base slot of Q, slot of U'_QQ, slot of U'_QQ of the next in-chunk source (-1 none), its nq,
    //  first slot and number of slots of the span of U' columns the targets touch, 0}
    std::vector<int32_t> t_info;
    std::vector<int32_t> map;         // per target: (column position in Q's pattern) | (column position in the chunk << 16)
    int64_t n_units() const { return (int64_t)u_info.size() / UREC; }
    int64_t n_tasks() const { return (int64_t)t_info.size() / TREC; }
};

// Front plan of the window LU (kb2_front.cpp): the same factorisation as the block plan, right-
// looking, with the active submatrix (rows and columns that have received or will receive
// updates and are not eliminated yet) resident in shared memory.  Front P = panel P: its pivot
// block, the rows below it that have P as a source block (Lrows) and its U-part columns (Ucols).
// Rows and columns own a window slot from their first touch until they are eliminated.
struct FrontPlan {
    bool ready = false;
    int32_t NF = 0, Wr = 0, Wc = 0, max_nl = 0, max_nu = 0, max_init = 0;
    static constexpr int FREC = 12;
    // per front: {nr, first pivot row, nu, nl, base slot of the panel, position of the diagonal block,
    //             offset into `lists`, offset into `init` (entries), init entries,
    //             1 if a row/column of this front takes a slot that the previous front gave up,
    //             1 if the pivot block may be factorised ahead (no entry of it is new at this front), 0}
    std::vector<int32_t> f_info;
    // per front, at its offset: prs[8] pcs[8] (window slots of the pivot rows / columns, -1 beyond nr),
    // ucs[nu] (column slots of Ucols, ascending), ujj[nu] (position of that column in the panel's U
    // part), lrs[nl] (row slots of Lrows), lgs[nl] (storage slot of L'(row, first pivot column) |
    // (rows of that row's panel - 1) << 28: the stride between pivots)
    std::vector<int32_t> lists;
    // window entries that become live when front P starts and have an original value (inactive
    // entries are zero): {window position rslot*Wc + cslot, (J entry + 1) << 1 | is_diagonal}
    std::vector<int32_t> init;
};

// Everything the kernels need that depends only on the network (shared by all members).
struct Symbolic {
    bool ready = false;
    int64_t nnzJ = 0, nnzLU = 0, n_fma = 0;
    // Jacobian pattern, CSC, rows ascending (pattern contract of SURVEY.md §8a R5)
    std::vector<int64_t> colptr, rowval;
    std::vector<int64_t> perm, iperm;                 // perm[a] = species at pivot position a
    // combined L\U pattern of P(W)P^T, row CSR, columns ascending
    std::vector<int64_t> rowptr, colidx, diagpos;
    // --- device-ready int32 tables ---
    // RHS gather CSR by species: entries (reaction, net coefficient)
    std::vector<int32_t> rhs_ptr, rhs_rxn, rhs_coef;
    // work order of the gather loops: rows / entries sorted by length, longest first; the first
    // n_long of them are split across the lanes of a member, the rest go one per lane
    static constexpr int RHS_LONG = 64, JAC_LONG = 16;
    std::vector<int32_t> rhs_order, j_order;
    int32_t rhs_nlong = 0, j_nlong = 0;
    // sliced ELL of the one-per-lane RHS rows (groups of ELL_G rows of rhs_order after the long ones):
    // ell[ell_ptr[g] + t*ELL_G + rho] = coef << 24 | reaction
    static constexpr int ELL_G = 64;
    std::vector<int32_t> ell_ptr, ell;
    // Jacobian: derivative-table slots per reaction (max distinct reactants), terms as
    // (index j*jslots + slot, coefficient), packed coef << 24 | index, and their sliced ELL
    int32_t jslots = 1;
    std::vector<int32_t> jt_idx, jt_coef, jt_pk, jell_ptr, jell;
    // first-touch layouts of the gathered tables: rate of reaction j at rate_pos[j], derivative
    // (j, s) at drate_pos[j*jslots + s]; rhs_src = rhs_rxn mapped through rate_pos
    std::vector<int32_t> rate_pos, drate_pos, rhs_src;
    // reaction descriptors: up to 3 distinct reactant species + exponents packed 8 bit each
    std::vector<int32_t> rdesc;                       // 4 ints per reaction
    // Jacobian terms by J entry (CSC order): (reaction, (coef*nu_l) << 2 | reactant slot)
    std::vector<int32_t> jt_ptr, jt_rxn, jt_pack;
    // per LU slot: ((J entry + 1) << 1) | is_diagonal
    std::vector<int32_t> slot_src;
    std::vector<int32_t> lu_rowptr, lu_colidx, lu_diagpos;
    int32_t max_rowlen = 0;
    PanelPlan panels;
    FrontPlan fronts;
};

// Builds derived stoichiometry; returns "" or an error message.
std::string build_network(Network &net);
// ordering: 0 min degree, 1 natural, 2 user (sym.perm preset), 3 natural with hub species last,
// 5 reverse Cuthill-McKee / 6, 7 Sloan (weights 1:2, 2:1) on the graph without the hubs, hubs last
// env_limit / fma_limit let `auto` drop a candidate whose fill explodes (a banded order on a network without
// locality) at bounded cost: give up ("profile beyond the limit") before the symbolic LU when the envelope of
// the permuted symmetrised pattern (an upper bound of the fill of a banded order) exceeds env_limit, and
// ("fill beyond the limit") as soon as the symbolic LU has counted more than fma_limit FMAs
std::string build_symbolic(const Network &net, int ordering, Symbolic &sym, int64_t fma_limit = INT64_MAX,
                           int64_t env_limit = INT64_MAX);
std::string build_panels(Symbolic &sym, int64_t S);
std::string build_fronts(Symbolic &sym, int64_t S);

}  // namespace kb2
