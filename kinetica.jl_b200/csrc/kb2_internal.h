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

// Panel plan of the register-blocked LU / panel triangular solves (kb2_panel.cpp).
struct PanelPlan {
    static constexpr int PR = 8;      // rows per panel
    static constexpr int CW = 128;    // columns per chunk (32 column lanes x 4 registers)
    bool ready = false;
    int64_t padded = 0, n_fma_padded = 0;
    int32_t max_width = 0;
    std::vector<int32_t> p_row0, p_nrows, p_width, p_next, p_base, p_cptr, cols;
    std::vector<int32_t> row_panel, row_r;
    std::vector<int32_t> slot_of;     // exact-pattern slot -> storage slot
    std::vector<int32_t> slot_src;    // per storage slot: ((J entry + 1) << 1) | is_diagonal
    std::vector<int32_t> diag_slot;   // per row
    struct Unit { int32_t panel, x0, x1, step0, n_pre, n_ext, map0, n_maps, diag_here, diag_before, block0 = 0, n_blocks = 0; };
    std::vector<Unit> units;
    std::vector<int32_t> s_e, s_k, s_src, s_map, maps;
    static constexpr int SEG = 64;    // source slots staged per step
    static constexpr int NB = 4;      // steps per block (one barrier round)
    std::vector<int32_t> b_info;      // [block] {first step (unit relative), steps, kind 0 pre / 1 in-chunk, owner lane}
    std::vector<int32_t> b_idx;       // [block][32 column lanes][NB] packed byte-index words
    std::vector<int32_t> s_meta;      // [step] {first source slot, slots, pivot column, flags}
    std::vector<int32_t> s_idx;       // [step][32 column lanes] 4 x int8 index into the staged range (-1 none)
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
    // reaction descriptors: up to 3 distinct reactant species + exponents packed 8 bit each
    std::vector<int32_t> rdesc;                       // 4 ints per reaction
    // Jacobian terms by J entry (CSC order): (reaction, (coef*nu_l) << 2 | reactant slot)
    std::vector<int32_t> jt_ptr, jt_rxn, jt_pack;
    // per LU slot: ((J entry + 1) << 1) | is_diagonal
    std::vector<int32_t> slot_src;
    std::vector<int32_t> lu_rowptr, lu_colidx, lu_diagpos;
    // elimination schedule: for L slot p (row i, pivot k): targets of U(k,:) as offsets into row i
    std::vector<uint32_t> tgt_off;                    // per slot (only L slots meaningful)
    std::vector<int32_t> tgt;                         // n_fma entries
    int32_t max_rowlen = 0;
    PanelPlan panels;
};

// Builds derived stoichiometry; returns "" or an error message.
std::string build_network(Network &net);
// ordering: 0 min degree, 1 natural, 2 user (sym.perm preset)
std::string build_symbolic(const Network &net, int ordering, Symbolic &sym);
std::string build_panels(Symbolic &sym, int64_t S);

}  // namespace kb2
