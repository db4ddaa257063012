// Register-blocked panel LU and panel triangular solves (sm_100a).
//
// Thread layout inside a tile CTA: blockDim = 32 * MB; thread = (m = tid % MB, cs = tid / MB),
// cs in 0..31 is the *column lane*.  While a unit (panel x column chunk) is being eliminated,
// thread (m, cs) holds the PR x NQ targets  w[r][q] = W[row r of the panel][column x0 + cs + 32 q]
// of member m in registers; pivots stream past: the multipliers l[0..PR) are shared through a
// double-buffered shared-memory slot (one block barrier per in-chunk pivot), the pivot row u_kj
// comes straight from global/L2 — one coalesced load feeds PR FMAs — and finished multipliers of
// earlier chunks are re-read from global without any barrier.  Per FMA this is 1/PR global loads
// and no shared-memory traffic on the targets, which is what lifts the sparse LU off the
// shared-memory roofline.
#pragma once
#include "kb2_kernels.cuh"

namespace kb2 {

constexpr int PR = 8;
constexpr int NQ = 4;
constexpr int CW = 32 * NQ;

struct DevPlan {
    int npanels, nunits, padded;
    const int *p_row0, *p_nrows, *p_width, *p_next, *p_base, *p_cptr, *cols;
    const int *u_panel, *u_x0, *u_x1, *u_step0, *u_npre, *u_next, *u_map0, *u_diag;   // u_diag: 0 later, 1 here, 2 before
    const int *s_e, *s_k, *s_src, *s_map, *maps;
};

// In-place LU of the padded panel storage.  lbuf: 2 * PR * MB doubles of shared memory.
template <int MB>
__device__ void tile_lu_panels(const Tile<MB> &tl, const DevPlan &pl, double *lu,
                               double *invd, double *lbuf)
{
    const int cs = tl.slot, m = tl.m, b = tl.b;
    const size_t Bp = tl.Bp;
    int par = 0;
    for (int un = 0; un < pl.nunits; ++un) {
        const int P = pl.u_panel[un], x0 = pl.u_x0[un], x1 = pl.u_x1[un];
        const int W = pl.p_width[P], nr = pl.p_nrows[P], base = pl.p_base[P], next = pl.p_next[P], p0 = pl.p_row0[P];
        const int s0 = pl.u_step0[un], npre = pl.u_npre[un], nin = pl.u_next[un], dflag = pl.u_diag[un];
        const int *__restrict__ maps = pl.maps + (size_t)pl.u_map0[un] * CW;
        double w[PR][NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int c = x0 + cs + 32 * q;
#pragma unroll
            for (int r = 0; r < PR; ++r)
                w[r][q] = (c < x1 && r < nr) ? lu[(size_t)(base + r * W + c) * Bp + b] : 0.0;
        }
        double l[PR];
        // ---- pivots left of this chunk: multipliers are final in global memory, no barrier ----
        for (int s = s0; s < s0 + npre; ++s) {
            const int e = pl.s_e[s], src = pl.s_src[s];
            const int *__restrict__ mp = maps + (size_t)pl.s_map[s] * CW;
#pragma unroll
            for (int r = 0; r < PR; ++r) l[r] = (r < nr) ? lu[(size_t)(base + r * W + e) * Bp + b] : 0.0;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const int mq = mp[cs + 32 * q];
                if (mq >= 0) {
                    const double u = lu[(size_t)(src + mq) * Bp + b];
#pragma unroll
                    for (int r = 0; r < PR; ++r) w[r][q] -= l[r] * u;
                }
            }
        }
        if (dflag == 2) {
            // the panel's own rows are pivots too (diagonal block in an earlier chunk)
#pragma unroll
            for (int r0 = 0; r0 < PR - 1; ++r0) {
                if (r0 < nr - 1) {
#pragma unroll
                    for (int r = r0 + 1; r < PR; ++r) l[r] = (r < nr) ? lu[(size_t)(base + r * W + next + r0) * Bp + b] : 0.0;
#pragma unroll
                    for (int q = 0; q < NQ; ++q)
#pragma unroll
                        for (int r = r0 + 1; r < PR; ++r) w[r][q] -= l[r] * w[r0][q];
                }
            }
        }
        // ---- external pivots inside this chunk: finalise the multiplier, publish, update ----
        for (int s = s0 + npre; s < s0 + npre + nin; ++s) {
            const int e = pl.s_e[s], ce = e - x0, src = pl.s_src[s], mi = pl.s_map[s];
            double *lb = lbuf + par * PR * MB;
            if (cs == (ce & 31)) {
                const double d = invd[(size_t)pl.s_k[s] * Bp + b];
                const int qe = ce >> 5;
#pragma unroll
                for (int r = 0; r < PR; ++r) {
                    double v = w[r][0];               // select-based register pick: keeps w[][] out of local memory
#pragma unroll
                    for (int q = 1; q < NQ; ++q) v = (qe == q) ? w[r][q] : v;
                    v *= d;
#pragma unroll
                    for (int q = 0; q < NQ; ++q) w[r][q] = (qe == q) ? v : w[r][q];
                    lb[r * MB + m] = v;
                }
            }
            __syncthreads();
            if (mi >= 0) {
                const int *__restrict__ mp = maps + (size_t)mi * CW;
#pragma unroll
                for (int r = 0; r < PR; ++r) l[r] = lb[r * MB + m];
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    const int cl = cs + 32 * q;
                    const int mq = mp[cl];
                    if (mq >= 0 && cl > ce) {
                        const double u = lu[(size_t)(src + mq) * Bp + b];
#pragma unroll
                        for (int r = 0; r < PR; ++r) w[r][q] -= l[r] * u;
                    }
                }
            }
            par ^= 1;
        }
        if (dflag == 1) {
            // ---- the panel's own diagonal block ----
#pragma unroll
            for (int r0 = 0; r0 < PR; ++r0) {
                if (r0 < nr) {
                    const int ce = next + r0 - x0;
                    double *lb = lbuf + par * PR * MB;
                    if (cs == (ce & 31)) {
                        const int qe = ce >> 5;
                        double piv = w[r0][0];
#pragma unroll
                        for (int q = 1; q < NQ; ++q) piv = (qe == q) ? w[r0][q] : piv;
                        const double inv = 1.0 / piv;
                        invd[(size_t)(p0 + r0) * Bp + b] = inv;
#pragma unroll
                        for (int r = r0 + 1; r < PR; ++r) {
                            double v = w[r][0];
#pragma unroll
                            for (int q = 1; q < NQ; ++q) v = (qe == q) ? w[r][q] : v;
                            v *= inv;
#pragma unroll
                            for (int q = 0; q < NQ; ++q) w[r][q] = (qe == q) ? v : w[r][q];
                            lb[r * MB + m] = v;
                        }
                    }
                    if (r0 < nr - 1) {
                        __syncthreads();
#pragma unroll
                        for (int r = r0 + 1; r < PR; ++r) l[r] = lb[r * MB + m];
#pragma unroll
                        for (int q = 0; q < NQ; ++q) {
                            if (cs + 32 * q > ce) {
#pragma unroll
                                for (int r = r0 + 1; r < PR; ++r) w[r][q] -= l[r] * w[r0][q];
                            }
                        }
                        par ^= 1;
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int c = x0 + cs + 32 * q;
            if (c < x1) {
#pragma unroll
                for (int r = 0; r < PR; ++r)
                    if (r < nr) lu[(size_t)(base + r * W + c) * Bp + b] = w[r][q];
            }
        }
        __syncthreads();
    }
}

// W x = rhs over the panel storage.  rhs, x in species order; y = permuted scratch [S][Bp].
// red: (blockDim/32) * PR * MB doubles of shared memory.
template <int MB>
__device__ void tile_trisolve_panels(const Tile<MB> &tl, const DevNet &net, const DevPlan &pl,
                                     const double *lu, const double *invd,
                                     const double *rhs, double *y,
                                     double *x, double *red)
{
    const int cs = tl.slot, m = tl.m, b = tl.b;
    const size_t Bp = tl.Bp;
    const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const bool wlead = (threadIdx.x & 31) < MB;
    // solver threads: (mm, r) with r fastest, so that the PR rows of one member sit in PR adjacent lanes
    const int t = threadIdx.x;
    const int nsolve = (MB * PR + 31) / 32 * 32;
    const bool swarp = t < nsolve;
    const bool sval = t < MB * PR;
    const int sm_ = sval ? t / PR : 0, sr = t % PR;
    const int sb = tl.b - m + sm_;
    // ---------------- forward:  L y = P rhs ----------------
    for (int P = 0; P < pl.npanels; ++P) {
        const int W = pl.p_width[P], nr = pl.p_nrows[P], base = pl.p_base[P], next = pl.p_next[P], p0 = pl.p_row0[P];
        const int *__restrict__ C = pl.cols + pl.p_cptr[P];
        double acc[PR];
#pragma unroll
        for (int r = 0; r < PR; ++r) acc[r] = 0.0;
        for (int c = cs; c < next; c += 32) {
            const double yc = y[(size_t)C[c] * Bp + b];
#pragma unroll
            for (int r = 0; r < PR; ++r)
                if (r < nr) acc[r] += lu[(size_t)(base + r * W + c) * Bp + b] * yc;
        }
        double lint[PR - 1], z = 0.0;
        if (swarp) {
            const bool ok = sval && sr < nr;
#pragma unroll
            for (int rp = 0; rp < PR - 1; ++rp)
                lint[rp] = (ok && rp < sr) ? lu[(size_t)(base + sr * W + next + rp) * Bp + sb] : 0.0;
            z = ok ? rhs[(size_t)net.perm[p0 + sr] * Bp + sb] : 0.0;
        }
        if (next > 0) {
#pragma unroll
            for (int r = 0; r < PR; ++r) {
                const double v = warp_sum<MB>(acc[r]);
                if (wlead) red[(warp * PR + r) * MB + m] = v;
            }
            __syncthreads();
        }
        if (swarp) {
            if (next > 0)
                for (int q = 0; q < nw; ++q) z -= red[(q * PR + sr) * MB + sm_];
#pragma unroll
            for (int rp = 0; rp < PR - 1; ++rp) {
                const double yv = __shfl_sync(0xffffffffu, z, rp, PR);
                if (sr > rp) z -= lint[rp] * yv;
            }
            if (sval && sr < nr) y[(size_t)(p0 + sr) * Bp + sb] = z;
        }
        __syncthreads();
    }
    // ---------------- backward:  U x = y ----------------
    for (int P = pl.npanels - 1; P >= 0; --P) {
        const int W = pl.p_width[P], nr = pl.p_nrows[P], base = pl.p_base[P], next = pl.p_next[P], p0 = pl.p_row0[P];
        const int *__restrict__ C = pl.cols + pl.p_cptr[P];
        double acc[PR];
#pragma unroll
        for (int r = 0; r < PR; ++r) acc[r] = 0.0;
        const int u0 = next + nr;
        for (int c = u0 + cs; c < W; c += 32) {
            const double yc = y[(size_t)C[c] * Bp + b];
#pragma unroll
            for (int r = 0; r < PR; ++r)
                if (r < nr) acc[r] += lu[(size_t)(base + r * W + c) * Bp + b] * yc;
        }
        double uint_[PR], z = 0.0, dinv = 0.0;
        if (swarp) {
            const bool ok = sval && sr < nr;
#pragma unroll
            for (int rp = 0; rp < PR; ++rp)
                uint_[rp] = (ok && rp > sr && rp < nr) ? lu[(size_t)(base + sr * W + next + rp) * Bp + sb] : 0.0;
            z = ok ? y[(size_t)(p0 + sr) * Bp + sb] : 0.0;
            dinv = ok ? invd[(size_t)(p0 + sr) * Bp + sb] : 0.0;
        }
        if (W > u0) {
#pragma unroll
            for (int r = 0; r < PR; ++r) {
                const double v = warp_sum<MB>(acc[r]);
                if (wlead) red[(warp * PR + r) * MB + m] = v;
            }
            __syncthreads();
        }
        if (swarp) {
            if (W > u0)
                for (int q = 0; q < nw; ++q) z -= red[(q * PR + sr) * MB + sm_];
#pragma unroll
            for (int rp = PR - 1; rp >= 0; --rp) {
                const double mine = z * dinv;                       // final value if this lane's row is rp
                const double xv = __shfl_sync(0xffffffffu, mine, rp, PR);
                if (sr < rp) z -= uint_[rp] * xv;
            }
            if (sval && sr < nr) {
                const double v = z * dinv;
                y[(size_t)(p0 + sr) * Bp + sb] = v;
                x[(size_t)net.perm[p0 + sr] * Bp + sb] = v;
            }
        }
        __syncthreads();
    }
}

}  // namespace kb2
