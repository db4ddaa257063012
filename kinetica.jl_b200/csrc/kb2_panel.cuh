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
    const int *u_block0, *u_nblocks, *b_info, *b_idx;
    int dbg;      // timing experiments only (KB2_DBG): bit0 no staging, bit1 no application, bit2 no owner work, bit3 no barriers, bit4 no unit load/store
    const int *s_e, *s_k, *s_meta, *s_idx;
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];\n" ::"l"(p)); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

constexpr int LU_NB = 4;      // pivot steps per block (one barrier round)
constexpr int LU_LOOK = 8;    // segments are staged up to this many steps beyond the current block
constexpr int LU_NS = 16;     // stages in the ring
constexpr int LU_SEG = 64;    // source slots per stage

// shared memory of tile_lu_panels, in doubles, for MB members per tile
__host__ __device__ constexpr size_t lu_smem_doubles(int nt, int mb)
{
    (void)nt;
    return (size_t)2 * LU_NB * PR * mb     // published multipliers of a block (double buffered)
           + (size_t)4 * LU_NB * PR * mb   // multipliers of pivots left of the chunk (ring over blocks)
           + (size_t)(LU_NS * LU_SEG + 1) * mb;  // staged pivot-row segments + one constant zero row
}

// In-place factorisation W = L' U' of the padded panel storage in Crout form: U' has a unit
// diagonal (rows are scaled by 1/d_i when their panel finishes), L' carries the pivots, so no
// division sits on the per-pivot critical path.  invd[i] = 1/d_i is kept for the forward
// substitution.
//
// Pivots are processed in blocks of up to LU_NB steps.  For every step the CTA copies the needed
// segment of the pivot row (one contiguous slot range of the source panel, MB members wide) into
// a shared-memory stage with one 16-byte cp.async per thread, up to LU_LOOK steps ahead; a lane
// finds its up to NQ targets with one packed byte-index word per step.  The NQ consecutive pivot
// columns of an in-chunk block live in ONE thread, which resolves the dependencies among them
// and publishes all their multipliers at once: two block barriers per LU_NB pivots.
template <int MB>
__device__ void tile_lu_panels(const Tile<MB> &tl, const DevPlan &pl, double *lu, double *invd, double *smem)
{
    const int cs = tl.slot, m = tl.m, b = tl.b, tid = threadIdx.x, nt = blockDim.x;
    const size_t Bp = tl.Bp;
    const int b0 = b - m;
    double *lbuf = smem;                                  // [2][LU_NB][PR][MB]
    double *ring_l = lbuf + 2 * LU_NB * PR * MB;          // [4][LU_NB][PR][MB]
    double *stage = ring_l + 4 * LU_NB * PR * MB;         // [LU_NS][LU_SEG][MB]
    constexpr int PIECE = (MB >= 2) ? 2 : 1;              // doubles per copied piece (16 bytes; 8 for MB = 1)
    constexpr int PPS = MB / PIECE;                       // pieces per slot
    const int my_slot = tid / PPS, my_part = tid % PPS;   // nt / PPS = 64 = LU_SEG slots: one piece per thread per stage
    // inactive targets read u = 0 from a constant zero row, which keeps the update loop branch-free
    double *zero_row = stage + (size_t)LU_NS * LU_SEG * MB;
    if (tid < MB) zero_row[tid] = 0.0;
    __syncthreads();
    int par = 0;
    for (int un = 0; un < pl.nunits; ++un) {
        const int P = pl.u_panel[un], x0 = pl.u_x0[un], x1 = pl.u_x1[un];
        const int W = pl.p_width[P], nr = pl.p_nrows[P], base = pl.p_base[P], next = pl.p_next[P], p0 = pl.p_row0[P];
        const int s0 = pl.u_step0[un], npre = pl.u_npre[un], n = npre + pl.u_next[un], dflag = pl.u_diag[un];
        const int blk0 = pl.u_block0[un], nblk = pl.u_nblocks[un];
        const int4 *meta = reinterpret_cast<const int4 *>(pl.s_meta) + s0;
        const int4 *binfo = reinterpret_cast<const int4 *>(pl.b_info) + blk0;
        const int4 *bidx = reinterpret_cast<const int4 *>(pl.b_idx) + (size_t)blk0 * 32 + cs;
        double w[PR][NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int c = x0 + cs * NQ + q;
#pragma unroll
            for (int r = 0; r < PR; ++r)
                w[r][q] = (c < x1 && r < nr && !(pl.dbg & 16)) ? lu[(size_t)(base + r * W + c) * Bp + b] : 0.0;
        }
        int issued = 0;                  // steps whose segment copy has been issued
        // stage the segments (and pre-chunk multipliers) of steps [issued, upto)
        auto issue_to = [&](int upto) {
            upto = upto < n ? upto : n;
            for (; issued < upto; ++issued) {
                const int4 mt = meta[issued];
                for (int sl = my_slot; sl < mt.y && !(pl.dbg & 1); sl += nt / PPS) {      // one trip for MB >= 2
                    double *dst = stage + ((size_t)(issued & (LU_NS - 1)) * LU_SEG + sl) * MB + my_part * PIECE;
                    const double *src = lu + (size_t)(mt.x + sl) * Bp + b0 + my_part * PIECE;
                    if (PIECE == 2) cp_async16(dst, src); else cp_async8(dst, src);
                }
            }
        };
        // multipliers of a pre-chunk block: lane (j, r) = (cs / PR, cs % PR) fetches L'(row r, pivot j)
        auto issue_pre_l = [&](int bi) {
            if (bi < nblk) {
                const int4 bi4 = binfo[bi];
                if (!(bi4.z & 1) && (cs / PR) < bi4.y && (cs % PR) < nr) {
                    const int e = meta[bi4.x + cs / PR].z;
                    cp_async8(ring_l + (((size_t)(bi & 3) * LU_NB + cs / PR) * PR + cs % PR) * MB + m,
                              lu + (size_t)(base + (cs % PR) * W + e) * Bp + b);
                }
            }
        };
        issue_to(LU_LOOK);
        issue_pre_l(0);
        cp_async_commit();
        issue_pre_l(1);
        cp_async_commit();
        for (int bi = 0; bi < nblk; ++bi) {
            const int4 bi4 = binfo[bi];                    // {first step, steps, kind, owner lane}
            const int4 iw4 = bidx[(size_t)bi * 32];        // this lane's index words for the block's steps
            if (bi + 3 < nblk) { prefetch_l1(bidx + (size_t)(bi + 3) * 32); prefetch_l1(binfo + bi + 3); prefetch_l1(meta + bi4.x + 12); }
            cp_async_wait<1>();
            issue_to(bi4.x + bi4.y + LU_LOOK);
            issue_pre_l(bi + 2);
            cp_async_commit();
            if (!(pl.dbg & 8)) __syncthreads();                               // stages (and pre-chunk multipliers) of this block are visible
            const double *lsrc;
            if (bi4.z & 1) {
                // in-chunk block: the owner lane resolves its pivots and publishes their multipliers
                double *lb = lbuf + (size_t)par * LU_NB * PR * MB;
                if (cs == bi4.w && !(pl.dbg & 4)) {
                    // (updates its own columns in place; it then skips the generic application below)
#pragma unroll
                    for (int j = 0; j < LU_NB; ++j) {
                        if (j < bi4.y) {
                            const int4 mt = meta[bi4.x + j];
                            const unsigned wd = (j == 0) ? iw4.x : (j == 1) ? iw4.y : (j == 2) ? iw4.z : iw4.w;
                            const int qe = (mt.z - x0) % NQ, lj = mt.w >> 8;
                            double lv[PR];
#pragma unroll
                            for (int r = 0; r < PR; ++r) {
                                double v = w[r][0];
#pragma unroll
                                for (int q = 1; q < NQ; ++q) v = (qe == q) ? w[r][q] : v;
                                lv[r] = v;
                                if (!(mt.w & 1)) lb[(lj * PR + r) * MB + m] = v;
                            }
                            const double *sv = stage + (size_t)((bi4.x + j) & (LU_NS - 1)) * LU_SEG * MB + m;
#pragma unroll
                            for (int q = 1; q < NQ; ++q) {      // only the owner's later pivot columns matter here
                                const int i8 = (wd >> (8 * q)) & 0xff;
                                if (i8 != LU_SEG) {
                                    const double u = sv[i8 * MB];
#pragma unroll
                                    for (int r = 0; r < PR; ++r) w[r][q] -= lv[r] * u;
                                }
                            }
                        }
                    }
                }
                if (!(pl.dbg & 8)) __syncthreads();
                lsrc = lb + m;
                par ^= 1;
                if (cs == bi4.w) continue;
            } else {
                lsrc = ring_l + (size_t)(bi & 3) * LU_NB * PR * MB + m;
            }
#pragma unroll
            for (int j = 0; j < LU_NB; ++j) {
                if (j < bi4.y && !(pl.dbg & 2)) {
                    const unsigned wd = (j == 0) ? iw4.x : (j == 1) ? iw4.y : (j == 2) ? iw4.z : iw4.w;
                    const int lj = (bi4.z >> (8 + 2 * j)) & 3;
                    const double *sv = stage + (size_t)((bi4.x + j) & (LU_NS - 1)) * LU_SEG * MB + m;
                    double l[PR], u[NQ];
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        const int i8 = (wd >> (8 * q)) & 0xff;
                        u[q] = (i8 == LU_SEG) ? zero_row[m] : sv[i8 * MB];
                    }
#pragma unroll
                    for (int r = 0; r < PR; ++r) l[r] = lsrc[(lj * PR + r) * MB];
#pragma unroll
                    for (int q = 0; q < NQ; ++q)
#pragma unroll
                        for (int r = 0; r < PR; ++r) w[r][q] -= l[r] * u[q];
                }
            }
        }
        cp_async_wait<0>();
        double l[PR];
        if (dflag == 2) {
            // diagonal block in an earlier chunk: the panel's own rows are pivots for these columns
            // and each row still has to be scaled by its 1/d
#pragma unroll
            for (int r0 = 0; r0 < PR; ++r0) {
                if (r0 < nr) {
                    const double inv = invd[(size_t)(p0 + r0) * Bp + b];
#pragma unroll
                    for (int q = 0; q < NQ; ++q) w[r0][q] *= inv;
#pragma unroll
                    for (int r = r0 + 1; r < PR; ++r) l[r] = (r < nr) ? lu[(size_t)(base + r * W + next + r0) * Bp + b] : 0.0;
#pragma unroll
                    for (int q = 0; q < NQ; ++q)
#pragma unroll
                        for (int r = r0 + 1; r < PR; ++r) w[r][q] -= l[r] * w[r0][q];
                }
            }
        }
        if (dflag == 1) {
            // ---- the panel's own diagonal block: pivot d = w[r0][diag]; scale row r0 right of it ----
#pragma unroll
            for (int r0 = 0; r0 < PR; ++r0) {
                if (r0 < nr) {
                    const int ce = next + r0 - x0;
                    double *lb = lbuf + par * PR * MB;
                    if (cs == ce / NQ) {
                        const int qe = ce % NQ;
#pragma unroll
                        for (int r = r0; r < PR; ++r) {
                            double v = w[r][0];
#pragma unroll
                            for (int q = 1; q < NQ; ++q) v = (qe == q) ? w[r][q] : v;
                            if (r == r0) { v = 1.0 / v; invd[(size_t)(p0 + r0) * Bp + b] = v; }
                            lb[r * MB + m] = v;
                        }
                    }
                    __syncthreads();
                    const double inv = lb[r0 * MB + m];
#pragma unroll
                    for (int r = r0 + 1; r < PR; ++r) l[r] = lb[r * MB + m];
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        if (cs * NQ + q > ce) {
                            w[r0][q] *= inv;
#pragma unroll
                            for (int r = r0 + 1; r < PR; ++r) w[r][q] -= l[r] * w[r0][q];
                        }
                    }
                    par ^= 1;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int c = x0 + cs * NQ + q;
            if (c < x1) {
#pragma unroll
                for (int r = 0; r < PR; ++r)
                    if (r < nr && !(pl.dbg & 16)) lu[(size_t)(base + r * W + c) * Bp + b] = w[r][q];
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
        double lint[PR - 1], z = 0.0, dinv = 0.0;
        if (swarp) {
            const bool ok = sval && sr < nr;
#pragma unroll
            for (int rp = 0; rp < PR - 1; ++rp)
                lint[rp] = (ok && rp < sr) ? lu[(size_t)(base + sr * W + next + rp) * Bp + sb] : 0.0;
            z = ok ? rhs[(size_t)net.perm[p0 + sr] * Bp + sb] : 0.0;
            dinv = ok ? invd[(size_t)(p0 + sr) * Bp + sb] : 0.0;
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
                const double yv = __shfl_sync(0xffffffffu, z * dinv, rp, PR);   // y_rp = z_rp / d_rp is final here
                if (sr > rp) z -= lint[rp] * yv;
            }
            if (sval && sr < nr) y[(size_t)(p0 + sr) * Bp + sb] = z * dinv;
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
        double uint_[PR], z = 0.0;
        if (swarp) {
            const bool ok = sval && sr < nr;
#pragma unroll
            for (int rp = 0; rp < PR; ++rp)
                uint_[rp] = (ok && rp > sr && rp < nr) ? lu[(size_t)(base + sr * W + next + rp) * Bp + sb] : 0.0;
            z = ok ? y[(size_t)(p0 + sr) * Bp + sb] : 0.0;
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
            for (int rp = PR - 1; rp > 0; --rp) {
                const double xv = __shfl_sync(0xffffffffu, z, rp, PR);          // U' has a unit diagonal
                if (sr < rp) z -= uint_[rp] * xv;
            }
            if (sval && sr < nr) {
                y[(size_t)(p0 + sr) * Bp + sb] = z;
                x[(size_t)net.perm[p0 + sr] * Bp + sb] = z;
            }
        }
        __syncthreads();
    }
}

}  // namespace kb2
