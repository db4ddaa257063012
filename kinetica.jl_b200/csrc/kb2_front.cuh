// Window LU (sm_100a): right-looking block LU with the active submatrix in shared memory.
//
// One CTA of 256 threads factorises MW members of a tile at a time (MW = MB, or a divisor of it so
// that two CTAs fit an SM).  The plan (kb2_front.cpp)
// walks the panels of the block plan in order; front P consists of the pivot block of panel P, the
// rows below it that have P as a source block (Lrows) and its U-part columns (Ucols).  Rows and
// columns own a slot of the window  Win[row slot][column slot][member]  from the first front that
// touches them until they are eliminated, so the Schur complement never leaves the SM:
//   init(P)   window entries that become live at P (row and column both active) receive their
//             original value from the compact Jacobian values: W = I/(h*gamma) - J is never
//             materialised in HBM; inactive entries are kept at zero
//   B(P)      warp 0: Crout LU of the nr x nr pivot block with shuffles (pivots on L', unit U').  Look-
//             ahead: B(P+1) runs on warp 0 DURING D(P), on a register copy of the block taken before
//             the update and brought up to date with the same strip values in the same order, so
//             the serial pivot chain is off the critical path (fronts whose pivot block has entries
//             that only become live at that front factorise it in a phase of their own)
//   C(P)      U strip  U'[:, j] = inv(L'_PP) w[:, j]   (one thread per column and member), and
//             L strip  L'[i, :] = x inv(U'_PP)          (one thread per row and member), in place
//             in the window and, once, to the panel storage in HBM that the sweeps read
//   D(P)      W[i, j] -= sum_k L'[i, k] U'[k, j]  over Lrows x Ucols, 8 x 4 register blocks
//   clear(P)  the slots of the pivot rows and columns are zeroed, then init(P+1) (behind one more
//             barrier in the rare front that re-uses a slot given up by the front before it)
// Three block barriers per front (four without look-ahead or with a re-used slot).  HBM traffic of a factorisation = the compact Jacobian values in,
// the factors out; nothing is read twice.  Every entry receives the same updates in the same order
// as in the left-looking block plan (tile_lu), so the two kernels agree bit for bit.
#pragma once
#include "kb2_kernels.cuh"

namespace kb2 {

struct DevFront {
    int NF, Wr, Wc, max_nl, max_nu;
    const int *f_info;      // FrontPlan::FREC ints per front
    const int *lists;
    const int2 *init;       // {window position, (J entry + 1) << 1 | is_diagonal}
};

constexpr int WL_NT = 256;          // threads per CTA
#ifndef KB2_WL_CB
#define KB2_WL_CB 4
#endif
constexpr int WL_CB = KB2_WL_CB;    // columns of a register block of the update (rows: 8)
constexpr int WL_PF = 4;            // original values of the next front a thread fetches ahead

__host__ __device__ inline int wl_list_cap(int max_nl, int max_nu) { return 16 + 2 * max_nu + 2 * max_nl; }
// mw = members per CTA
__host__ __device__ inline size_t wl_smem_bytes(int mw, int Wr, int Wc, int max_nl, int max_nu)
{
    return (size_t)8 * mw * ((size_t)Wr * Wc + 2 * (64 + 8)) + (size_t)8 * WL_PF * WL_NT + (size_t)3 * wl_list_cap(max_nl, max_nu) * 4;
}

__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}

// rank-nr update of one 8 x WL_CB register block of the window.  The pivot loop is unrolled by
// KU only: the strip values of KU pivots are in flight while the previous ones are applied (fully
// unrolled, the compiler hoists all 8 x 12 loads and runs out of registers).
#ifndef KB2_WL_KU
#define KB2_WL_KU 2
#endif
constexpr int WL_KU = KB2_WL_KU;
template <int MW>
__device__ __forceinline__ void wl_update_block(double *WinM, const int *prs, const int *pcs, const int (&ro)[8], const int (&co)[WL_CB],
                                                const bool (&rok)[8], const bool (&cok)[WL_CB], int nr, int Wc)
{
    // WinM = Win + member; ro[i] = row slot * Wc * MW, co[c] = column slot * MW
    double acc[8][WL_CB];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < WL_CB; ++c) acc[i][c] = WinM[ro[i] + co[c]];
#pragma unroll WL_KU
    for (int k = 0; k < nr; ++k) {
        const int pc = pcs[k] * MW, pr = prs[k] * Wc * MW;
        double l[8], u[WL_CB];
#pragma unroll
        for (int i = 0; i < 8; ++i) l[i] = WinM[ro[i] + pc];
#pragma unroll
        for (int c = 0; c < WL_CB; ++c) u[c] = WinM[pr + co[c]];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int c = 0; c < WL_CB; ++c) acc[i][c] -= l[i] * u[c];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < WL_CB; ++c)
            if (rok[i] && cok[c]) WinM[ro[i] + co[c]] = acc[i][c];
}

// Crout LU of the nr x nr pivot block held one row per lane (lane = row * MW + member, D[j] = entry
// (row, j)); pivot rows are broadcast with shuffles.  Writes L'_PP \ U'_PP and the pivot reciprocals
// to shared memory (for the strips) and to the panel storage / invd in HBM.  All 32 lanes call.
template <int MB, int MW>
__device__ __forceinline__ void wl_pivot_block(double (&D)[8], int lane, int nr, double *Dl, double *dinv, double *invd_g, double *lu_diag)
{
    const int ln = lane / MW, mw = lane % MW;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (j < nr) {
            const double piv = __shfl_sync(FULL, D[j], j * MW + mw);
            const double inv = 1.0 / piv;
            if (ln == j) { invd_g[(size_t)j * MB] = inv; dinv[j * MW + mw] = inv; }
#pragma unroll
            for (int i = j + 1; i < 8; ++i) {
                const double uji = __shfl_sync(FULL, D[i], j * MW + mw) * inv;
                if (ln == j) D[i] = uji;
                else if (ln > j && ln < 8) D[i] -= D[j] * uji;
            }
        }
    }
    if (ln < nr) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j < nr) {
                Dl[(ln * 8 + j) * MW + mw] = D[j];
                lu_diag[((size_t)j * nr + ln) * MB] = D[j];
            }
    }
}

// One CTA factorises MW members of a tile (MW divides MB: a tile may be shared by MB/MW CTAs).
// Window layout [row slot][column slot][MW]; HBM arrays keep the tile layout [index][MB].
// hg: per-member 1/(h*gamma) (kernel-level entry point), or null: taken from the control state,
// work items without a running member are skipped.
template <int MB, int MW>
__global__ void __launch_bounds__(WL_NT, 1) k_lu_window(DevNet net, DevPlan pl, DevFront fr, DevEns en, const double *hg, int ntiles, int stagger_ns)
{
    extern __shared__ double smem[];
    constexpr int NX = WL_NT / MW, FREC = 12, NPART = MB / MW;
    constexpr int NXL = (WL_NT - 32) / MW;                // x-threads outside warp 0
    const int Wc = fr.Wc, Wr = fr.Wr;
    double *Win = smem;                                   // [Wr*Wc][MW]
    double *Dl2 = Win + (size_t)Wr * Wc * MW;             // [2][8][8][MW]: L'_PP (lower + pivots) \ U'_PP (strict upper) of this front and the next
    double *dinv2 = Dl2 + 2 * 64 * MW;                    // [2][8][MW]
    double *stage = dinv2 + 2 * 8 * MW;                   // [WL_PF][WL_NT]: Jacobian values of the next front's new entries
    const int LCAP = wl_list_cap(fr.max_nl, fr.max_nu);
    int *lst = reinterpret_cast<int *>(stage + WL_PF * WL_NT);   // [3][LCAP]: lists of this front and the next two
    const int tid = threadIdx.x, mw = tid % MW, x = tid / MW, lane = tid & 31, warp = tid >> 5;
    const int xl = x - 32 / MW;                           // index among the x-threads outside warp 0 (negative in warp 0)
    if (stagger_ns > 0 && blockIdx.x >= gridDim.x / 2) __nanosleep(stagger_ns);
    for (int work = blockIdx.x; work < ntiles * NPART; work += gridDim.x) {
        const int tile = work / NPART, m = (work % NPART) * MW + mw;
        const int b = tile * MB + m;
        double hgi;
        if (hg) hgi = hg[b];
        else {
            const Ctl *c = en.ctl + b;
            if (!__syncthreads_or(c->active)) continue;
            hgi = 1.0 / (c->hs * kGamma);
        }
        const double *jv = en.jv + (size_t)tile * net.nnzJ * MB + m;
        double *lu = en.lu + (size_t)tile * pl.padded * MB + m;
        double *invd = en.invd + (size_t)tile * net.S * MB + m;
        double *WinM = Win + mw;
        auto init_value = [&](int src) {
            double v = (src >> 1) ? -jv[(size_t)((src >> 1) - 1) * MB] : 0.0;
            if (src & 1) v += hgi;
            return v;
        };
        // pivot block of front Q straight from the window (warp 0; all its lanes call)
        auto pivot_block_from_window = [&](const int *f, const int *L, int buf) {
            const int nr = f[0], ln = lane / MW;
            const int *prs = L, *pcs = L + 8;
            double D[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) D[j] = (ln < nr && j < nr) ? WinM[(prs[ln] * Wc + pcs[j]) * MW] : 0.0;
            wl_pivot_block<MB, MW>(D, lane, nr, Dl2 + buf * 64 * MW, dinv2 + buf * 8 * MW, invd + (size_t)f[1] * MB,
                                   lu + ((size_t)f[4] + (size_t)f[5] * nr) * MB);
        };
        // ---- prologue: clear the window (inactive entries are zero from here on), lists of fronts
        // 0 and 1, original values of front 0, its pivot block ----
        {
            const int n = Wr * Wc * MW;
            for (int i = tid; i < n; i += WL_NT) Win[i] = 0.0;
            for (int q = 0; q < 2 && q < fr.NF; ++q) {
                const int *f = fr.f_info + q * FREC;
                const int len = 16 + 2 * f[2] + 2 * f[3];
                for (int i = tid; i < len; i += WL_NT) lst[q * LCAP + i] = fr.lists[f[6] + i];
            }
            __syncthreads();
            const int *f = fr.f_info;
            for (int e = x; e < f[8]; e += NX) {
                const int2 ent = fr.init[f[7] + e];
                WinM[ent.x * MW] = init_value(ent.y);
            }
            __syncthreads();
            if (warp == 0) pivot_block_from_window(f, lst, 0);
        }
        __syncthreads();
        for (int P = 0; P < fr.NF; ++P) {
            const int *f = fr.f_info + (size_t)P * FREC;
            const int nr = f[0], nu = f[2], nl = f[3], base = f[4], next = f[5];
            const int *L0 = lst + (P % 3) * LCAP, *L1 = lst + ((P + 1) % 3) * LCAP;
            const int *prs = L0, *pcs = L0 + 8, *ucs = L0 + 16, *ujj = ucs + nu, *lrs = ujj + nu, *lgs = lrs + nl;
            const double *Dl = Dl2 + (P & 1) * 64 * MW, *dinv = dinv2 + (P & 1) * 8 * MW;
            const bool has_next = P + 1 < fr.NF;
            const int *fn = f + (has_next ? FREC : 0);
            const int ni = has_next ? fn[8] : 0, ioff = fn[7], hot = has_next ? fn[9] : 0;
            const bool la = has_next && fn[10];           // the next pivot block is factorised while this front updates
            // ---- look ahead: the lists of front P+2 go to the third buffer (cp.async); the original
            // values of front P+1 are fetched by the warps that have no pivot block to prepare ----
            if (P + 2 < fr.NF) {
                const int *f2 = f + 2 * FREC;
                int *L2 = lst + ((P + 2) % 3) * LCAP;
                const int len = 16 + 2 * f2[2] + 2 * f2[3];
                for (int i = tid; i < len; i += WL_NT) cp_async4(L2 + i, fr.lists + f2[6] + i);
            }
            int ppos[WL_PF], psrc[WL_PF];
#pragma unroll
            for (int q = 0; q < WL_PF; ++q) { ppos[q] = -1; psrc[q] = 0; }
            double Dn[8];                                 // warp 0: row `ln` of the next pivot block, before this front's update
#pragma unroll
            for (int j = 0; j < 8; ++j) Dn[j] = 0.0;
            if (warp == 0) {
                cp_async_commit();
                if (la) {
                    const int ln = lane / MW, nr1 = fn[0];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (ln < nr1 && j < nr1) Dn[j] = WinM[(L1[ln] * Wc + L1[8 + j]) * MW];
                }
            } else {
                // the next front's new entries: positions and sources now (in flight during the strips) ...
#pragma unroll
                for (int q = 0; q < WL_PF; ++q) {
                    const int e = xl + q * NXL;
                    int2 ent = make_int2(-1, 0);
                    if (e < ni) ent = fr.init[ioff + e];
                    ppos[q] = ent.x; psrc[q] = ent.y;
                }
            }
            // ---- C: strips, in place (the pivot block of this front is in Dl / dinv).  A U column and
            // an L row are the same substitution  v[r] -= sum_{a<r} c(r,a) v[a]  with c = L'_PP[r][a]
            // (then scaled by 1/pivot) or c = U'_PP[a][r]; a thread runs two of them interleaved (two
            // independent dependency chains: the phase is latency-bound with eight warps per SM). ----
            {
                int offU[8], offL[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) { offU[r] = r < nr ? prs[r] * Wc * MW : 0; offL[r] = r < nr ? pcs[r] * MW : 0; }
                const int ntask = nu + nl;
                for (int t = x; t < ntask; t += 2 * NX) {
                    int tt[2] = {t, t + NX < ntask ? t + NX : t};
                    const bool two = t + NX < ntask;
                    bool isu[2];
                    int wb[2], gst[2];
                    double *g[2], v[2][8];
#pragma unroll
                    for (int z = 0; z < 2; ++z) {
                        isu[z] = tt[z] < nu;
                        if (isu[z]) {
                            wb[z] = ucs[tt[z]] * MW;
                            g[z] = lu + ((size_t)base + (size_t)(next + nr + ujj[tt[z]]) * nr) * MB;
                            gst[z] = MB;
                        } else {
                            const int ii = tt[z] - nu, gs = lgs[ii];
                            wb[z] = lrs[ii] * Wc * MW;
                            g[z] = lu + (size_t)(gs & 0x0fffffff) * MB;
                            gst[z] = ((gs >> 28) + 1) * MB;
                        }
#pragma unroll
                        for (int r = 0; r < 8; ++r) v[z][r] = r < nr ? WinM[wb[z] + (isu[z] ? offU[r] : offL[r])] : 0.0;
                    }
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        if (r < nr) {
#pragma unroll
                            for (int a = 0; a < r; ++a) {
#pragma unroll
                                for (int z = 0; z < 2; ++z) v[z][r] -= Dl[((isu[z] ? r * 8 + a : a * 8 + r)) * MW + mw] * v[z][a];
                            }
#pragma unroll
                            for (int z = 0; z < 2; ++z) v[z][r] *= isu[z] ? dinv[r * MW + mw] : 1.0;
                        }
                    }
#pragma unroll
                    for (int z = 0; z < 2; ++z) {
                        if (z == 0 || two) {
#pragma unroll
                            for (int r = 0; r < 8; ++r)
                                if (r < nr) { WinM[wb[z] + (isu[z] ? offU[r] : offL[r])] = v[z][r]; g[z][(size_t)r * gst[z]] = v[z][r]; }
                        }
                    }
                }
            }
            if (warp != 0) {
                // ... and their Jacobian values straight into shared memory (cp.async: no register waits
                // for them), in flight during the update and consumed after it
#pragma unroll
                for (int q = 0; q < WL_PF; ++q)
                    if (ppos[q] >= 0 && (psrc[q] >> 1)) cp_async8(stage + q * WL_NT + tid, jv + (size_t)((psrc[q] >> 1) - 1) * MB);
                cp_async_commit();
            }
            __syncthreads();
            // ---- D: rank-nr update of Lrows x Ucols.  With look-ahead, warp 0 instead brings its copy
            // of the next pivot block up to date (same strip values, same order) and factorises it. ----
            if (la && warp == 0) {
                const int ln = lane / MW, nr1 = fn[0];
                if (ln < nr1) {
                    const double *lrow = WinM + L1[ln] * Wc * MW;
                    for (int k = 0; k < nr; ++k) {
                        const double l = lrow[pcs[k] * MW];
                        const double *urow = WinM + prs[k] * Wc * MW;
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (j < nr1) Dn[j] -= l * urow[L1[8 + j] * MW];
                    }
                }
                wl_pivot_block<MB, MW>(Dn, lane, nr1, Dl2 + ((P + 1) & 1) * 64 * MW, dinv2 + ((P + 1) & 1) * 8 * MW,
                                       invd + (size_t)fn[1] * MB, lu + ((size_t)fn[4] + (size_t)fn[5] * nr1) * MB);
            } else if (nu > 0 && nl > 0) {
                const int ncb = (nu + WL_CB - 1) / WL_CB, nrb = (nl + 7) / 8;
                const int t0 = la ? xl : x, tstep = la ? NXL : NX;
                for (int t = t0; t < nrb * ncb; t += tstep) {
                    const int rb = t / ncb, cb = t - rb * ncb;
                    int ro[8], co[WL_CB];
                    bool rok[8], cok[WL_CB];
#pragma unroll
                    for (int i = 0; i < 8; ++i) { rok[i] = rb * 8 + i < nl; ro[i] = lrs[min(rb * 8 + i, nl - 1)] * Wc * MW; }
#pragma unroll
                    for (int c = 0; c < WL_CB; ++c) { cok[c] = cb + c * ncb < nu; co[c] = ucs[min(cb + c * ncb, nu - 1)] * MW; }
                    wl_update_block<MW>(WinM, prs, pcs, ro, co, rok, cok, nr, Wc);
                }
            }
            __syncthreads();
            // ---- the pivot rows and columns of P are dead: clear their slots (inactive entries stay
            // zero), then the next front's original values (which may land in those slots) ----
            // (warp r clears pivot r: its row slot is contiguous, its column slot strided)
            if (warp < nr) {
                const int r = warp;
                double *row = Win + prs[r] * Wc * MW;
                for (int i = lane; i < Wc * MW; i += 32) row[i] = 0.0;
                double *col = Win + pcs[r] * MW + lane % MW;
                for (int z = lane / MW; z < Wr; z += 32 / MW) col[z * Wc * MW] = 0.0;
            }
            cp_async_wait_all();
            if (hot) __syncthreads();
#pragma unroll
            for (int q = 0; q < WL_PF; ++q)
                if (ppos[q] >= 0) {
                    double v = (psrc[q] >> 1) ? -stage[q * WL_NT + tid] : 0.0;
                    if (psrc[q] & 1) v += hgi;
                    WinM[ppos[q] * MW] = v;
                }
            if (xl >= 0)       // fronts with more new values than the staged ones (the first front of a block): four at a time
                for (int e0 = xl + WL_PF * NXL; e0 < ni; e0 += 4 * NXL) {
                    int2 ent[4];
                    double v[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) ent[q] = e0 + q * NXL < ni ? fr.init[ioff + e0 + q * NXL] : make_int2(-1, 0);
#pragma unroll
                    for (int q = 0; q < 4; ++q) v[q] = ent[q].x >= 0 ? init_value(ent[q].y) : 0.0;
#pragma unroll
                    for (int q = 0; q < 4; ++q) if (ent[q].x >= 0) WinM[ent[q].x * MW] = v[q];
                }
            __syncthreads();
            // a pivot block with entries that are new at its own front waits for them
            if (has_next && !la) {
                if (warp == 0) pivot_block_from_window(fn, L1, (P + 1) & 1);
                __syncthreads();
            }
        }
    }
}

// compact Jacobian values of every tile, CSC order, tile-major [tile][nnzJ][MB] (input of the window
// LU); NW warps of one CTA share a tile like in the right-hand side
template <int MB, int NW>
__global__ void __launch_bounds__(NW * 32, KB2_RHS_MINB) k_step_jac(DevNet net, DevPlan pl, DevEns en, int ntiles, int use_ctl)
{
    extern __shared__ double smem[];
    BulkChan ch; ch.bar = 0; ch.par = nullptr;
    const int w = threadIdx.x >> 5;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        WTile<MB> tl(tile, net, pl, en, ch);
        if (use_ctl && !__syncthreads_or(en.ctl[tl.b].active)) continue;
        if (use_ctl && en.continuous) {
            // continuous rate updates: the Jacobian belongs to k(T(t)) at the start of the step (the
            // stages of the previous attempt left k at their own times)
            const Ctl *c = en.ctl + tl.b;
            tile_rates<MB, NW>(tl, net, profile_eval(en.pkind[tl.b], en.pparams + (size_t)tl.b * 16, c->t), c->active != 0, -1, w);
        }
        const double *u = tl.u;
        if (en.u_smem) {
            stage_vector<MB, NW>(tl, smem, u, net.S * MB);
            u = smem;
        }
        tile_jac_csc<MB, NW>(tl, net, u, en.jv + (size_t)tile * net.nnzJ * MB, w);
    }
}

}  // namespace kb2
