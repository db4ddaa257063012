// CUDA kernels of the kinetic-solve hot path (sm_100a).
//
// Layout: every ensemble array is [index][Bp] — species/reaction/slot-major, member-minor —
// so that consecutive members are consecutive in memory and every warp-level access is a
// full-sector coalesced FP64 load/store.  A CTA owns one *tile* of MB consecutive members
// (MB in {1,2,4,8,16,32}); its threads are laid out as (m = tid % MB, slot = tid / MB): the
// member index is the fast axis (coalescing), `slot` strides over species / reactions / LU
// slots.  All index tables are shared by every member, so control flow is uniform inside a
// tile and nothing diverges.  No atomics are used on the data path (gather CSR), results are
// deterministic run to run.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace kb2 {

constexpr double kR = 8.314462618;       // reference src/constants.jl:4
constexpr double kNA = 6.02214076e23;    // reference src/constants.jl:5
constexpr double kGamma = 0.25;          // Rodas4

// Rodas4 (Hairer & Wanner RODAS) in transformed K-form; verified against the order
// conditions in tests/test_rodas_tableau.py.
__constant__ double cA[6][6] = {
    {0, 0, 0, 0, 0, 0},
    {0.1544000000000000e+01, 0, 0, 0, 0, 0},
    {0.9466785280815826e+00, 0.2557011698983284e+00, 0, 0, 0, 0},
    {0.3314825187068521e+01, 0.2896124015972201e+01, 0.9986419139977817e+00, 0, 0, 0},
    {0.1221224509226641e+01, 0.6019134481288629e+01, 0.1253708332932087e+02, -0.6878860361058950e+00, 0, 0},
    {0.1221224509226641e+01, 0.6019134481288629e+01, 0.1253708332932087e+02, -0.6878860361058950e+00, 1.0, 0}};
__constant__ double cC[6][6] = {
    {0, 0, 0, 0, 0, 0},
    {-0.5668800000000000e+01, 0, 0, 0, 0, 0},
    {-0.2430093356833875e+01, -0.2063599157091915e+00, 0, 0, 0, 0},
    {-0.1073529058151375e+00, -0.9594562251023355e+01, -0.2047028614809616e+02, 0, 0, 0},
    {0.7496443313967647e+01, -0.1024680431464352e+02, -0.3399990352819905e+02, 0.1170890893206160e+02, 0, 0},
    {0.8083246795921522e+01, -0.7981132988064893e+01, -0.3152159432874371e+02, 0.1631930543123136e+02,
     -0.6058818238834054e+01, 0}};

struct DevNet {
    int S, R, nnzJ, nnzLU, max_rowlen;
    const int *rhs_ptr, *rhs_rxn, *rhs_coef;
    const int4 *rdesc;
    const int *jt_ptr, *jt_rxn, *jt_pack;
    const int *slot_src, *rowptr, *colidx, *diagpos, *perm;
    const unsigned *tgt_off;
    const int *tgt;
    // calculator
    int calc_mode;            // 0 Arrhenius, 1 rate table
    const double *A, *Ea, *n; // n may be null
    double k_max, t_mult;     // k_max NaN = uncapped
    const double *ktab, *kinit;
};

struct DevEns {
    int B, Bp;
    double *u, *ua, *rv, *y, *K[6], *k, *lu, *invd;
    // conditions
    int nstops;               // row length of the per-member stop tables
    const double *stop_t;     // [b*nstops + s]
    const int *stop_flags, *stop_ridx, *stop_cnt;   // [b*nstops + s], [b*nstops + s], [b]
    const double *Ttab;       // [b*nstops + s] or null
    const int *pkind;
    const double *pparams;    // [b*16]
    // outputs
    int Ns;
    double *out_u, *out_umax;
    int *status;
    long long *stats;
    // controls
    double t0, abstol, reltol, dtmin;
    long long maxiters;
    int ban_neg;
};

__device__ __forceinline__ double pw(double x, int e)
{
    double r = 1.0;
    for (; e > 0; --e) r *= x;
    return r;
}

// rate_j = k_j * prod_m u_m^nu_mj  (Catalyst mass action, combinatoric_ratelaws=false;
// reference src/solving/solve_utils.jl:318-334)
__device__ __forceinline__ double rate_of(const int4 d, const double *__restrict__ u, size_t Bp, int b, double kj)
{
    double r = kj;
    if (d.x >= 0) r *= pw(u[(size_t)d.x * Bp + b], d.w & 255);
    if (d.y >= 0) r *= pw(u[(size_t)d.y * Bp + b], (d.w >> 8) & 255);
    if (d.z >= 0) r *= pw(u[(size_t)d.z * Bp + b], (d.w >> 16) & 255);
    return r;
}

// d(rate_j)/du_l / nu_l for the reactant in descriptor slot s (nu_l is folded into the term coefficient)
__device__ __forceinline__ double drate_of(const int4 d, int s, const double *__restrict__ u, size_t Bp, int b, double kj)
{
    double r = kj;
    if (d.x >= 0) r *= pw(u[(size_t)d.x * Bp + b], (d.w & 255) - (s == 0));
    if (d.y >= 0) r *= pw(u[(size_t)d.y * Bp + b], ((d.w >> 8) & 255) - (s == 1));
    if (d.z >= 0) r *= pw(u[(size_t)d.z * Bp + b], ((d.w >> 16) & 255) - (s == 2));
    return r;
}

// k = A*T^n*exp(-Ea/(R*T))*N_A*t_mult, optional harmonic cap — operation order of
// reference src/solving/calculator.jl:223-232 (T^n is this build's extension, n = 0 by default).
__device__ __forceinline__ double arrhenius(const DevNet &net, int r, double T)
{
    double kr = net.A[r] * exp(-net.Ea[r] / (kR * T));
    if (net.n) kr *= pow(T, net.n[r]);
    kr = kr * kNA * net.t_mult;
    if (isnan(net.k_max)) return kr;
    return 1.0 / ((1.0 / net.k_max) + (1.0 / kr));
}

// Condition value X(t) of one member (reference src/conditions/*.jl; closed forms of the
// gradient profiles, which the reference integrates numerically: gradient_variable.jl:35-64).
__device__ inline double profile_eval(int kind, const double *__restrict__ p, double t)
{
    switch (kind) {
    case 0: case 1: return p[0];
    case 2: {   // LinearDirectProfile, direct_variable.jl:144-150
        double rate = p[0], Xs = p[1], Xe = p[2], te = p[3];
        if (t <= 0.0) return Xs;
        if (t <= te) return Xs + rate * t;
        return Xe;
    }
    case 3: {   // LinearGradientProfile: grad = rate for t <= t_end, gradient_variable.jl:165-170
        double rate = p[0], Xs = p[1], te = p[3];
        double tt = t < te ? t : te;
        if (tt < 0.0) tt = 0.0;
        return Xs + rate * tt;
    }
    case 4: {   // DoubleRampGradientProfile (plain :277-285, blended :287-299), integrated from 0
        double X = p[0], r1 = p[1], r2 = p[2], tb = p[7];
        if (t <= 0.0) return X;
        double ts[2] = {p[3], p[5]}, te[2] = {p[4], p[6]}, rr[2] = {r1, r2};
        for (int q = 0; q < 2; ++q) {
            double r = rr[q];
            if (tb > 0.0) {
                // ramp-up blend on [ts-tb, ts+tb): grad = r*(t-ts-tb)/(2tb) + r
                double a = ts[q] - tb, bnd = ts[q] + tb;
                double lo = fmax(a, 0.0), hi = fmin(bnd, t);
                if (hi > lo) X += 0.5 * (r / (2 * tb)) * ((hi - a) * (hi - a) - (lo - a) * (lo - a));
                lo = fmax(bnd, 0.0); hi = fmin(te[q] - tb, t);
                if (hi > lo) X += r * (hi - lo);
                a = te[q] - tb; bnd = te[q] + tb;
                lo = fmax(a, 0.0); hi = fmin(bnd, t);
                if (hi > lo) X += r * (hi - lo) - 0.5 * (r / (2 * tb)) * ((hi - a) * (hi - a) - (lo - a) * (lo - a));
            } else {
                double lo = fmax(ts[q], 0.0), hi = fmin(te[q], t);
                if (hi > lo) X += r * (hi - lo);
            }
        }
        return X;
    }
    default: return p[0];
    }
}

// ---------------------------------------------------------------------------------------------
// Tile primitives.  `b` = global member column, `slot`/`nslot` = this thread's stride lane.
// ---------------------------------------------------------------------------------------------
template <int MB>
struct Tile {
    int m, slot, nslot, b;
    size_t Bp;
    __device__ Tile(int tile, int Bp_) : Bp((size_t)Bp_)
    {
        m = threadIdx.x % MB;
        slot = threadIdx.x / MB;
        nslot = blockDim.x / MB;
        b = tile * MB + m;
    }
};

// K1: k[r][b] for the member's current condition value T (masked by `upd`)
template <int MB>
__device__ void tile_rates(const Tile<MB> &tl, const DevNet &net, double *__restrict__ k, double T, bool upd, int ridx)
{
    if (!upd) return;
    if (net.calc_mode == 0) {
        for (int r = tl.slot; r < net.R; r += tl.nslot) k[(size_t)r * tl.Bp + tl.b] = arrhenius(net, r, T);
    } else {
        const double *src = ridx < 0 ? net.kinit : net.ktab + (size_t)ridx * net.R;
        for (int r = tl.slot; r < net.R; r += tl.nslot) k[(size_t)r * tl.Bp + tl.b] = src[r];
    }
}

// K2: du_i = sum_e coef_e * rate_{j(e)} over the gather CSR of species i (ascending reaction
// order, no atomics).  out_i = du_i + sum_q cs[q]*Kq_i  (stage right-hand side fusion).
template <int MB>
__device__ void tile_rhs(const Tile<MB> &tl, const DevNet &net, const double *__restrict__ u,
                         const double *__restrict__ k, double *__restrict__ out, int nk,
                         double *const *Kq, const double *cs)
{
    for (int i = tl.slot; i < net.S; i += tl.nslot) {
        double acc = 0.0;
        const int e1 = net.rhs_ptr[i + 1];
        for (int e = net.rhs_ptr[i]; e < e1; ++e) {
            const int j = net.rhs_rxn[e];
            acc += (double)net.rhs_coef[e] * rate_of(net.rdesc[j], u, tl.Bp, tl.b, k[(size_t)j * tl.Bp + tl.b]);
        }
        const size_t o = (size_t)i * tl.Bp + tl.b;
        for (int q = 0; q < nk; ++q) acc += cs[q] * Kq[q][o];
        out[o] = acc;
    }
}

// K3: analytic Jacobian entry p = (i,l):  J_p = sum_t coef_t * k_j * d(prod)/du_l
template <int MB>
__device__ __forceinline__ double jac_entry(const Tile<MB> &tl, const DevNet &net, int p,
                                            const double *__restrict__ u, const double *__restrict__ k)
{
    double acc = 0.0;
    const int t1 = net.jt_ptr[p + 1];
    for (int t = net.jt_ptr[p]; t < t1; ++t) {
        const int j = net.jt_rxn[t], pk = net.jt_pack[t];
        acc += (double)(pk >> 2) * drate_of(net.rdesc[j], pk & 3, u, tl.Bp, tl.b, k[(size_t)j * tl.Bp + tl.b]);
    }
    return acc;
}

template <int MB>
__device__ void tile_jac_csc(const Tile<MB> &tl, const DevNet &net, const double *__restrict__ u,
                             const double *__restrict__ k, double *__restrict__ Jval)
{
    for (int p = tl.slot; p < net.nnzJ; p += tl.nslot) Jval[(size_t)p * tl.Bp + tl.b] = jac_entry(tl, net, p, u, k);
}

// W = I/(h*gamma) - J assembled straight into the L\U slots (fill slots zeroed)
template <int MB>
__device__ void tile_assemble_w(const Tile<MB> &tl, const DevNet &net, const double *__restrict__ u,
                                const double *__restrict__ k, double hg_inv, double *__restrict__ lu)
{
    for (int q = tl.slot; q < net.nnzLU; q += tl.nslot) {
        const int src = net.slot_src[q];
        double v = (src & 1) ? hg_inv : 0.0;
        const int p = (src >> 1) - 1;
        if (p >= 0) v -= jac_entry(tl, net, p, u, k);
        lu[(size_t)q * tl.Bp + tl.b] = v;
    }
}

// K4: in-place sparse LU (row-wise, no pivoting) over the shared symbolic factorisation.
// The target row lives in shared memory (w[offset][m]); for each pivot k in L(i,:) every thread
// forms l_ik = w_k/u_kk redundantly and the updates over U(k,:) are spread across the slots.
template <int MB>
__device__ void tile_lu(const Tile<MB> &tl, const DevNet &net, double *__restrict__ lu,
                        double *__restrict__ invd, double *__restrict__ w)
{
    for (int i = 0; i < net.S; ++i) {
        const int r0 = net.rowptr[i], r1 = net.rowptr[i + 1], dg = net.diagpos[i];
        if (dg == r0) {   // no L part: the row is already final
            if (tl.slot == 0) invd[(size_t)i * tl.Bp + tl.b] = 1.0 / lu[(size_t)dg * tl.Bp + tl.b];
            continue;
        }
        for (int o = tl.slot; o < r1 - r0; o += tl.nslot) w[o * MB + tl.m] = lu[(size_t)(r0 + o) * tl.Bp + tl.b];
        __syncthreads();
        for (int p = r0; p < dg; ++p) {
            const int kk = net.colidx[p];
            const double l = w[(p - r0) * MB + tl.m] * invd[(size_t)kk * tl.Bp + tl.b];
            const int ub = net.diagpos[kk] + 1, nu = net.rowptr[kk + 1] - ub;
            const int *__restrict__ tg = net.tgt + net.tgt_off[p];
            for (int e = tl.slot; e < nu; e += tl.nslot)
                w[tg[e] * MB + tl.m] -= l * lu[(size_t)(ub + e) * tl.Bp + tl.b];
            if (tl.slot == 0) lu[(size_t)p * tl.Bp + tl.b] = l;
            __syncthreads();
        }
        for (int o = dg - r0 + tl.slot; o < r1 - r0; o += tl.nslot) lu[(size_t)(r0 + o) * tl.Bp + tl.b] = w[o * MB + tl.m];
        if (tl.slot == 0) invd[(size_t)i * tl.Bp + tl.b] = 1.0 / w[(dg - r0) * MB + tl.m];
        __syncthreads();
    }
}

// sum over the slots of one member in a fixed order (deterministic): shuffle tree across the
// sub-lanes of each warp, then one pass over the per-warp partials; red has (nthreads/32)*MB doubles
template <int MB>
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int off = 16; off >= MB; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

template <int MB>
__device__ __forceinline__ double tile_sum(const Tile<MB> &tl, double v, double *red)
{
    v = warp_sum<MB>(v);
    const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if ((threadIdx.x & 31) < MB) red[warp * MB + tl.m] = v;
    __syncthreads();
    double s = 0.0;
    for (int q = 0; q < nw; ++q) s += red[q * MB + tl.m];
    __syncthreads();
    return s;
}

// K5: forward/back substitution  W x = rhs  (rhs, x in species order; y = permuted scratch)
template <int MB>
__device__ void tile_trisolve(const Tile<MB> &tl, const DevNet &net, const double *__restrict__ lu,
                              const double *__restrict__ invd, const double *__restrict__ rhs,
                              double *__restrict__ y, double *__restrict__ x, double *red)
{
    const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const bool wlead = (threadIdx.x & 31) < MB;
    // rows without an L part depend on nothing: do them all at once
    for (int i = threadIdx.x / MB; i < net.S; i += tl.nslot)
        if (net.diagpos[i] == net.rowptr[i]) y[(size_t)i * tl.Bp + tl.b] = rhs[(size_t)net.perm[i] * tl.Bp + tl.b];
    __syncthreads();
    for (int i = 0; i < net.S; ++i) {
        const int r0 = net.rowptr[i], dg = net.diagpos[i];
        if (dg - r0 > 0) {
            double part = 0.0;
            for (int p = r0 + tl.slot; p < dg; p += tl.nslot)
                part += lu[(size_t)p * tl.Bp + tl.b] * y[(size_t)net.colidx[p] * tl.Bp + tl.b];
            part = warp_sum<MB>(part);
            if (wlead) red[warp * MB + tl.m] = part;
            __syncthreads();
            if (tl.slot == 0) {
                double s = 0.0;
                for (int q = 0; q < nw; ++q) s += red[q * MB + tl.m];
                y[(size_t)i * tl.Bp + tl.b] = rhs[(size_t)net.perm[i] * tl.Bp + tl.b] - s;
            }
            __syncthreads();
        }
    }
    // rows without a U part (beyond the diagonal) only need scaling
    for (int i = threadIdx.x / MB; i < net.S; i += tl.nslot)
        if (net.rowptr[i + 1] - net.diagpos[i] == 1) {
            const double v = y[(size_t)i * tl.Bp + tl.b] * invd[(size_t)i * tl.Bp + tl.b];
            y[(size_t)i * tl.Bp + tl.b] = v;
            x[(size_t)net.perm[i] * tl.Bp + tl.b] = v;
        }
    __syncthreads();
    for (int i = net.S - 1; i >= 0; --i) {
        const int dg = net.diagpos[i], r1 = net.rowptr[i + 1];
        if (r1 - dg - 1 > 0) {
            double part = 0.0;
            for (int p = dg + 1 + tl.slot; p < r1; p += tl.nslot)
                part += lu[(size_t)p * tl.Bp + tl.b] * y[(size_t)net.colidx[p] * tl.Bp + tl.b];
            part = warp_sum<MB>(part);
            if (wlead) red[warp * MB + tl.m] = part;
            __syncthreads();
            if (tl.slot == 0) {
                double s = 0.0;
                for (int q = 0; q < nw; ++q) s += red[q * MB + tl.m];
                const double v = (y[(size_t)i * tl.Bp + tl.b] - s) * invd[(size_t)i * tl.Bp + tl.b];
                y[(size_t)i * tl.Bp + tl.b] = v;
                x[(size_t)net.perm[i] * tl.Bp + tl.b] = v;
            }
            __syncthreads();
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Stand-alone kernels (kernel-level C-ABI entry points, per-kernel roofline timing)
// ---------------------------------------------------------------------------------------------
template <int MB>
__global__ void k_rates(DevNet net, DevEns en, const double *T, int ntiles)
{
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        Tile<MB> tl(tile, en.Bp);
        tile_rates(tl, net, en.k, T[tl.b], true, -1);
    }
}

template <int MB>
__global__ void k_rhs(DevNet net, DevEns en, int ntiles)
{
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        Tile<MB> tl(tile, en.Bp);
        tile_rhs(tl, net, en.u, en.k, en.rv, 0, nullptr, nullptr);
    }
}

template <int MB>
__global__ void k_jac(DevNet net, DevEns en, double *Jval, int ntiles)
{
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        Tile<MB> tl(tile, en.Bp);
        tile_jac_csc(tl, net, en.u, en.k, Jval);
    }
}

template <int MB>
__global__ void k_factor(DevNet net, DevEns en, const double *hg_inv, int ntiles)
{
    extern __shared__ double smem[];
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        Tile<MB> tl(tile, en.Bp);
        tile_assemble_w(tl, net, en.u, en.k, hg_inv[tl.b], en.lu);
        __syncthreads();
        tile_lu(tl, net, en.lu, en.invd, smem);
        __syncthreads();
    }
}

template <int MB>
__global__ void k_trisolve(DevNet net, DevEns en, int ntiles)
{
    extern __shared__ double smem[];
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        Tile<MB> tl(tile, en.Bp);
        tile_trisolve(tl, net, en.lu, en.invd, en.rv, en.y, en.ua, smem);
        __syncthreads();
    }
}

__global__ void k_profile(int B, int nt, const int *kind, const double *params, const double *t, double *X)
{
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * nt) return;
    int b = idx / nt, s = idx % nt;
    X[idx] = profile_eval(kind[b], params + (size_t)b * 16, t[s]);
}

// [S][Bp] -> member-major [B][S] pack for the allgather (transpose fused into the pack)
__global__ void k_pack_bs(int S, int B, int Bp, const double *__restrict__ src, double *__restrict__ dst)
{
    __shared__ double tile[32][33];
    int b0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int i = i0 + r, b = b0 + threadIdx.x;
        tile[r][threadIdx.x] = (i < S && b < Bp) ? src[(size_t)i * Bp + b] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int b = b0 + r, i = i0 + threadIdx.x;
        if (b < B && i < S) dst[(size_t)b * S + i] = tile[threadIdx.x][r];
    }
}

// ---------------------------------------------------------------------------------------------
// The fused solve: one CTA integrates one tile of MB members from t0 to the last stop.
// Rodas4 with per-member adaptive h; accept/reject and stop handling are masked per member while
// the tile moves in lock-step.  Replaces init/solve!/reinit! of `pars.solver` and the
// PresetTimeCallback rate update (reference src/solving/methods.jl:655-714,
// src/solving/solve_utils.jl:376-450).
// ---------------------------------------------------------------------------------------------
template <int MB>
struct Ctl {
    double t[MB], h[MB], hs[MB], hold[MB], errold[MB], T[MB];
    long long iters[MB];
    int ns[MB], si[MB], isave[MB], status[MB], hit[MB], active[MB], rejlast[MB], firstacc[MB], accept[MB], upd[MB], ridx[MB], sav[MB];
    int nacc[MB], nrej[MB], nlu[MB], nrhs[MB];
};

enum { ST_RUNNING = -1 };

template <int MB>
__device__ void tile_process_stop(const Tile<MB> &tl, const DevNet &net, const DevEns &en, Ctl<MB> &c, bool at_start)
{
    // slot-0 thread of each member decides what its member does at this stop
    if (tl.slot == 0) {
        const int m = tl.m;
        c.upd[m] = 0; c.sav[m] = -1;
        const size_t sb = (size_t)tl.b * en.nstops;
        const bool due = at_start ? (c.status[m] == ST_RUNNING && c.si[m] < c.ns[m] && en.stop_t[sb + c.si[m]] <= en.t0)
                                  : (c.accept[m] && c.hit[m]);
        if (due) {
            const int s = c.si[m], fl = en.stop_flags[sb + s];
            if (fl & 1) {
                double T = en.Ttab ? en.Ttab[sb + s] : nan("");
                if (isnan(T) && net.calc_mode == 0) T = profile_eval(en.pkind[tl.b], en.pparams + (size_t)tl.b * 16, en.stop_t[sb + s]);
                c.T[m] = T; c.upd[m] = 1; c.ridx[m] = en.stop_ridx[sb + s];
            }
            if (fl & 2) c.sav[m] = c.isave[m]++;
            c.si[m] = s + 1;
            if (c.si[m] >= c.ns[m]) c.status[m] = 0;   // reached the end of tspan
        }
    }
    __syncthreads();
    const int m = tl.m;
    if (__syncthreads_or(c.upd[m])) tile_rates(tl, net, en.k, c.T[m], c.upd[m] != 0, c.ridx[m]);
    const int sv = c.sav[m];
    if (__syncthreads_or(sv >= 0)) {
        if (sv >= 0)
            for (int i = tl.slot; i < net.S; i += tl.nslot) {
                const double v = en.u[(size_t)i * tl.Bp + tl.b];
                en.out_u[((size_t)sv * net.S + i) * tl.Bp + tl.b] = v;
                double *mx = en.out_umax + (size_t)i * tl.Bp + tl.b;
                *mx = (sv == 0) ? v : fmax(*mx, v);
            }
    }
    __syncthreads();
}

// Starting step size (Hairer-Nørsett-Wanner II.4, order 4).  Called once at t0 (`initial`) and
// again after every discrete rate update, where the RHS jumps: members flagged in c.upd get
// h = min(h, estimate).  Uses rv, ua, y as scratch.
template <int MB>
__device__ void tile_hinit(const Tile<MB> &tl, const DevNet &net, const DevEns &en, Ctl<MB> &c, double *red, bool initial)
{
    const int m = tl.m, b = tl.b;
    const size_t Bp = tl.Bp;
    tile_rhs(tl, net, en.u, en.k, en.rv, 0, nullptr, nullptr);
    __syncthreads();
    double d0 = 0, d1 = 0;
    for (int i = tl.slot; i < net.S; i += tl.nslot) {
        const double ui = en.u[(size_t)i * Bp + b], fi = en.rv[(size_t)i * Bp + b];
        const double sc = en.abstol + en.reltol * fabs(ui);
        d0 += (ui / sc) * (ui / sc); d1 += (fi / sc) * (fi / sc);
    }
    d0 = sqrt(tile_sum(tl, d0, red) / net.S);
    d1 = sqrt(tile_sum(tl, d1, red) / net.S);
    const double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
    for (int i = tl.slot; i < net.S; i += tl.nslot)
        en.ua[(size_t)i * Bp + b] = en.u[(size_t)i * Bp + b] + h0 * en.rv[(size_t)i * Bp + b];
    __syncthreads();
    tile_rhs(tl, net, en.ua, en.k, en.y, 0, nullptr, nullptr);
    __syncthreads();
    double d2 = 0;
    for (int i = tl.slot; i < net.S; i += tl.nslot) {
        const double sc = en.abstol + en.reltol * fabs(en.u[(size_t)i * Bp + b]);
        const double q = (en.y[(size_t)i * Bp + b] - en.rv[(size_t)i * Bp + b]) / sc;
        d2 += q * q;
    }
    d2 = sqrt(tile_sum(tl, d2, red) / net.S) / h0;
    const double dm = fmax(d1, d2);
    const double h1 = (dm <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(0.01 / dm, 0.2);
    const double hn = fmin(100.0 * h0, h1);
    if (tl.slot == 0) {
        if (initial) { c.h[m] = hn; c.hold[m] = hn; c.nrhs[m] += 2; }
        else if (c.upd[m] && c.status[m] == ST_RUNNING) { c.h[m] = fmin(c.h[m], hn); c.nrhs[m] += 2; }
    }
    __syncthreads();
}

template <int MB>
__device__ void solve_tile(int tile, const DevNet &net, const DevEns &en, Ctl<MB> &c, double *w, double *red)
{
    Tile<MB> tl(tile, en.Bp);
    const int m = tl.m;
    const size_t Bp = tl.Bp;
    const int b = tl.b;
    if (tl.slot == 0) {
        c.t[m] = en.t0; c.si[m] = 0; c.isave[m] = 0; c.iters[m] = 0;
        c.ns[m] = en.stop_cnt[b];
        c.status[m] = (b < en.B && c.ns[m] > 0) ? ST_RUNNING : 0;
        c.nacc[m] = c.nrej[m] = c.nlu[m] = c.nrhs[m] = 0;
        c.rejlast[m] = 0; c.firstacc[m] = 1; c.accept[m] = 0; c.hit[m] = 0;
        c.errold[m] = 1.0;
        // initial conditions: static -> value, variable -> X_start (condition_set.jl:111-121);
        // both sit in the profile's X(0) for every supported kind
        c.T[m] = (net.calc_mode == 0) ? profile_eval(en.pkind[b], en.pparams + (size_t)b * 16, -1.0) : 0.0;
    }
    __syncthreads();
    tile_rates(tl, net, en.k, c.T[m], true, -1);     // k(initial conditions), methods.jl:668
    __syncthreads();
    tile_process_stop(tl, net, en, c, true);
    tile_hinit(tl, net, en, c, red, true);
    // ---- main loop ----
    for (;;) {
        if (tl.slot == 0) {
            int act = (c.status[m] == ST_RUNNING);
            double hs = 1.0;
            int hit = 0;
            if (act) {
                if (++c.iters[m] > en.maxiters) { c.status[m] = 1; act = 0; }
                else {
                    const double tstop = en.stop_t[(size_t)b * en.nstops + c.si[m]];
                    hs = c.h[m];
                    if (c.t[m] + 1.01 * hs >= tstop) { hs = tstop - c.t[m]; hit = 1; }
                    if (hs < en.dtmin && !hit) { c.status[m] = 2; act = 0; hs = 1.0; }
                }
            }
            c.active[m] = act; c.hs[m] = hs; c.hit[m] = hit; c.accept[m] = 0;
        }
        __syncthreads();
        if (!__syncthreads_or(c.active[m])) break;
        const double hs = c.hs[m];
        tile_assemble_w(tl, net, en.u, en.k, 1.0 / (hs * kGamma), en.lu);
        __syncthreads();
        tile_lu(tl, net, en.lu, en.invd, w);
        __syncthreads();
        for (int s = 0; s < 6; ++s) {
            const double *Us = en.u;
            if (s > 0) {
                for (int i = tl.slot; i < net.S; i += tl.nslot) {
                    const size_t o = (size_t)i * Bp + b;
                    double a = en.u[o];
                    for (int q = 0; q < s; ++q) a += cA[s][q] * en.K[q][o];
                    en.ua[o] = a;
                }
                __syncthreads();
                Us = en.ua;
            }
            double cs[5];
            for (int q = 0; q < s; ++q) cs[q] = cC[s][q] / hs;
            tile_rhs(tl, net, Us, en.k, en.rv, s, en.K, cs);
            __syncthreads();
            tile_trisolve(tl, net, en.lu, en.invd, en.rv, en.y, en.K[s], red);
            __syncthreads();
        }
        // error estimate = K6; new solution = ua + K6
        double e2 = 0.0;
        int neg = 0;
        for (int i = tl.slot; i < net.S; i += tl.nslot) {
            const size_t o = (size_t)i * Bp + b;
            const double k6 = en.K[5][o], un = en.ua[o] + k6;
            const double sc = en.abstol + en.reltol * fmax(fabs(en.u[o]), fabs(un));
            e2 += (k6 / sc) * (k6 / sc);
            neg |= (un < 0.0);
        }
        double err = sqrt(tile_sum(tl, e2, red) / net.S);
        const double nneg = en.ban_neg ? tile_sum(tl, (double)neg, red) : 0.0;
        if (!(err < INFINITY)) err = INFINITY;          // NaN/Inf (singular pivot, overflow) -> reject
        if (nneg > 0.0) err = fmax(err, 1e4);            // isoutofdomain, methods.jl:169-171
        if (tl.slot == 0 && c.active[m]) {
            double fac = (err < INFINITY) ? fmax(1.0 / 6.0, fmin(5.0, pow(err, 0.25) / 0.9)) : 5.0;
            double hnew = hs / fac;
            c.nlu[m]++; c.nrhs[m] += 6;
            if (err <= 1.0) {
                c.nacc[m]++;
                if (!c.firstacc[m]) {
                    double facgus = (c.hold[m] / hs) * pow(err * err / c.errold[m], 0.25) / 0.9;
                    facgus = fmax(1.0 / 6.0, fmin(5.0, facgus));
                    fac = fmax(fac, facgus);
                    hnew = hs / fac;
                }
                c.firstacc[m] = 0;
                c.hold[m] = hs; c.errold[m] = fmax(1e-2, err);
                if (c.rejlast[m]) hnew = fmin(hnew, hs);
                c.rejlast[m] = 0;
                c.accept[m] = 1;
                if (c.hit[m]) { c.t[m] = en.stop_t[(size_t)b * en.nstops + c.si[m]]; c.h[m] = fmax(hnew, c.h[m]); }
                else { c.t[m] += hs; c.h[m] = hnew; }
            } else {
                c.nrej[m]++; c.rejlast[m] = 1; c.h[m] = hnew;
                if (hnew < en.dtmin) c.status[m] = 2;
            }
        }
        __syncthreads();
        if (c.accept[m])
            for (int i = tl.slot; i < net.S; i += tl.nslot) {
                const size_t o = (size_t)i * Bp + b;
                en.u[o] = en.ua[o] + en.K[5][o];
            }
        __syncthreads();
        tile_process_stop(tl, net, en, c, false);
        if (__syncthreads_or(c.upd[m])) tile_hinit(tl, net, en, c, red, false);
    }
    if (tl.slot == 0 && b < en.B) {
        en.status[b] = c.status[m] == ST_RUNNING ? 5 : c.status[m];
        long long *st = en.stats + (size_t)b * 8;
        st[0] = c.nacc[m]; st[1] = c.nrej[m]; st[2] = c.nlu[m]; st[3] = c.nrhs[m];
        st[4] = c.isave[m]; st[5] = c.si[m]; st[6] = 0; st[7] = 0;
    }
    __syncthreads();
}

template <int MB>
__global__ void k_solve(DevNet net, DevEns en, int ntiles, int *tile_counter)
{
    extern __shared__ double smem[];
    __shared__ Ctl<MB> c;
    __shared__ int s_tile;
    double *w = smem;
    double *red = smem + (size_t)net.max_rowlen * MB;
    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1);
        __syncthreads();
        const int tile = s_tile;
        __syncthreads();
        if (tile >= ntiles) break;
        solve_tile<MB>(tile, net, en, c, w, red);
    }
}

}  // namespace kb2
