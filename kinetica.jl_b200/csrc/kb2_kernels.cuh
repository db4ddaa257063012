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

// x^e for the small non-negative integer stoichiometries of mass action; exponents 0, 1 and 2
// (uni/bimolecular steps, max_molecularity = 2 in the reference, network.jl:250) are branch-free selects
__device__ __forceinline__ double pw(double x, int e)
{
    if (e <= 2) return e == 1 ? x : (e == 2 ? x * x : 1.0);
    double r = x * x * x;
    for (e -= 3; e > 0; --e) r *= x;
    return r;
}

// rate_j = k_j * prod_m u_m^nu_mj  (Catalyst mass action, combinatoric_ratelaws=false;
// reference src/solving/solve_utils.jl:318-334)
__device__ __forceinline__ double rate_of(const int4 d, const double *u, size_t Bp, int b, double kj)
{
    double r = kj;
    if (d.x >= 0) r *= pw(u[(size_t)d.x * Bp + b], d.w & 255);
    if (d.y >= 0) r *= pw(u[(size_t)d.y * Bp + b], (d.w >> 8) & 255);
    if (d.z >= 0) r *= pw(u[(size_t)d.z * Bp + b], (d.w >> 16) & 255);
    return r;
}

// d(rate_j)/du_l / nu_l for the reactant in descriptor slot s (nu_l is folded into the term coefficient)
__device__ __forceinline__ double drate_of(const int4 d, int s, const double *u, size_t Bp, int b, double kj)
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
__device__ inline double profile_eval(int kind, const double *p, double t)
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
__device__ void tile_rates(const Tile<MB> &tl, const DevNet &net, double *k, double T, bool upd, int ridx)
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
__device__ void tile_rhs(const Tile<MB> &tl, const DevNet &net, const double *u,
                         const double *k, double *out, int nk,
                         double *const *Kq, const double *cs)
{
    // four species per thread at a time: their gather chains (index -> descriptor -> k, u) are
    // independent, which gives the memory system four times as many loads in flight
    constexpr int U = 4;
    for (int i0 = tl.slot; i0 < net.S; i0 += U * tl.nslot) {
        int e[U], e1[U];
        double acc[U];
        int len = 0;
#pragma unroll
        for (int v = 0; v < U; ++v) {
            const int i = i0 + v * tl.nslot;
            e[v] = i < net.S ? net.rhs_ptr[i] : 0;
            e1[v] = i < net.S ? net.rhs_ptr[i + 1] : 0;
            acc[v] = 0.0;
            len = max(len, e1[v] - e[v]);
        }
        for (int t = 0; t < len; ++t) {
#pragma unroll
            for (int v = 0; v < U; ++v) {
                if (e[v] + t < e1[v]) {
                    const int j = net.rhs_rxn[e[v] + t];
                    acc[v] += (double)net.rhs_coef[e[v] + t] * rate_of(net.rdesc[j], u, tl.Bp, tl.b, k[(size_t)j * tl.Bp + tl.b]);
                }
            }
        }
#pragma unroll
        for (int v = 0; v < U; ++v) {
            const int i = i0 + v * tl.nslot;
            if (i < net.S) {
                const size_t o = (size_t)i * tl.Bp + tl.b;
                double a = acc[v];
                for (int q = 0; q < nk; ++q) a += cs[q] * Kq[q][o];
                out[o] = a;
            }
        }
    }
}

// K3: analytic Jacobian entry p = (i,l):  J_p = sum_t coef_t * k_j * d(prod)/du_l
template <int MB>
__device__ __forceinline__ double jac_entry(const Tile<MB> &tl, const DevNet &net, int p,
                                            const double *u, const double *k)
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
__device__ void tile_jac_csc(const Tile<MB> &tl, const DevNet &net, const double *u,
                             const double *k, double *Jval)
{
    for (int p = tl.slot; p < net.nnzJ; p += tl.nslot) Jval[(size_t)p * tl.Bp + tl.b] = jac_entry(tl, net, p, u, k);
}

// W = I/(h*gamma) - J assembled straight into the L\U slots (fill slots zeroed); four slots per
// thread in flight
template <int MB>
__device__ void tile_assemble_w(const Tile<MB> &tl, const DevNet &net, const double *u,
                                const double *k, double hg_inv, double *lu)
{
    constexpr int U = 4;
    for (int q0 = tl.slot; q0 < net.nnzLU; q0 += U * tl.nslot) {
        int t[U], t1[U];
        double v[U];
        int len = 0;
#pragma unroll
        for (int x = 0; x < U; ++x) {
            const int q = q0 + x * tl.nslot;
            const int src = q < net.nnzLU ? net.slot_src[q] : 0;
            v[x] = (src & 1) ? hg_inv : 0.0;
            const int p = (src >> 1) - 1;
            t[x] = p >= 0 ? net.jt_ptr[p] : 0;
            t1[x] = p >= 0 ? net.jt_ptr[p + 1] : 0;
            len = max(len, t1[x] - t[x]);
        }
        for (int z = 0; z < len; ++z) {
#pragma unroll
            for (int x = 0; x < U; ++x) {
                if (t[x] + z < t1[x]) {
                    const int j = net.jt_rxn[t[x] + z], pk = net.jt_pack[t[x] + z];
                    v[x] -= (double)(pk >> 2) * drate_of(net.rdesc[j], pk & 3, u, tl.Bp, tl.b, k[(size_t)j * tl.Bp + tl.b]);
                }
            }
        }
#pragma unroll
        for (int x = 0; x < U; ++x) {
            const int q = q0 + x * tl.nslot;
            if (q < net.nnzLU) lu[(size_t)q * tl.Bp + tl.b] = v[x];
        }
    }
}

// K4: in-place sparse LU (row-wise, no pivoting) over the shared symbolic factorisation.
// The target row lives in shared memory (w[offset][m]); for each pivot k in L(i,:) every thread
// forms l_ik = w_k/u_kk redundantly and the updates over U(k,:) are spread across the slots.
template <int MB>
__device__ void tile_lu(const Tile<MB> &tl, const DevNet &net, double *lu,
                        double *invd, double *w)
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
__device__ void tile_trisolve(const Tile<MB> &tl, const DevNet &net, const double *lu,
                              const double *invd, const double *rhs,
                              double *y, double *x, double *red)
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

}  // namespace kb2
