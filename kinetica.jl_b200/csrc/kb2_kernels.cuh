// CUDA device code of the kinetic-solve hot path (sm_100a): data structures and tile primitives.
//
// Execution model: the ensemble is cut into TILES of MB consecutive members (MB in {1,2,4}); a
// tile is worked on by one warp (lane = ln * MB + m with m the member inside the tile and ln one of
// LN = 32/MB work lanes of that member) or, in the streaming phases, by NW warps of one CTA that
// deal the reactions / rows / ELL groups among themselves (virtual lane = w * LN + ln).  The solve
// is a sequence of phase kernels per attempted step (kb2_solve.cuh); the factorisation has its
// own CTA-wide kernel (kb2_front.cuh).
// Layout: every per-member array is tile-major [tile][index][MB] — species / reaction / LU-slot
// major, member minor — so the threads that work on the same index touch one MB*8-byte segment
// (a full 32-byte sector for MB = 4) and consecutive indices are consecutive in memory.
// All index tables are shared by every member, control flow is uniform inside a warp.  No atomics
// on the data path (gather tables), results are deterministic run to run.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace kb2 {

constexpr double kR = 8.314462618;       // reference src/constants.jl:4
constexpr double kNA = 6.02214076e23;    // reference src/constants.jl:5
constexpr double kGamma = 0.25;          // Rodas4

// Rodas4 (Hairer & Wanner RODAS) in transformed K-form; verified against the order
// conditions in tests/test_oracle_golden.py::test_rodas4_tableau_order_conditions.
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

// non-autonomous form (continuous rate updates): stage times t + cT[i]*h and the time-derivative
// weights h*cD[i]*df/dt;  c = A d (row sums), checked in tests/test_oracle_golden.py
__constant__ double cT[6] = {0.0, 0.386, 0.21, 0.63, 1.0, 1.0};
__constant__ double cD[6] = {0.25, -0.1043, 0.1035, -0.3620000000000023e-01, 0.0, 0.0};

struct DevNet {
    int S, R, nnzJ;
    const int *rhs_ptr, *rhs_rxn, *rhs_coef;   // gather CSR by species; rhs_rxn holds positions in the rate table
    const int *rate_pos, *drate_pos;           // first-touch layouts: rate of reaction j at rate_pos[j], derivative (j, s) at drate_pos[j*jslots + s]
    const int4 *rdesc;
    const int *jt_ptr, *jt_pk;             // Jacobian terms by entry (CSC order): coef << 24 | (reaction*jslots + reactant slot)
    const int *jell_ptr, *jell;            // their sliced ELL (entries of j_order after the first j_nlong)
    int jell_ngroups, jslots;
    const int *rhs_order, *j_order;        // work orders, longest first; the first n_long are split across lanes
    int rhs_nlong, j_nlong;
    const int *ell_ptr, *ell;              // sliced ELL of the one-per-lane RHS rows: ell[ell_ptr[g] + t*64 + slot] = coef << 24 | reaction
    int ell_ngroups;
    const int *jslot, *diag_slot, *perm;   // storage slot of every Jacobian entry / pivot; perm[a] = species at pivot a
    // calculator
    int calc_mode;            // 0 Arrhenius, 1 rate table
    const double *A, *Ea, *n; // n may be null
    double k_max, t_mult;     // k_max NaN = uncapped
    const double *ktab, *kinit;
};

// Block plan of the LU / triangular solves (kb2_panel.cpp), shared by all members.
struct DevPlan {
    int npanels, nunits, padded;
    const int *p_row0, *p_nrows, *p_width, *p_next, *p_base, *p_cptr, *cols;
    const int4 *u_info;       // 3 per unit: {panel, x0, x1, task0}, {ntask, dmode, first U'_QQ slot, its nq}, {nr, next, row0, base}
    const int4 *t_info;       // 3 per task: {Q, lpos | inchunk << 30, ntargets, map0}, {nq, base of Q, U'_QQ slot, next U'_QQ slot},
                              //             {next nq, first slot of the target span, slots of the span, 0}
    const int *map;           // (column position in Q) | (column position in the chunk << 16)
};

// Control state of one member between the phase kernels of a step (kb2_solve.cuh).
struct Ctl {
    double t, h, hs, hold, errold, T, hfirst;
    double atol, rtol;                 // tolerances in force (tightened by the chunk retry)
    double tchunk, Tchunk;             // start of the current chunk: time, condition value of its rate constants
    long long iters;
    int ns, si, isave, status, hit, active, rejlast, firstacc, accept, upd, ridx, sav, fresh;
    int nacc, nrej, nlu, nrhs;
    int si_chunk, isave_chunk, ridx_chunk, retries, chunk, nretry;   // chunk = a chunk boundary was just passed
};

// Ensemble state.  Every per-member array is TILE-MAJOR: [tile][index][MB] with MB members per
// tile (species / reaction / LU-slot major, member minor inside the tile), so one warp owns one
// contiguous block per array and every access of the warp covers whole MB*8-byte segments.
struct DevEns {
    int B, Bp, MB;
    int u_smem;               // 1: a tile's state vector (S*MB doubles) fits the warp's shared memory and is staged there for the gathers
    double *u, *ua, *rv, *y, *K[6], *k, *rate, *drate, *lu, *invd;
    double *jv;               // compact Jacobian values, CSC order: [tile][nnzJ][MB]
    double *uc;               // state at the start of the current chunk (chunk retry), like u
    int continuous;           // continuous rate updates (methods.jl:363-458): k(T(t)) at every stage time, df/dt term
    double *kdot, *ft;        // dk/dt [tile][R][MB] and df/dt [tile][S][MB] of the current step (continuous mode)
    // conditions
    int nstops;               // row length of the per-member stop tables
    const double *stop_t;     // [b*nstops + s]
    const int *stop_flags, *stop_ridx, *stop_cnt;   // [b*nstops + s], [b*nstops + s], [b]
    const double *Ttab;       // [b*nstops + s] or null
    const int *pkind;
    const double *pparams;    // [b*16]
    // outputs
    int Ns;
    double *out_u, *out_umax; // [tile][s][i][MB], [tile][i][MB]
    int *status;
    long long *stats;
    Ctl *ctl;                 // [Bp]
    int *flags;               // [KB2_FLAG_SLOTS] members still running at the start of a round (host loop control)
    // controls
    double t0, abstol, reltol, dtmin;
    long long maxiters;
    int ban_neg;
    int chunk_retry;          // a failed chunk is repeated from its start with tolerances x0.1 (adaptive_solve! per chunk)
    int update_tols;          // tightened tolerances stay in force for the following chunks
};

// x^e for the small non-negative integer stoichiometries of mass action; exponents 0, 1 and 2
// (uni/bimolecular steps, max_molecularity = 2 in the reference, network.jl:250) are branch-free selects
__device__ __forceinline__ double pw(double x, int e)
{
    const double x2 = x * x;
    double r = e == 1 ? x : (e == 2 ? x2 : (e == 3 ? x2 * x : 1.0));
    if (e > 3) { r = x2 * x2; for (e -= 4; e > 0; --e) r *= x; }
    return r;
}

// x^e for e in 0..3 without branches: three selects and two multiplications
__device__ __forceinline__ double pw3(double x, int e)
{
    const double a = e >= 1 ? x : 1.0, b = e >= 2 ? x : 1.0, c = e >= 3 ? x : 1.0;
    return (a * b) * c;
}

// rate_j = k_j * prod_m u_m^nu_mj  (Catalyst mass action, combinatoric_ratelaws=false;
// reference src/solving/solve_utils.jl:318-334).  Branch-free: an absent reactant slot reads
// species 0 with exponent 0 (its byte of d.w is 0), so the three gathers of a reaction are always
// issued together; exponents above 3 take the general (rare) path.
__device__ __forceinline__ double rate_of(const int4 d, const double *u, int MB, int m, double kj)
{
    const double x0 = u[max(d.x, 0) * MB + m], x1 = u[max(d.y, 0) * MB + m], x2 = u[max(d.z, 0) * MB + m];
    const int e0 = d.w & 255, e1 = (d.w >> 8) & 255, e2 = (d.w >> 16) & 255;
    if (d.w & 0x00fcfcfc) return kj * pw(x0, e0) * pw(x1, e1) * pw(x2, e2);
    return (kj * pw3(x0, e0)) * (pw3(x1, e1) * pw3(x2, e2));
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

// dk/dT of the same expression (continuous rate updates: the df/dt term of the Rosenbrock stages)
__device__ __forceinline__ double arrhenius_dT(const DevNet &net, int r, double T)
{
    double kr = net.A[r] * exp(-net.Ea[r] / (kR * T));
    double dln = net.Ea[r] / (kR * T * T);                 // d ln k_r / dT
    if (net.n) { kr *= pow(T, net.n[r]); dln += net.n[r] / T; }
    kr = kr * kNA * net.t_mult;
    const double dkr = kr * dln;
    if (isnan(net.k_max)) return dkr;
    const double k = 1.0 / ((1.0 / net.k_max) + (1.0 / kr));
    return (kr > 0.0) ? dkr * (k / kr) * (k / kr) : 0.0;    // k = 1/(1/kmax + 1/kr)  =>  dk = (k/kr)^2 dkr
}

// dX/dt of a member's condition profile (right-continuous at the kinks, which are forced step ends)
__device__ inline double profile_grad(int kind, const double *p, double t)
{
    switch (kind) {
    case 2: case 3: return (t >= 0.0 && t < p[3]) ? p[0] : 0.0;
    case 4: {
        const double r1 = p[1], r2 = p[2], tb = p[7];
        const double ts[2] = {p[3], p[5]}, te[2] = {p[4], p[6]}, rr[2] = {r1, r2};
        double g = 0.0;
        for (int q = 0; q < 2; ++q) {
            if (tb > 0.0) {
                if (t >= ts[q] - tb && t < ts[q] + tb) g += rr[q] * (t - ts[q] + tb) / (2 * tb);
                else if (t >= ts[q] + tb && t < te[q] - tb) g += rr[q];
                else if (t >= te[q] - tb && t < te[q] + tb) g += rr[q] * (1.0 - (t - te[q] + tb) / (2 * tb));
            } else if (t >= ts[q] && t < te[q]) g += rr[q];
        }
        return g;
    }
    default: return 0.0;
    }
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
// Warp-tile primitives.
// ---------------------------------------------------------------------------------------------
#ifndef KB2_LU_PREFETCH
#define KB2_LU_PREFETCH 1     // U' values of the next source block are loaded while the current one is applied
#endif
#ifndef KB2_CHUNK_BULK
#define KB2_CHUNK_BULK 1      // panel chunks go to shared memory with one bulk copy
#endif
#ifndef KB2_TRI_AHEAD
#define KB2_TRI_AHEAD 0        // >= 0: the panel two links ahead of the sweep is pulled into L2 with a bulk prefetch; -1: off
#endif
#ifndef KB2_LU_L2PF
#define KB2_LU_L2PF 1         // bulk L2 prefetch of the next unit's chunk and target spans
#endif
#ifndef KB2_NC
#define KB2_NC 3               // target columns per lane and pass in the LU update
#endif
#ifndef KB2_K_STREAM
#define KB2_K_STREAM 0         // k is loaded with an L2 evict_first policy in the per-reaction passes
#endif
#ifndef KB2_RHS_MINB
#define KB2_RHS_MINB 4           // resident CTAs per SM requested for the multi-warp streaming kernels (register cap)
#endif
constexpr int PR = 8;          // rows per panel (PanelPlan::PR)
constexpr int CWMAX = 96;      // columns per chunk (PanelPlan::CW)
constexpr unsigned FULL = 0xffffffffu;

// ---- bulk asynchronous copies (TMA, 1-D): the large contiguous transfers of a warp — a panel
// chunk, the columns of a panel for a sweep, the state vector — are handed to the copy engine
// with ONE instruction and complete on a shared-memory mbarrier.  They never sit in the
// load/store unit, so the short latency-critical loads of the other warps on the SM (which are
// in other phases of their own solves) do not queue behind them.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_init(unsigned bar)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n"
                 ::"r"(bar), "r"(parity) : "memory");
}
// Per-warp copy channel: one mbarrier and its phase parity, both in the warp's shared memory.
struct BulkChan {
    unsigned bar;      // shared-memory address of the mbarrier (0: bulk copies disabled, e.g. MB = 1 alignment)
    int *par;          // phase parity of the next completion
};
// all lanes call; the copy is issued by lane 0 after the warp has finished with the destination.
// `fence`: the source was written with ordinary stores by this warp since the last fence.
__device__ __forceinline__ void bulk_issue(const BulkChan &ch, void *dst, const void *src, unsigned bytes, int lane, bool fence)
{
    if (fence) fence_proxy_async();
    __syncwarp();
    if (lane == 0) bulk_g2s(smem_u32(dst), src, bytes, ch.bar);
}
__device__ __forceinline__ void bulk_wait(const BulkChan &ch, int lane)
{
    const int p = *ch.par;
    mbar_wait(ch.bar, (unsigned)p);
    __syncwarp();
    if (lane == 0) *ch.par = p ^ 1;
    __syncwarp();
}

// one channel per warp (= per CTA), set up once at kernel start; the mbarrier lives behind the
// data_bytes of dynamic shared memory the kernel was launched with
template <int MB>
__device__ __forceinline__ BulkChan chan_setup(double *smem, int data_bytes)
{
    BulkChan ch;
    char *tail = reinterpret_cast<char *>(smem) + data_bytes;
    ch.bar = MB >= 2 ? smem_u32(tail) : 0u;        // MB = 1: tiles are only 8-byte aligned, bulk copies need 16
    ch.par = reinterpret_cast<int *>(tail + 8);
    if ((threadIdx.x & 31) == 0) {
        if (ch.bar) mbar_init(ch.bar);
        *ch.par = 0;
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncwarp();
    return ch;
}

template <int MB>
struct WTile {
    static constexpr int LN = 32 / MB;
    int lane, m, ln, b;
    double *u, *ua, *rv, *y, *K[6], *k, *rate, *drate, *lu, *invd, *out_u, *out_umax, *uc;
    BulkChan ch;
    __device__ WTile(int tile, const DevNet &net, const DevPlan &pl, const DevEns &en, const BulkChan &chan)
    {
        ch = chan;
        lane = threadIdx.x & 31;
        m = lane % MB;
        ln = lane / MB;
        b = tile * MB + m;
        const size_t vs = (size_t)tile * net.S * MB;
        u = en.u + vs; ua = en.ua + vs; rv = en.rv + vs; y = en.y + vs; invd = en.invd + vs; uc = en.uc + vs;
#pragma unroll
        for (int q = 0; q < 6; ++q) K[q] = en.K[q] + vs;
        k = en.k + (size_t)tile * net.R * MB;
        rate = en.rate + (size_t)tile * net.R * MB;
        drate = en.drate + (size_t)tile * net.R * net.jslots * MB;
        lu = en.lu + (size_t)tile * pl.padded * MB;
        out_u = en.out_u + (size_t)tile * en.Ns * net.S * MB;
        out_umax = en.out_umax + vs;
    }
};

// sum over the LN lanes of each member in a fixed butterfly order (deterministic; every lane of
// the member ends up with the same bits)
template <int MB>
__device__ __forceinline__ double member_sum(double v)
{
#pragma unroll
    for (int off = 16; off >= MB; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
    return v;
}

__device__ __forceinline__ void prefetch_l2_bulk(const void *p, unsigned bytes)
{
    // bulk L2 prefetch (sm_90+): one instruction pulls a whole panel towards L2
    if (bytes >= 16) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(p), "r"(bytes) : "memory");
}

// L2 residency control.  The factors stream through L2 once per triangular sweep (evict_first),
// while the finished U' blocks of the last few panels are re-read by the panels that follow
// (evict_last): with the hints the 126 MB L2 keeps that window instead of the stream.
__device__ __forceinline__ unsigned long long l2_policy_last()
{
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long l2_policy_first()
{
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void st2_hint(double2 *a, double2 v, unsigned long long p)
{
#ifndef KB2_L2_HINTS
    (void)p; *a = v; return;
#endif
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(a), "d"(v.x), "d"(v.y), "l"(p) : "memory");
}
__device__ __forceinline__ double ld_hint(const double *a, unsigned long long p)
{
    double v;
#ifndef KB2_L2_HINTS
    (void)p; return *a;
#endif
    asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(a), "l"(p) : "memory");
    return v;
}
// streamed once per pass (k): keep it from displacing the rate / derivative tables that the
// gather pass is about to re-read from L2
__device__ __forceinline__ double ld_stream(const double *a, unsigned long long p)
{
#if KB2_K_STREAM
    double v;
    asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(a), "l"(p) : "memory");
    return v;
#else
    (void)p; return *a;
#endif
}

// A tile may be worked on by NW warps of one CTA (the streaming phases: right-hand side, Jacobian
// values); phase boundaries inside the tile primitives are then block barriers.
template <int NW>
__device__ __forceinline__ void tile_sync()
{
    if (NW > 1) __syncthreads(); else __syncwarp();
}

// K1: k[r][m] for the member's current condition value T (masked by `upd`); with `kdot` also
// dk/dt = dk/dT * Tdot (continuous rate updates).  NW warps of a CTA may share the tile.
template <int MB, int NW = 1>
__device__ void tile_rates(const WTile<MB> &tl, const DevNet &net, double T, bool upd, int ridx, int w = 0,
                           double *kdot = nullptr, double Tdot = 0.0)
{
    constexpr int LN = 32 / MB, VL = LN * NW;
    if (upd) {
        if (net.calc_mode == 0) {
            for (int r = w * LN + tl.ln; r < net.R; r += VL) {
                tl.k[r * MB + tl.m] = arrhenius(net, r, T);
                if (kdot) kdot[r * MB + tl.m] = Tdot != 0.0 ? arrhenius_dT(net, r, T) * Tdot : 0.0;
            }
        } else {
            const double *src = ridx < 0 ? net.kinit : net.ktab + (size_t)ridx * net.R;
            for (int r = w * LN + tl.ln; r < net.R; r += VL) tl.k[r * MB + tl.m] = src[r];
        }
    }
    tile_sync<NW>();
}

// K2: mass-action right-hand side in two gather passes (no atomics, fixed summation order):
//   rate_j = k_j * prod u^nu                      (lanes over reactions, coalesced k / rate)
//   du_i   = sum_e coef_e * rate_{j(e)}           (gather rows of species i, ascending reactions)
// out_i = du_i (+ out_i when `accumulate`: the caller has put the stage combination there).
constexpr int ELL_G = 64;      // Symbolic::ELL_G
template <int U>
__device__ __forceinline__ void ell_load(const int *p, int (&ix)[U])
{
    if constexpr (U >= 4) {
#pragma unroll
        for (int q = 0; q < U / 4; ++q) {
            const int4 v = reinterpret_cast<const int4 *>(p)[q];
            ix[4 * q] = v.x; ix[4 * q + 1] = v.y; ix[4 * q + 2] = v.z; ix[4 * q + 3] = v.w;
        }
    } else {
        const int2 v = *reinterpret_cast<const int2 *>(p);
        ix[0] = v.x; ix[1] = v.y;
    }
}

// Gather-sum over a sliced ELL:  for every item i = order[nlong + slot] (slot = g*64 + ln*U + v)
//   put(i, pre(i), sum_t coef_t * src[index_t])      (terms in ascending t: fixed order)
// src is tile-major [index][MB]; pre(i) is evaluated a group ahead (whatever the output needs
// that is a load: a base value, a storage slot).  The packed indices of a step depend on no
// data: inside a group they are fetched two steps ahead and the gathers of step t+1 are in flight
// while step t is accumulated; across groups the group record is three ahead, and the items,
// the first two index steps and the first gathers of group g+1 are issued while group g finishes
// (most Jacobian entries have one or two terms: there the group turnaround is what counts).
template <int MB, class Pre, class Put>
__device__ __forceinline__ void ell_gather(const WTile<MB> &tl, const int *order, int n, int nlong, const int *ell_ptr, const int *ell,
                                           int ngroups, const double *src, Pre pre, Put put)
{
    constexpr int LN = 32 / MB, U = ELL_G / LN;
    typedef decltype(pre(0)) H;
    const int m = tl.m, lo = tl.ln * U;
    if (ngroups <= 0) return;
    int q0 = ell_ptr[0], q1 = ell_ptr[1], q2 = ell_ptr[min(2, ngroups)];
    int sp[U], ia[U], ib[U];
    double ra[U];
    H hh[U];
#pragma unroll
    for (int v = 0; v < U; ++v) {
        sp[v] = nlong + lo + v < n ? order[nlong + lo + v] : -1;
        ia[v] = ib[v] = 0; ra[v] = 0.0;
    }
    if (q1 > q0) {
        ell_load<U>(ell + q0 + lo, ia);
        ell_load<U>(ell + q0 + lo + min(1, (q1 - q0) / ELL_G - 1) * ELL_G, ib);
    }
#pragma unroll
    for (int v = 0; v < U; ++v) hh[v] = sp[v] >= 0 ? pre(sp[v]) : H();
    if (q1 > q0) {
#pragma unroll
        for (int v = 0; v < U; ++v) ra[v] = src[(ia[v] & 0xffffff) * MB + m];
    }
    for (int g = 0; g < ngroups; ++g) {
        const int len = (q1 - q0) / ELL_G, lenN = (q2 - q1) / ELL_G;
        const int q3 = ell_ptr[min(g + 3, ngroups)];
        // group g+1: items and the first two index steps
        const int zn = nlong + (g + 1) * ELL_G + lo;
        int spn[U], ja[U], jb[U];
#pragma unroll
        for (int v = 0; v < U; ++v) {
            spn[v] = zn + v < n ? order[zn + v] : -1;
            ja[v] = jb[v] = 0;
        }
        if (lenN > 0) {
            ell_load<U>(ell + q1 + lo, ja);
            ell_load<U>(ell + q1 + lo + min(1, lenN - 1) * ELL_G, jb);
        }
        // group g
        double acc[U];
#pragma unroll
        for (int v = 0; v < U; ++v) acc[v] = 0.0;
        const int *ep = ell + q0 + lo;
        for (int t = 0; t < len; ++t) {
            int ic[U];
            double rb[U];
            ell_load<U>(ep + min(t + 2, len - 1) * ELL_G, ic);
#pragma unroll
            for (int v = 0; v < U; ++v) rb[v] = src[(ib[v] & 0xffffff) * MB + m];
#pragma unroll
            for (int v = 0; v < U; ++v) acc[v] += (double)(ia[v] >> 24) * ra[v];
#pragma unroll
            for (int v = 0; v < U; ++v) { ia[v] = ib[v]; ra[v] = rb[v]; ib[v] = ic[v]; }
        }
        // group g+1: what its outputs need, and its first gathers
        H hn[U];
        double rn[U];
#pragma unroll
        for (int v = 0; v < U; ++v) hn[v] = spn[v] >= 0 ? pre(spn[v]) : H();
#pragma unroll
        for (int v = 0; v < U; ++v) rn[v] = lenN > 0 ? src[(ja[v] & 0xffffff) * MB + m] : 0.0;
#pragma unroll
        for (int v = 0; v < U; ++v)
            if (sp[v] >= 0) put(sp[v], hh[v], acc[v]);
#pragma unroll
        for (int v = 0; v < U; ++v) { sp[v] = spn[v]; hh[v] = hn[v]; ia[v] = ja[v]; ib[v] = jb[v]; ra[v] = rn[v]; }
        q0 = q1; q1 = q2; q2 = q3;
    }
}

// Contiguous share [g0, g1) of the ELL groups for warp w of NW, balanced by ELL length (the groups
// are sorted by decreasing length, so equal counts would not be equal work).
template <int NW>
__device__ __forceinline__ void ell_share(const int *ell_ptr, int ngroups, int w, int &g0, int &g1)
{
    if (NW == 1) { g0 = 0; g1 = ngroups; return; }
    const int total = ell_ptr[ngroups] - ell_ptr[0];
    auto cut = [&](int k) {        // first group whose start is at or beyond k/NW of the total (plus one group per warp as a floor)
        if (k <= 0) return 0;
        if (k >= NW) return ngroups;
        const long long target = (long long)total * k / NW + ell_ptr[0];
        int lo = 0, hi = ngroups;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (ell_ptr[mid] < target) lo = mid + 1; else hi = mid; }
        return lo;
    };
    g0 = cut(w); g1 = cut(w + 1);
}

// state vector of the tile -> shared memory (all NW warps copy, 16 bytes per thread and step)
template <int MB, int NW>
__device__ __forceinline__ void stage_vector(const WTile<MB> &tl, double *su, const double *u, int n)
{
    if (NW == 1 && tl.ch.bar) {
        bulk_issue(tl.ch, su, u, (unsigned)(n * 8), tl.lane, true);
        bulk_wait(tl.ch, tl.lane);
        return;
    }
    const int tid = NW == 1 ? tl.lane : (int)threadIdx.x;
    if (MB >= 2) {
        const double2 *s2 = reinterpret_cast<const double2 *>(u);
        double2 *d2 = reinterpret_cast<double2 *>(su);
        for (int i = tid; i < n / 2; i += NW * 32) d2[i] = s2[i];
    } else {
        for (int i = tid; i < n; i += NW * 32) su[i] = u[i];
    }
    tile_sync<NW>();
}

template <int MB, int NW = 1>
__device__ void tile_rhs(const WTile<MB> &tl, const DevNet &net, const double *u, double *out, bool accumulate, double *su, int w = 0)
{
    constexpr int LN = 32 / MB;
    const int m = tl.m;
    if (su) {       // stage the state vector in shared memory: the reactant gathers never leave the SM
        stage_vector<MB, NW>(tl, su, u, net.S * MB);
        u = su;
    }
    {
        constexpr int UR = 4, VL = LN * NW;       // virtual lanes of a member over the NW warps
        const unsigned long long pol_k = l2_policy_first();
        for (int j0 = w * LN + tl.ln; j0 < net.R; j0 += UR * VL) {
            int4 d[UR];
            double kj[UR];
#pragma unroll
            for (int v = 0; v < UR; ++v) {
                const int j = min(j0 + v * VL, net.R - 1);
                d[v] = net.rdesc[j];
                kj[v] = ld_stream(tl.k + j * MB + m, pol_k);
            }
            double rt[UR];
            int ps[UR];
#pragma unroll
            for (int v = 0; v < UR; ++v) ps[v] = net.rate_pos[min(j0 + v * VL, net.R - 1)];
#pragma unroll
            for (int v = 0; v < UR; ++v) rt[v] = rate_of(d[v], u, MB, m, kj[v]);
#pragma unroll
            for (int v = 0; v < UR; ++v)
                if (j0 + v * VL < net.R) tl.rate[ps[v] * MB + m] = rt[v];
        }
    }
    tile_sync<NW>();
    // rows of hub species (hundreds of terms): the LN lanes of the member stride over the terms
    for (int z = w; z < net.rhs_nlong; z += NW) {
        const int i = net.rhs_order[z];
        const int e1 = net.rhs_ptr[i + 1];
        const double b0 = (accumulate && tl.ln == 0) ? out[i * MB + m] : 0.0;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int e = net.rhs_ptr[i] + tl.ln;
        for (; e + 3 * LN < e1; e += 4 * LN) {
            a0 += (double)net.rhs_coef[e] * tl.rate[net.rhs_rxn[e] * MB + m];
            a1 += (double)net.rhs_coef[e + LN] * tl.rate[net.rhs_rxn[e + LN] * MB + m];
            a2 += (double)net.rhs_coef[e + 2 * LN] * tl.rate[net.rhs_rxn[e + 2 * LN] * MB + m];
            a3 += (double)net.rhs_coef[e + 3 * LN] * tl.rate[net.rhs_rxn[e + 3 * LN] * MB + m];
        }
        for (; e < e1; e += LN) a0 += (double)net.rhs_coef[e] * tl.rate[net.rhs_rxn[e] * MB + m];
        const double a = member_sum<MB>((a0 + a1) + (a2 + a3));
        if (tl.ln == 0) out[i * MB + m] = a + b0;
    }
    // the other rows one per lane slot as a sliced ELL (groups of 64 rows of about equal length)
    int g0, g1;
    ell_share<NW>(net.ell_ptr, net.ell_ngroups, w, g0, g1);
    ell_gather<MB>(tl, net.rhs_order, net.S, net.rhs_nlong + g0 * ELL_G, net.ell_ptr + g0, net.ell, g1 - g0, tl.rate,
                   [&](int i) { return accumulate ? out[i * MB + m] : 0.0; },
                   [&](int i, double b0, double a) { out[i * MB + m] = a + b0; });
    tile_sync<NW>();
}

// K3: analytic Jacobian entries  J_p = sum_t coef_t * k_j * d(prod)/du_l  for every entry p of
// the fixed CSC pattern, handed to `put(p, J_p)`.  Two passes like the right-hand side:
//   d[j][s] = k_j * d(prod_j)/du_(slot s) / nu_s     (lanes over reactions: k, descriptors and
//                                                      the table itself are coalesced streams)
//   J_p     = sum_t coef_t * d[index_t]               (gather-sum: entries with many terms (hub
//             columns) are split across the lanes of the member, the others go through the ELL)
template <int MB, class Pre, class Put, int NW = 1>
__device__ __forceinline__ void tile_jac_entries(const WTile<MB> &tl, const DevNet &net, const double *u, Pre pre, Put put, int w = 0)
{
    constexpr int LN = 32 / MB;
    const int m = tl.m, ns = net.jslots;
    {
        constexpr int UR = 4, VL = LN * NW;
        const unsigned long long pol_k = l2_policy_first();
        for (int j0 = w * LN + tl.ln; j0 < net.R; j0 += UR * VL) {
            int4 d[UR];
            double kj[UR];
#pragma unroll
            for (int v = 0; v < UR; ++v) {
                const int j = min(j0 + v * VL, net.R - 1);
                d[v] = net.rdesc[j];
                kj[v] = ld_stream(tl.k + j * MB + m, pol_k);
            }
#pragma unroll
            for (int v = 0; v < UR; ++v) {
                const int j = j0 + v * VL;
                const double x0 = u[max(d[v].x, 0) * MB + m], x1 = u[max(d[v].y, 0) * MB + m], x2 = u[max(d[v].z, 0) * MB + m];
                const int e0 = d[v].w & 255, e1 = (d[v].w >> 8) & 255, e2 = (d[v].w >> 16) & 255;
                double g0, g1, g2;
                if (d[v].w & 0x00fcfcfc) {
                    g0 = kj[v] * pw(x0, max(e0 - 1, 0)) * pw(x1, e1) * pw(x2, e2);
                    g1 = kj[v] * pw(x0, e0) * pw(x1, max(e1 - 1, 0)) * pw(x2, e2);
                    g2 = kj[v] * pw(x0, e0) * pw(x1, e1) * pw(x2, max(e2 - 1, 0));
                } else {
                    // full powers p_s and powers reduced by one q_s: d/du_s = k * q_s * prod_{r != s} p_r
                    const double q0 = pw3(x0, e0 - 1), q1 = pw3(x1, e1 - 1), q2 = pw3(x2, e2 - 1);
                    const double p0 = e0 >= 1 ? q0 * x0 : 1.0, p1 = e1 >= 1 ? q1 * x1 : 1.0, p2 = e2 >= 1 ? q2 * x2 : 1.0;
                    g0 = (kj[v] * q0) * (p1 * p2);
                    g1 = (kj[v] * q1) * (p0 * p2);
                    g2 = (kj[v] * q2) * (p0 * p1);
                }
                if (j < net.R) {
                    const int *dq = net.drate_pos + j * ns;
                    tl.drate[dq[0] * MB + m] = g0;
                    if (ns > 1) tl.drate[dq[1] * MB + m] = g1;
                    if (ns > 2) tl.drate[dq[2] * MB + m] = g2;
                }
            }
        }
    }
    tile_sync<NW>();
    for (int z = w; z < net.j_nlong; z += NW) {
        const int p = net.j_order[z];
        const int t1 = net.jt_ptr[p + 1];
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int t = net.jt_ptr[p] + tl.ln;
        for (; t + 3 * LN < t1; t += 4 * LN) {
            const int k0 = net.jt_pk[t], k1 = net.jt_pk[t + LN], k2 = net.jt_pk[t + 2 * LN], k3 = net.jt_pk[t + 3 * LN];
            a0 += (double)(k0 >> 24) * tl.drate[(k0 & 0xffffff) * MB + m];
            a1 += (double)(k1 >> 24) * tl.drate[(k1 & 0xffffff) * MB + m];
            a2 += (double)(k2 >> 24) * tl.drate[(k2 & 0xffffff) * MB + m];
            a3 += (double)(k3 >> 24) * tl.drate[(k3 & 0xffffff) * MB + m];
        }
        for (; t < t1; t += LN) {
            const int k0 = net.jt_pk[t];
            a0 += (double)(k0 >> 24) * tl.drate[(k0 & 0xffffff) * MB + m];
        }
        const double a = member_sum<MB>((a0 + a1) + (a2 + a3));
        if (tl.ln == 0) put(p, pre(p), a);
    }
    int g0, g1;
    ell_share<NW>(net.jell_ptr, net.jell_ngroups, w, g0, g1);
    ell_gather<MB>(tl, net.j_order, net.nnzJ, net.j_nlong + g0 * ELL_G, net.jell_ptr + g0, net.jell, g1 - g0, tl.drate, pre, put);
}

template <int MB, int NW = 1>
__device__ void tile_jac_csc(const WTile<MB> &tl, const DevNet &net, const double *u, double *Jval, int w = 0)
{
    const int m = tl.m;
    auto pre = [](int p) { return p; };
    auto put = [&](int, int p, double v) { Jval[p * MB + m] = v; };
    tile_jac_entries<MB, decltype(pre), decltype(put), NW>(tl, net, u, pre, put, w);
    tile_sync<NW>();
}

// W = I/(h*gamma) - J assembled straight into the padded block storage: zero fill (coalesced
// 16-byte stores), -J entries scattered to their slots (four entries per lane in flight), then
// the diagonal shift.
template <int MB>
__device__ void tile_assemble_w(const WTile<MB> &tl, const DevNet &net, const DevPlan &pl, const double *u, double hg_inv, double *su)
{
    constexpr int LN = 32 / MB;
    const int m = tl.m;
    if (su) {
        stage_vector<MB, 1>(tl, su, u, net.S * MB);
        u = su;
    }
    if (MB == 1) {
        for (int i = tl.lane; i < pl.padded; i += 32) tl.lu[i] = 0.0;
    } else {
        double2 *z = reinterpret_cast<double2 *>(tl.lu);       // padded*MB is even: every tile is 16-byte aligned
        const int n2 = pl.padded * MB / 2;
        const unsigned long long pol = l2_policy_first();
        for (int i = tl.lane; i < n2; i += 32) st2_hint(z + i, make_double2(0.0, 0.0), pol);
    }
    __syncwarp();
    tile_jac_entries<MB>(tl, net, u, [&](int p) { return net.jslot[p]; }, [&](int, int slot, double v) { tl.lu[(size_t)slot * MB + m] = -v; });
    __syncwarp();
    for (int i = tl.ln; i < net.S; i += LN) tl.lu[(size_t)net.diag_slot[i] * MB + m] += hg_inv;
    fence_proxy_async();        // the factorisation reads these values with bulk copies
    __syncwarp();
}

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_but1() { asm volatile("cp.async.wait_group 1;\n" ::: "memory"); }

// shared memory of a factorising warp, in doubles: one staged chunk of a panel, the unit upper
// diagonal block U'_QQ of the next in-chunk source (prefetched with cp.async) and the finished
// L'_PQ block of a source that lies in an earlier chunk
__host__ __device__ constexpr size_t lu_smem_doubles(int mb) { return (size_t)(CWMAX * PR + 3 * PR * PR) * mb; }

// K4: block Crout LU of the padded panel storage, in place:  W = L' U'  with the pivots on L'
// and a unit diagonal on U'; invd[i] = 1/L'_ii is kept for the substitutions.
// A unit (panel x column chunk) is staged in shared memory Wp[c][r][m]; for every source block
// Q in the L part:   L_PQ = X * inv(U_QQ)  (lane ln owns row ln),  then every lane updates its
// share of the target columns, three columns per pass,
//     w[:, j] -= L_PQ * U_Q[:, j]
// with L_PQ broadcast from shared memory (8 loads feed 24 FMAs).
// The source blocks of a unit form a software pipeline, because one warp has nobody else to hide
// its latencies behind: while source k is applied, the task record and target map of source k+1
// are already in registers, its U' values (first pass) are in flight into registers, and its
// U'_QQ or L'_PQ block is in flight into shared memory with cp.async; the next unit's chunk and
// target spans are pulled towards L2 with bulk prefetches.
template <int MB>
__device__ void tile_lu(const WTile<MB> &tl, const DevPlan &pl, double *Wp)
{
    constexpr int LN = 32 / MB, NC = KB2_NC;
    constexpr int RL = LN < PR ? LN : PR, RPL = PR / RL;      // lanes that own rows, rows per such lane (row = ln + rr*RL)
    const int m = tl.m, ln = tl.ln, lane = tl.lane;
    double *lu = tl.lu;
    double *Uq = Wp + CWMAX * PR * MB;       // [q][a][m]: U'_QQ of the next in-chunk source
    double *Ls = Uq + PR * PR * MB;          // [2][q][r][m]: L'_PQ of sources from an earlier chunk (double buffered)
    const unsigned long long pol_first = l2_policy_first(), pol_last = l2_policy_last();
    int4 na = pl.u_info[0], nb = pl.u_info[1], nc = pl.u_info[2];
    for (int un = 0; un < pl.nunits; ++un) {
        const int4 ua = na, ub = nb, uc = nc;
        const int x0 = ua.y, x1 = ua.z, task0 = ua.w, ntask = ub.x, dmode = ub.y;
        const int nr = uc.x, next = uc.y, p0 = uc.z;
        const int cw = x1 - x0;
        double *gP = lu + (size_t)uc.w * MB;
        const int4 *trec = pl.t_info + (size_t)3 * task0;
        // ---- the chunk goes to shared memory with one bulk copy (cp.async group C without a copy
        // engine channel): it is in flight while the prologue below runs ----
        {
            const double *src = gP + (size_t)x0 * nr * MB;
            const int n = cw * nr * MB;
            if (KB2_CHUNK_BULK && tl.ch.bar) {
                bulk_issue(tl.ch, Wp, src, (unsigned)(n * 8), lane, false);
            } else if (MB == 1) {
                for (int i = lane; i < n; i += 32) cp_async8(Wp + i, src + i);
            } else {
                // eight 16-byte loads per lane in flight
                const double2 *s2 = reinterpret_cast<const double2 *>(src);
                double2 *d2 = reinterpret_cast<double2 *>(Wp);
                const int n2 = n / 2;
                for (int i0 = lane; i0 < n2; i0 += 32 * 8) {
                    double2 v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = i0 + 32 * j < n2 ? s2[i0 + 32 * j] : make_double2(0.0, 0.0);
#pragma unroll
                    for (int j = 0; j < 8; ++j) if (i0 + 32 * j < n2) d2[i0 + 32 * j] = v[j];
                }
            }
            cp_async_commit();
        }
        // ---- prologue of the task pipeline: records of the first two sources, the first one's
        // staged block and target map ----
        int4 ca = make_int4(0, 0, 0, 0), cb = ca, cc = ca, xa = ca, xb = ca, xc = ca;
        if (ntask > 0) { ca = trec[0]; cb = trec[1]; cc = trec[2]; }
        if (ntask > 1) { xa = trec[3]; xb = trec[4]; xc = trec[5]; }
        if (ntask > 0 && !(ca.y >> 30)) {        // group A(-1): L' block of a first source that lies in an earlier chunk
            const double *src = gP + (size_t)(ca.y & 0x3fffffff) * nr * MB;
            for (int i = lane; i < cb.x * nr * MB; i += 32) cp_async8(Ls + i, src + i);
        }
        cp_async_commit();
        if (ub.z >= 0) {                         // group B(-1): U'_QQ of the first in-chunk source
            const double *src = lu + (size_t)ub.z * MB;
            for (int i = lane; i < ub.w * ub.w * MB; i += 32) cp_async8(Uq + i, src + i);
        }
        cp_async_commit();
        // next unit: its record, and L2 prefetch of its chunk and of its sources' target spans
        if (un + 1 < pl.nunits) {
            na = pl.u_info[3 * un + 3]; nb = pl.u_info[3 * un + 4]; nc = pl.u_info[3 * un + 5];
        }
        if (KB2_LU_L2PF && un + 1 < pl.nunits) {
            if (lane == 31) {
                const char *a = (const char *)(lu + ((size_t)nc.w + (size_t)na.y * nc.x) * MB);
                const size_t nbytes = (size_t)(na.z - na.y) * nc.x * MB * 8;
                const size_t a16 = (size_t)a & ~(size_t)15;
                prefetch_l2_bulk((const void *)a16, (unsigned)(((size_t)a + nbytes - a16) & ~(size_t)15));
            } else {
                for (int tk = lane; tk < nb.x; tk += 31) {
                    const int4 t2 = pl.t_info[(size_t)3 * (na.w + tk) + 2];
                    if (t2.z > 0) {
                        const char *a = (const char *)(lu + (size_t)t2.y * MB);
                        const size_t nbytes = (size_t)t2.z * MB * 8;
                        const size_t a16 = (size_t)a & ~(size_t)15;
                        prefetch_l2_bulk((const void *)a16, (unsigned)(((size_t)a + nbytes - a16) & ~(size_t)15));
                    }
                }
            }
        }
        int cmap[NC];
        double cu[NC][PR];
#pragma unroll
        for (int c = 0; c < NC; ++c) cmap[c] = (ntask > 0 && ca.z > 0) ? pl.map[ca.w + min(ln + c * LN, ca.z - 1)] : 0;
        // U' values of the first source's first pass
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const double *up = lu + ((size_t)cb.y + (size_t)(cmap[c] & 0xffff) * cb.x) * MB + m;
#pragma unroll
            for (int q = 0; q < PR; ++q) cu[c][q] = (KB2_LU_PREFETCH && ntask > 0 && ca.z > 0 && q < cb.x) ? ld_hint(up + q * MB, pol_last) : 0.0;
        }
        if (KB2_CHUNK_BULK && tl.ch.bar) bulk_wait(tl.ch, lane);      // the chunk has landed
        __syncwarp();
        for (int tk = 0; tk < ntask; ++tk) {
            const int lpos = ca.y & 0x3fffffff, inch = ca.y >> 30, ntg = ca.z, map0 = ca.w;
            const int nq = cb.x;
            const double *gQ = lu + (size_t)cb.y * MB;
            const bool has_next = tk + 1 < ntask;
            // ---- look ahead: record of source k+2, target map of source k+1, staged L' of k+1 ----
            int4 ya = make_int4(0, 0, 0, 0), yb = ya, yc = ya;
            if (tk + 2 < ntask) { ya = trec[3 * (tk + 2)]; yb = trec[3 * (tk + 2) + 1]; yc = trec[3 * (tk + 2) + 2]; }
            int xmap[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) xmap[c] = (has_next && xa.z > 0) ? pl.map[xa.w + min(ln + c * LN, xa.z - 1)] : 0;
            if (has_next && !(xa.y >> 30)) {     // group A(k): L' block of source k+1
                const double *src = gP + (size_t)(xa.y & 0x3fffffff) * nr * MB;
                double *dst = Ls + ((tk + 1) & 1) * PR * PR * MB;
                for (int i = lane; i < xb.x * nr * MB; i += 32) cp_async8(dst + i, src + i);
            }
            cp_async_commit();
            cp_async_wait_but1();                // everything but A(k) has landed: L' or U'_QQ of source k
            __syncwarp();
            const double *ls;                    // L'_PQ as [q][r][m] in shared memory
            if (inch) {
#pragma unroll
                for (int rr = 0; rr < RPL; ++rr) {
                    const int row = ln + rr * RL;
                    if (ln < RL && row < nr) {
                        double X[PR];
                        double *xp = Wp + ((lpos - x0) * nr + row) * MB + m;
#pragma unroll
                        for (int q = 0; q < PR; ++q) X[q] = q < nq ? xp[q * nr * MB] : 0.0;
                        const double *uq = Uq + m;                 // U_QQ[a][q] at (q*nq + a)*MB
#pragma unroll
                        for (int a = 0; a < PR - 1; ++a)
#pragma unroll
                            for (int q = a + 1; q < PR; ++q)
                                if (q < nq) X[q] -= X[a] * uq[(q * nq + a) * MB];
#pragma unroll
                        for (int q = 1; q < PR; ++q) if (q < nq) xp[q * nr * MB] = X[q];
                    }
                }
                __syncwarp();
                if (cb.w >= 0) {                 // group B(k): U'_QQ of the next in-chunk source
                    const double *src = lu + (size_t)cb.w * MB;
                    for (int i = lane; i < cc.x * cc.x * MB; i += 32) cp_async8(Uq + i, src + i);
                }
                ls = Wp + (lpos - x0) * nr * MB + m;
            } else {
                ls = Ls + (tk & 1) * PR * PR * MB + m;
            }
            cp_async_commit();
            // ---- first pass: U' values were loaded during the previous source ----
            double xu[NC][PR];
            if (ntg > 0) {
                double w[NC][PR];
                double *wp[NC];
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    wp[c] = Wp + (cmap[c] >> 16) * nr * MB + m;
                    if (!KB2_LU_PREFETCH) {
                        const double *up = gQ + (size_t)(cmap[c] & 0xffff) * nq * MB + m;
#pragma unroll
                        for (int q = 0; q < PR; ++q) cu[c][q] = q < nq ? ld_hint(up + q * MB, pol_last) : 0.0;
                    }
                    // rows >= nr of w and of the L' block are whatever follows in shared memory: they are
                    // computed but never stored, and rows do not mix in the update
#pragma unroll
                    for (int r = 0; r < PR; ++r) w[c][r] = wp[c][r * MB];
                }
                // U' values of source k+1 fly while source k is applied
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const double *up = lu + ((size_t)xb.y + (size_t)(xmap[c] & 0xffff) * xb.x) * MB + m;
#pragma unroll
                    for (int q = 0; q < PR; ++q) xu[c][q] = (KB2_LU_PREFETCH && has_next && xa.z > 0 && q < xb.x) ? ld_hint(up + q * MB, pol_last) : 0.0;
                }
                {
                    // L' column q+1 is loaded while column q is applied
                    double l[PR], ln1[PR];
#pragma unroll
                    for (int r = 0; r < PR; ++r) l[r] = ls[r * MB];
#pragma unroll
                    for (int q = 0; q < PR; ++q) {
                        if (q < nq) {
                            if (q + 1 < PR) {
                                const double *lq = ls + min(q + 1, nq - 1) * nr * MB;
#pragma unroll
                                for (int r = 0; r < PR; ++r) ln1[r] = lq[r * MB];
                            }
#pragma unroll
                            for (int c = 0; c < NC; ++c)
#pragma unroll
                                for (int r = 0; r < PR; ++r) w[c][r] -= l[r] * cu[c][q];
#pragma unroll
                            for (int r = 0; r < PR; ++r) l[r] = ln1[r];
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < NC; ++c)
                    if (ln + c * LN < ntg) {
#pragma unroll
                        for (int r = 0; r < PR; ++r) if (r < nr) wp[c][r * MB] = w[c][r];
                    }
            } else {
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const double *up = lu + ((size_t)xb.y + (size_t)(xmap[c] & 0xffff) * xb.x) * MB + m;
#pragma unroll
                    for (int q = 0; q < PR; ++q) xu[c][q] = (KB2_LU_PREFETCH && has_next && xa.z > 0 && q < xb.x) ? ld_hint(up + q * MB, pol_last) : 0.0;
                }
            }
            // ---- further passes of a wide source: loaded on demand ----
            for (int t = ln + NC * LN; t < ntg; t += NC * LN) {
                double uu[NC][PR], w[NC][PR];
                double *wp[NC];
                bool ok[NC];
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    ok[c] = t + c * LN < ntg;
                    const int e = pl.map[map0 + (ok[c] ? t + c * LN : t)];
                    const double *up = gQ + (size_t)(e & 0xffff) * nq * MB + m;
                    wp[c] = Wp + (e >> 16) * nr * MB + m;
#pragma unroll
                    for (int q = 0; q < PR; ++q) uu[c][q] = q < nq ? ld_hint(up + q * MB, pol_last) : 0.0;
                }
#pragma unroll
                for (int c = 0; c < NC; ++c)
#pragma unroll
                    for (int r = 0; r < PR; ++r) w[c][r] = wp[c][r * MB];
#pragma unroll
                for (int q = 0; q < PR; ++q) {
                    if (q < nq) {
                        double l[PR];
#pragma unroll
                        for (int r = 0; r < PR; ++r) l[r] = ls[(q * nr + r) * MB];
#pragma unroll
                        for (int c = 0; c < NC; ++c)
#pragma unroll
                            for (int r = 0; r < PR; ++r) w[c][r] -= l[r] * uu[c][q];
                    }
                }
#pragma unroll
                for (int c = 0; c < NC; ++c)
                    if (ok[c]) {
#pragma unroll
                        for (int r = 0; r < PR; ++r) if (r < nr) wp[c][r * MB] = w[c][r];
                    }
            }
            __syncwarp();
            // rotate the pipeline
            ca = xa; cb = xb; cc = xc; xa = ya; xb = yb; xc = yc;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                cmap[c] = xmap[c];
#pragma unroll
                for (int q = 0; q < PR; ++q) cu[c][q] = xu[c][q];
            }
        }
        cp_async_wait_all();
        __syncwarp();
        if (dmode) {
            const int dpos = next - x0;      // position of the diagonal block relative to the chunk (negative for dmode 2)
            if (dmode == 1) {
                // Crout LU of the nr x nr diagonal block: lane ln (< RL) owns rows ln + rr*RL, pivot
                // rows are broadcast with shuffles
                double D[RPL][PR];
#pragma unroll
                for (int rr = 0; rr < RPL; ++rr) {
                    const int row = ln + rr * RL;
                    const double *dp = Wp + (dpos * nr + row) * MB + m;
#pragma unroll
                    for (int j = 0; j < PR; ++j) D[rr][j] = (ln < RL && row < nr && j < nr) ? dp[j * nr * MB] : 0.0;
                }
#pragma unroll
                for (int j = 0; j < PR; ++j) {
                    if (j < nr) {
                        const int oj = j % RL, sj = j / RL;          // owner lane and register slot of pivot row j
                        const double piv = __shfl_sync(FULL, D[sj][j], oj * MB + m);
                        const double inv = 1.0 / piv;
                        if (ln == oj) tl.invd[(p0 + j) * MB + m] = inv;
#pragma unroll
                        for (int i = j + 1; i < PR; ++i) {
                            const double uji = __shfl_sync(FULL, D[sj][i], oj * MB + m) * inv;
#pragma unroll
                            for (int rr = 0; rr < RPL; ++rr) {
                                const int row = ln + rr * RL;
                                if (ln == oj && rr == sj) D[rr][i] = uji;
                                else if (row > j && ln < RL) D[rr][i] -= D[rr][j] * uji;
                            }
                        }
                    }
                }
#pragma unroll
                for (int rr = 0; rr < RPL; ++rr) {
                    const int row = ln + rr * RL;
                    if (ln < RL && row < nr) {
                        double *dp = Wp + (dpos * nr + row) * MB + m;
#pragma unroll
                        for (int j = 0; j < PR; ++j) if (j < nr) dp[j * nr * MB] = D[rr][j];
                    }
                }
                __syncwarp();
            }
            // U part of the chunk:  U'_P[:, j] = inv(L'_PP) * w[:, j]  (forward substitution, thread local)
            const int c0 = dmode == 1 ? dpos + nr : 0;
            if (c0 < cw) {
                double Lpp[PR][PR], inv[PR];
                const double *lp = dmode == 1 ? (const double *)(Wp + dpos * nr * MB + m) : (gP + (size_t)next * nr * MB + m);
#pragma unroll
                for (int r = 0; r < PR; ++r) {
                    inv[r] = r < nr ? tl.invd[(p0 + r) * MB + m] : 0.0;
#pragma unroll
                    for (int a = 0; a < PR; ++a) Lpp[r][a] = (a < r && r < nr) ? lp[(a * nr + r) * MB] : 0.0;
                }
                for (int t = c0 + ln; t < cw; t += LN) {
                    double *wp = Wp + t * nr * MB + m;
                    double w[PR];
#pragma unroll
                    for (int r = 0; r < PR; ++r) w[r] = r < nr ? wp[r * MB] : 0.0;
#pragma unroll
                    for (int r = 0; r < PR; ++r) {
#pragma unroll
                        for (int a = 0; a < r; ++a) w[r] -= Lpp[r][a] * w[a];
                        w[r] *= inv[r];
                    }
#pragma unroll
                    for (int r = 0; r < PR; ++r) if (r < nr) wp[r * MB] = w[r];
                }
                __syncwarp();
            }
        }
        {
            double *dst = gP + (size_t)x0 * nr * MB;
            const int n = cw * nr * MB;
            if (MB == 1) {
                for (int i = lane; i < n; i += 32) dst[i] = Wp[i];
            } else {
                // L' columns are only read again by the forward sweeps (stream them through L2);
                // the diagonal block and U' columns are re-read by the next panels (keep them)
                double2 *d2 = reinterpret_cast<double2 *>(dst);
                const double2 *s2 = reinterpret_cast<const double2 *>(Wp);
                const int nl = min(max(next - x0, 0), cw) * nr * MB / 2;
                for (int i = lane; i < nl; i += 32) st2_hint(d2 + i, s2[i], pol_first);
                for (int i = nl + lane; i < n / 2; i += 32) st2_hint(d2 + i, s2[i], pol_last);
            }
        }
        __syncwarp();
    }
}

// K5: W x = rhs over the block storage.  rhs, x in species order; y = the permuted intermediate.
// A sweep is a chain over the panels, so what counts is the latency of one link.  Everything a
// link needs that does not depend on the previous link is loaded ahead of it:
//   * the permuted vector y lives in the warp's shared memory for both sweeps (when S*MB doubles
//     fit; the factorisation is done with the buffer by then), so the only data-dependent loads of
//     the chain are shared-memory reads;
//   * during the dot products every lane of a member is a COLUMN lane: it owns the columns
//     cb + ln, cb + ln + LN, ... of the panel and accumulates all eight rows (one index load and
//     one y read per column instead of one per row, the eight values of a column at immediate
//     offsets), then a reduce-scatter hands row r to lane r for the in-panel substitution;
//   * the first batch of columns of the NEXT panel, its right-hand side, pivots and diagonal block
//     are loaded into registers before the substitution chain of the current panel starts, the
//     panel after that is pulled into L2 with one bulk prefetch, panel records are three deep.
struct PanelMeta { int nr, next, p0, cptr, base, width; };
__device__ __forceinline__ PanelMeta load_pm(const DevPlan &pl, int P)
{
    PanelMeta q;
    const bool ok = P >= 0 && P < pl.npanels;
    const int Pc = ok ? P : 0;
    q.nr = ok ? pl.p_nrows[Pc] : 0;
    q.next = ok ? pl.p_next[Pc] : 0;
    q.p0 = ok ? pl.p_row0[Pc] : 0;
    q.cptr = ok ? pl.p_cptr[Pc] : 0;
    q.base = ok ? pl.p_base[Pc] : 0;
    q.width = ok ? pl.p_width[Pc] : 0;
    return q;
}

// acc[r] = partial sum of row r on every lane of a member -> out[rr] = total of row
// (ln % RL) + rr*RL.  Column-group bits (LN > 8) are folded with a butterfly, the row-lane bits
// with a reduce-scatter: at the level of bit hb a lane keeps the rows whose bit hb equals its own
// and sends the others to its partner.  Fixed order, deterministic.
template <int MB>
__device__ __forceinline__ void rows_reduce(double (&acc)[PR], int ln, double (&out)[PR / ((32 / MB) < PR ? (32 / MB) : PR)])
{
    constexpr int LN = 32 / MB, RL = LN < PR ? LN : PR, RPL = PR / RL;
#pragma unroll
    for (int off = 16; off >= RL * MB; off >>= 1)
#pragma unroll
        for (int r = 0; r < PR; ++r) acc[r] += __shfl_xor_sync(FULL, acc[r], off);
#pragma unroll
    for (int hb = RL / 2; hb >= 1; hb >>= 1) {
        const bool up = (ln & hb) != 0;
        const int done = (RL - 1) & ~(2 * hb - 1);      // row bits already scattered
#pragma unroll
        for (int r = 0; r < PR; ++r) {
            if ((r & hb) == 0 && (r & done) == 0) {
                const double send = up ? acc[r] : acc[r | hb];
                const double keep = up ? acc[r | hb] : acc[r];
                acc[r] = keep + __shfl_xor_sync(FULL, send, hb * MB);
            }
        }
    }
#pragma unroll
    for (int rr = 0; rr < RPL; ++rr) out[rr] = acc[rr * RL];
}

#ifndef KB2_TRI_TU
#define KB2_TRI_TU 4           // columns per lane in one batch of a sweep (LN*TU columns per batch)
#endif

template <int MB, bool YS, bool FWD>
__device__ __forceinline__ void tri_sweep(const WTile<MB> &tl, const DevNet &net, const DevPlan &pl, const double *rhs, double *x, double *y)
{
    constexpr int LN = 32 / MB, RL = LN < PR ? LN : PR, RPL = PR / RL, TU = KB2_TRI_TU;
    const int m = tl.m, ln = tl.ln, r0 = ln % RL, cg = ln / RL;
    const double *lu = tl.lu + m;
    const int Pbeg = FWD ? 0 : pl.npanels - 1, dP = FWD ? 1 : -1;
    int ix[TU];
    double lv[TU][PR];
    double z[RPL], dinv[RPL], tri[RPL][PR - 1];
    int prm[RPL], nprm[RPL];
    auto col_begin = [&](const PanelMeta &q) { return FWD ? 0 : q.next + q.nr; };
    auto col_end = [&](const PanelMeta &q) { return FWD ? q.next : q.width; };
    auto load_perm = [&](const PanelMeta &q, int (&p)[RPL]) {
#pragma unroll
        for (int rr = 0; rr < RPL; ++rr) {
            const int row = r0 + rr * RL;
            p[rr] = row < q.nr ? net.perm[q.p0 + row] : 0;
        }
    };
    // one batch: columns c0 + j*LN (c0 already includes the lane's offset), all eight rows of each
    auto load_batch = [&](const PanelMeta &q, int c0, int ce) {
#pragma unroll
        for (int j = 0; j < TU; ++j) {
            const int c = c0 + j * LN;
            const bool ok = c < ce;
            ix[j] = ok ? pl.cols[q.cptr + c] : 0;
            const double *vp = lu + (q.base + c * q.nr) * MB;
#pragma unroll
            for (int r = 0; r < PR; ++r) lv[j][r] = (ok && r < q.nr) ? vp[r * MB] : 0.0;
        }
    };
    auto consume = [&](int c0, int ce, double (&acc)[PR]) {
        double yv[TU];
#pragma unroll
        for (int j = 0; j < TU; ++j) yv[j] = (c0 + j * LN < ce) ? y[ix[j] * MB + m] : 0.0;
#pragma unroll
        for (int j = 0; j < TU; ++j)
#pragma unroll
            for (int r = 0; r < PR; ++r) acc[r] += lv[j][r] * yv[j];
    };
    // everything of panel q that does not depend on the panels before it
    auto preload = [&](const PanelMeta &q, const int (&p)[RPL]) {
        load_batch(q, col_begin(q) + ln, col_end(q));
#pragma unroll
        for (int rr = 0; rr < RPL; ++rr) {
            const int row = r0 + rr * RL;
            const bool ok = row < q.nr;
            const double *dp = lu + (q.base + q.next * q.nr + row) * MB;      // row `row` of the diagonal block
            if (FWD) {
                z[rr] = ok ? rhs[p[rr] * MB + m] : 0.0;
                dinv[rr] = ok ? tl.invd[(q.p0 + row) * MB + m] : 0.0;
#pragma unroll
                for (int a = 0; a < PR - 1; ++a) tri[rr][a] = (ok && a < row) ? dp[a * q.nr * MB] : 0.0;
            } else {
                z[rr] = ok ? y[(q.p0 + row) * MB + m] : 0.0;
                dinv[rr] = 0.0;
#pragma unroll
                for (int a = 1; a < PR; ++a) tri[rr][a - 1] = (ok && a > row && a < q.nr) ? dp[a * q.nr * MB] : 0.0;
            }
        }
    };
    PanelMeta cur = load_pm(pl, Pbeg), nxt = load_pm(pl, Pbeg + dP), nn = load_pm(pl, Pbeg + 2 * dP);
    load_perm(cur, prm);
    preload(cur, prm);
    for (int P = Pbeg; FWD ? P < pl.npanels : P >= 0; P += dP) {
        const PanelMeta pm = cur;
        load_perm(nxt, nprm);
        if (KB2_TRI_AHEAD >= 0 && tl.lane == 0 && col_end(nn) > col_begin(nn)) {
            const size_t a = (size_t)(tl.lu + (size_t)(nn.base + col_begin(nn) * nn.nr) * MB);
            const size_t nbytes = (size_t)(col_end(nn) - col_begin(nn)) * nn.nr * MB * 8;
            const size_t a16 = a & ~(size_t)15;
            prefetch_l2_bulk((const void *)a16, (unsigned)((a + nbytes - a16) & ~(size_t)15));
        }
        double acc[PR];
#pragma unroll
        for (int r = 0; r < PR; ++r) acc[r] = 0.0;
        const int cb = col_begin(pm), ce = col_end(pm);
        consume(cb + ln, ce, acc);
        for (int cc = cb + LN * TU; cc < ce; cc += LN * TU) {       // wide panels: further batches on demand
            load_batch(pm, cc + ln, ce);
            consume(cc + ln, ce, acc);
        }
        // the finish of this panel keeps its own copies; the registers of the batch are free again
        double cz[RPL], cdinv[RPL], ctri[RPL][PR - 1];
        int cprm[RPL];
#pragma unroll
        for (int rr = 0; rr < RPL; ++rr) {
            cz[rr] = z[rr]; cdinv[rr] = dinv[rr]; cprm[rr] = prm[rr]; prm[rr] = nprm[rr];
#pragma unroll
            for (int a = 0; a < PR - 1; ++a) ctri[rr][a] = tri[rr][a];
        }
        cur = nxt; nxt = nn; nn = load_pm(pl, P + 3 * dP);
        preload(cur, prm);
        // ---- finish panel P: row totals, in-panel substitution with shuffles ----
        double tot[RPL];
        rows_reduce<MB>(acc, ln, tot);
#pragma unroll
        for (int rr = 0; rr < RPL; ++rr) cz[rr] -= tot[rr];
        if (FWD) {
#pragma unroll
            for (int a = 0; a < PR - 1; ++a) {
                const int oa = a % RL, sa = a / RL;                       // owner lane and slot of row a
                const double yv = __shfl_sync(FULL, cz[sa] * cdinv[sa], (cg * RL + oa) * MB + m);   // y_a is final here
#pragma unroll
                for (int rr = 0; rr < RPL; ++rr)
                    if (r0 + rr * RL > a) cz[rr] -= ctri[rr][a] * yv;
            }
#pragma unroll
            for (int rr = 0; rr < RPL; ++rr) {
                const int row = r0 + rr * RL;
                if (cg == 0 && row < pm.nr) y[(pm.p0 + row) * MB + m] = cz[rr] * cdinv[rr];
            }
        } else {
#pragma unroll
            for (int a = PR - 1; a > 0; --a) {
                const int oa = a % RL, sa = a / RL;
                const double xv = __shfl_sync(FULL, cz[sa], (cg * RL + oa) * MB + m);        // U' has a unit diagonal
#pragma unroll
                for (int rr = 0; rr < RPL; ++rr)
                    if (r0 + rr * RL < a) cz[rr] -= ctri[rr][a - 1] * xv;
            }
#pragma unroll
            for (int rr = 0; rr < RPL; ++rr) {
                const int row = r0 + rr * RL;
                if (cg == 0 && row < pm.nr) {
                    y[(pm.p0 + row) * MB + m] = cz[rr];
                    x[cprm[rr] * MB + m] = cz[rr];
                }
            }
        }
        __syncwarp();
    }
}

// ys: the warp's shared memory if S*MB doubles fit there (en.u_smem), else null (y stays in HBM)
template <int MB>
__device__ void tile_trisolve(const WTile<MB> &tl, const DevNet &net, const DevPlan &pl, const double *rhs, double *x, double *ys)
{
    __syncwarp();
    if (ys) {
        tri_sweep<MB, true, true>(tl, net, pl, rhs, x, ys);       // forward:  L' y = P rhs
        tri_sweep<MB, true, false>(tl, net, pl, rhs, x, ys);      // backward: U' x = y
    } else {
        tri_sweep<MB, false, true>(tl, net, pl, rhs, x, tl.y);
        tri_sweep<MB, false, false>(tl, net, pl, rhs, x, tl.y);
    }
}

}  // namespace kb2
