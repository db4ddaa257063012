// Kernels that compose the tile primitives: stand-alone entry points (kernel-level C ABI, timing)
// and the phase kernels of the solve.  A tile of MB members is worked on by one warp (sweeps,
// block-plan LU) or by NW warps of one CTA (right-hand side, Jacobian values, step end).
#pragma once
#include "kb2_kernels.cuh"

namespace kb2 {

__host__ __device__ constexpr size_t lu_smem_bytes(int mb) { return lu_smem_doubles(mb) * sizeof(double); }
// Shared memory of a warp: what the factorisation needs, or the tile's state vector if that is
// larger and still leaves seven warps per SM (232448 bytes / 7, minus 1 KB the system reserves)
inline size_t warp_smem_bytes(int mb, int64_t S, bool *u_fits)
{
    const size_t base = lu_smem_bytes(mb), ub = (size_t)S * mb * sizeof(double), cap = 232448 / 7 - 1024;
    const size_t smem = (ub > base && ub + 16 <= cap) ? ub : base;
    if (u_fits) *u_fits = ub <= smem;
    return smem;            // data bytes; kernels are launched with 16 more for the copy channel's mbarrier
}

// ---------------------------------------------------------------------------------------------
// Layout conversion between the caller's row-major [row][B] arrays and the tile-major device
// layout [tile][row][MB]
// ---------------------------------------------------------------------------------------------
template <int MB>
__global__ void k_rows_to_tiles(const double *rows, double *tiles, size_t nrows, int B, int Bt, size_t row_stride)
{
    const size_t n = nrows * (size_t)Bt;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t i = idx / Bt;
        const int b = (int)(idx % Bt);
        const double v = b < B ? rows[i * row_stride + (row_stride ? b : 0)] : 0.0;
        tiles[((size_t)(b / MB) * nrows + i) * MB + b % MB] = v;
    }
}

template <int MB>
__global__ void k_tiles_to_rows(const double *tiles, double *rows, size_t nrows, int B)
{
    const size_t n = nrows * (size_t)B;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t i = idx / B;
        const int b = (int)(idx % B);
        rows[idx] = tiles[((size_t)(b / MB) * nrows + i) * MB + b % MB];
    }
}

// tile-major [tile][S][MB] -> member-major [B][S] pack for the allgather
template <int MB>
__global__ void k_pack_bs(int S, int B, const double *tiles, double *dst)
{
    const size_t n = (size_t)S * B;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(idx / S), i = (int)(idx % S);
        dst[idx] = tiles[((size_t)(b / MB) * S + i) * MB + b % MB];
    }
}

// ---------------------------------------------------------------------------------------------
// Stand-alone kernels (kernel-level C-ABI entry points, per-kernel roofline timing)
// ---------------------------------------------------------------------------------------------
template <int MB>
__global__ void __launch_bounds__(32) k_rates(DevNet net, DevPlan pl, DevEns en, const double *T, int ntiles)
{
    BulkChan ch; ch.bar = 0; ch.par = nullptr;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        WTile<MB> tl(tile, net, pl, en, ch);
        tile_rates(tl, net, T[tl.b], true, -1);
    }
}

template <int MB, int NW>
__global__ void __launch_bounds__(NW * 32, KB2_RHS_MINB) k_rhs(DevNet net, DevPlan pl, DevEns en, int ntiles)
{
    extern __shared__ double smem[];
    BulkChan ch; ch.bar = 0; ch.par = nullptr;
    const int w = threadIdx.x >> 5;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        WTile<MB> tl(tile, net, pl, en, ch);
        tile_rhs<MB, NW>(tl, net, tl.u, tl.rv, false, en.u_smem ? smem : nullptr, w);
    }
}

template <int MB>
__global__ void __launch_bounds__(32) k_factor(DevNet net, DevPlan pl, DevEns en, const double *hg_inv, int ntiles, int mode, int data_bytes)
{
    extern __shared__ double smem[];
    const BulkChan ch = chan_setup<MB>(smem, data_bytes);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        WTile<MB> tl(tile, net, pl, en, ch);
        if (mode & 1) tile_assemble_w(tl, net, pl, tl.u, hg_inv[tl.b], en.u_smem ? smem : nullptr);
        if (mode & 2) tile_lu(tl, pl, smem);
    }
}

template <int MB>
__global__ void __launch_bounds__(32) k_trisolve(DevNet net, DevPlan pl, DevEns en, int ntiles, int data_bytes)
{
    extern __shared__ double smem[];
    const BulkChan ch = chan_setup<MB>(smem, data_bytes);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        WTile<MB> tl(tile, net, pl, en, ch);
        tile_trisolve(tl, net, pl, tl.rv, tl.ua, en.u_smem ? smem : nullptr);
    }
}

__global__ void k_profile(int B, int nt, const int *kind, const double *params, const double *t, double *X)
{
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * nt) return;
    int b = idx / nt, s = idx % nt;
    X[idx] = profile_eval(kind[b], params + (size_t)b * 16, t[s]);
}

// ---------------------------------------------------------------------------------------------
// The solve: Rodas4 with per-member adaptive h, as a sequence of PHASE KERNELS per attempted step
//     k_step_jac | k_lu_window | 6 x ( k_stage_rhs(s) | k_stage_sweep(s) ) | k_step_end
// (k_step_lu = block-plan assembly + LU instead of the first two when the window does not fit)
// launched back to back on one stream by the host loop of kb2_solve_run.  Every kernel walks the
// tiles of the ensemble (one warp-tile of MB members at a time); the control state of a member
// (Ctl) lives in global memory between kernels and is replicated in the registers of the member's
// lanes inside one (same inputs, same arithmetic, same decisions on every lane).  Accept/reject
// and stop handling are masked per member while a tile moves in lock-step; a tile without a
// running member is skipped by every kernel.  Kernel boundaries keep all tiles in the same phase
// of the step (each phase has its own code and arrays: instruction cache, L2 and TLB see one
// working set at a time) and give every phase its own launch shape, registers and shared memory.
// Replaces init/solve!/reinit! of `pars.solver` and the PresetTimeCallback rate update
// (reference src/solving/methods.jl:655-714, src/solving/solve_utils.jl:376-450).
// ---------------------------------------------------------------------------------------------
enum { ST_RUNNING = -1 };
constexpr int KB2_FLAG_SLOTS = 64;
constexpr int END_NW = 4;          // warps per tile of k_solve_init / k_step_end

// Sum over all lanes of a member when NW warps of a CTA share the tile: warp butterfly, then the
// warps' partial sums through shared memory in warp order (deterministic; every lane of the member
// ends up with the same bits).  red: NW*MB doubles of shared memory.
template <int MB, int NW>
__device__ __forceinline__ double member_sum_cta(double v, double *red, int w, int m, int ln)
{
    v = member_sum<MB>(v);
    if (NW == 1) return v;
    if (ln == 0) red[w * MB + m] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < NW; ++q) t += red[q * MB + m];
    __syncthreads();
    return t;
}

template <int MB, int NW>
__device__ void tile_process_stop(const WTile<MB> &tl, const DevNet &net, const DevEns &en, Ctl &c, bool at_start, int w)
{
    constexpr int LN = 32 / MB, VL = LN * NW;
    const int vl = w * LN + tl.ln;
    c.upd = 0; c.sav = -1; c.chunk = 0;
    const size_t sb = (size_t)tl.b * en.nstops;
    const bool due = at_start ? (c.status == ST_RUNNING && c.si < c.ns && en.stop_t[sb + c.si] <= en.t0)
                              : (c.accept && c.hit);
    if (due) {
        const int s = c.si, fl = en.stop_flags[sb + s];
        if (fl & 1) {
            double T = en.Ttab ? en.Ttab[sb + s] : nan("");
            if (isnan(T) && net.calc_mode == 0) T = profile_eval(en.pkind[tl.b], en.pparams + (size_t)tl.b * 16, en.stop_t[sb + s]);
            c.T = T; c.upd = 1; c.ridx = en.stop_ridx[sb + s];
        }
        if (fl & 2) c.sav = c.isave++;
        if (fl & 4) c.chunk = 1;
        c.si = s + 1;
        if (c.si >= c.ns) c.status = 0;   // reached the end of tspan
    }
    if (__any_sync(FULL, c.upd)) tile_rates<MB, NW>(tl, net, c.T, c.upd != 0, c.ridx, w);
    const int sv = c.sav;
    if (__any_sync(FULL, sv >= 0)) {
        if (sv >= 0)
            for (int i = vl; i < net.S; i += VL) tl.out_u[((size_t)sv * net.S + i) * MB + tl.m] = tl.u[i * MB + tl.m];
        tile_sync<NW>();
    }
    // a chunk starts here (reference chunkwise solves, methods.jl:185-303, 717-865: the integrator is
    // re-initialised at every multiple of solve_chunkstep): iteration count and tolerances start
    // afresh, and the state is kept for a repeat of the chunk (adaptive_solve! per chunk)
    const bool cstart = (at_start || c.chunk) && c.status == ST_RUNNING;
    if (cstart) {
        c.iters = 0; c.retries = 0;
        c.tchunk = c.t; c.Tchunk = c.T; c.ridx_chunk = c.ridx; c.si_chunk = c.si; c.isave_chunk = c.isave;
        if (!en.update_tols) { c.atol = en.abstol; c.rtol = en.reltol; }
    }
    if (en.chunk_retry && __any_sync(FULL, cstart)) {
        if (cstart)
            for (int i = vl; i < net.S; i += VL) tl.uc[i * MB + tl.m] = tl.u[i * MB + tl.m];
        tile_sync<NW>();
    }
}

// Starting step size (Hairer-Nørsett-Wanner II.4, order 4).  Called once at t0 (`initial`) and
// again after a discrete rate update, where the RHS jumps, for members that have no step-size
// memory yet: they get h = min(h, 0.1 * estimate).  Uses rv, ua, y as scratch.
template <int MB, int NW>
__device__ void tile_hinit(const WTile<MB> &tl, const DevNet &net, const DevEns &en, Ctl &c, bool initial, double *su, double *red, int w)
{
    constexpr int LN = 32 / MB, VL = LN * NW;
    const int m = tl.m, vl = w * LN + tl.ln;
    tile_rhs<MB, NW>(tl, net, tl.u, tl.rv, false, su, w);
    double d0 = 0, d1 = 0;
    for (int i = vl; i < net.S; i += VL) {
        const double ui = tl.u[i * MB + m], fi = tl.rv[i * MB + m];
        const double sc = c.atol + c.rtol * fabs(ui);
        d0 += (ui / sc) * (ui / sc); d1 += (fi / sc) * (fi / sc);
    }
    d0 = sqrt(member_sum_cta<MB, NW>(d0, red, w, m, tl.ln) / net.S);
    d1 = sqrt(member_sum_cta<MB, NW>(d1, red, w, m, tl.ln) / net.S);
    const double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
    for (int i = vl; i < net.S; i += VL) tl.ua[i * MB + m] = tl.u[i * MB + m] + h0 * tl.rv[i * MB + m];
    tile_sync<NW>();
    tile_rhs<MB, NW>(tl, net, tl.ua, tl.y, false, su, w);
    double d2 = 0;
    for (int i = vl; i < net.S; i += VL) {
        const double sc = c.atol + c.rtol * fabs(tl.u[i * MB + m]);
        const double q = (tl.y[i * MB + m] - tl.rv[i * MB + m]) / sc;
        d2 += q * q;
    }
    d2 = sqrt(member_sum_cta<MB, NW>(d2, red, w, m, tl.ln) / net.S) / h0;
    const double dm = fmax(d1, d2);
    const double h1 = (dm <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(0.01 / dm, 0.2);
    const double hn = fmin(100.0 * h0, h1);
    if (initial) { c.h = hn; c.hold = hn; c.nrhs += 2; }
    else if ((c.upd || c.chunk) && c.status == ST_RUNNING && !(c.hfirst > 0.0)) { c.h = fmin(c.h, 0.1 * hn); c.nrhs += 2; }
    tile_sync<NW>();
}

// Step size after a discrete rate update.  The jump in k throws the fast species off their
// quasi-steady state, so the step must drop to the fast time scale again; consecutive updates of
// a profile are alike, so the step size that the controller found optimal right after the
// previous update (c.hfirst) is the prediction for this one.  Members without that memory fall
// back to the starting-step estimate.
template <int MB, int NW>
__device__ void tile_restart_h(const WTile<MB> &tl, const DevNet &net, const DevEns &en, Ctl &c, double *su, double *red, int w)
{
    const bool upd = (c.upd || c.chunk) && c.status == ST_RUNNING;      // a chunk boundary re-initialises the integrator as well
    if (upd) { c.fresh = 1; c.firstacc = 1; }
    if (__any_sync(FULL, upd && !(c.hfirst > 0.0))) tile_hinit<MB, NW>(tl, net, en, c, false, su, red, w);
    if (upd && c.hfirst > 0.0) c.h = fmin(c.h, c.hfirst);
}

// What the next attempted step of a member is: its size hs (cut at the next stop), whether it
// ends on the stop, or why the member stops running (maxiters, dtmin: methods.jl:164-165).
__device__ __forceinline__ void plan_attempt(const DevEns &en, Ctl &c, int b)
{
    int act = (c.status == ST_RUNNING);
    double hs = 1.0;
    int hit = 0;
    if (act) {
        if (++c.iters > en.maxiters) { c.status = 1; act = 0; }
        else {
            const double tstop = en.stop_t[(size_t)b * en.nstops + c.si];
            hs = c.h;
            if (c.t + 1.01 * hs >= tstop) { hs = tstop - c.t; hit = 1; }
            if (hs < en.dtmin && !hit) { c.status = 2; act = 0; hs = 1.0; }
        }
    }
    c.active = act; c.hs = hs; c.hit = hit; c.accept = 0;
}

// adaptive_solve! per chunk (solve_utils.jl:376-424 inside the chunk loops of methods.jl): a member
// whose chunk failed (maxiters, dtmin) goes back to the start of the chunk with abstol / reltol x0.1,
// at most five attempts and never below eps; then the failure stands.
template <int MB, int NW>
__device__ void tile_chunk_retry(const WTile<MB> &tl, const DevNet &net, const DevEns &en, Ctl &c, double *su, double *red, int w)
{
    constexpr int LN = 32 / MB, VL = LN * NW;
    const double mintol = 2.220446049250313e-16;
    const bool redo = (c.status == 1 || c.status == 2 || c.status == 3) && c.retries < 4 && c.atol / 10 > mintol && c.rtol / 10 > mintol;
    if (!__any_sync(FULL, redo)) return;
    if (redo) {
        for (int i = w * LN + tl.ln; i < net.S; i += VL) tl.u[i * MB + tl.m] = tl.uc[i * MB + tl.m];
        c.t = c.tchunk; c.si = c.si_chunk; c.isave = c.isave_chunk; c.T = c.Tchunk; c.ridx = c.ridx_chunk;
        c.atol /= 10; c.rtol /= 10; c.retries++; c.nretry++;
        c.iters = 0; c.status = ST_RUNNING;
        c.rejlast = 0; c.firstacc = 1; c.fresh = 0; c.errold = 1.0; c.hfirst = 0.0;
    }
    tile_sync<NW>();
    tile_rates<MB, NW>(tl, net, c.T, redo, c.ridx, w);          // the rate constants the chunk started with
    {
        // fresh starting step for the repeated members only
        Ctl tmp = c;
        tile_hinit<MB, NW>(tl, net, en, tmp, true, su, red, w);
        if (redo) { c.h = tmp.h; c.hold = tmp.hold; c.nrhs += 2; }
    }
    if (redo) plan_attempt(en, c, tl.b);
}

// end of a kernel that changes the control state: one lane per member writes it back, results of
// finished members are published, and the tile reports whether it still has work
template <int MB>
__device__ __forceinline__ void store_ctl(const WTile<MB> &tl, const DevEns &en, const Ctl &c, int slot, int w = 0)
{
    if (w != 0) return;
    if (tl.ln == 0) {
        en.ctl[tl.b] = c;
        if (tl.b < en.B && c.status != ST_RUNNING) {
            en.status[tl.b] = c.status;
            long long *st = en.stats + (size_t)tl.b * 8;
            st[0] = c.nacc; st[1] = c.nrej; st[2] = c.nlu; st[3] = c.nrhs;
            st[4] = c.isave; st[5] = c.si; st[6] = c.iters; st[7] = c.nretry;
        }
    }
    if (__any_sync(FULL, c.active) && tl.lane == 0) en.flags[slot] = 1;
}

template <int MB, int NW>
__global__ void __launch_bounds__(NW * 32, KB2_RHS_MINB) k_solve_init(DevNet net, DevPlan pl, DevEns en, int ntiles)
{
    extern __shared__ double smem[];
    BulkChan ch; ch.bar = 0; ch.par = nullptr;
    double *su = en.u_smem ? smem : nullptr;
    double *red = smem + (en.u_smem ? (size_t)net.S * MB : 0);
    const int w = threadIdx.x >> 5;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        WTile<MB> tl(tile, net, pl, en, ch);
        const int b = tl.b;
        Ctl c;
        c.t = en.t0; c.si = 0; c.isave = 0; c.iters = 0;
        c.ns = en.stop_cnt[b];
        c.status = (b < en.B && c.ns > 0) ? ST_RUNNING : 0;
        c.nacc = c.nrej = c.nlu = c.nrhs = 0;
        c.rejlast = 0; c.firstacc = 1; c.accept = 0; c.hit = 0; c.active = 0; c.upd = 0; c.ridx = -1; c.sav = -1;
        c.errold = 1.0; c.h = 0.0; c.hs = 1.0; c.hold = 0.0; c.hfirst = 0.0; c.fresh = 0;
        c.atol = en.abstol; c.rtol = en.reltol; c.tchunk = en.t0; c.Tchunk = 0.0;
        c.si_chunk = 0; c.isave_chunk = 0; c.ridx_chunk = -1; c.retries = 0; c.chunk = 0; c.nretry = 0;
        // initial conditions: static -> value, variable -> X_start (condition_set.jl:111-121); both sit
        // in the profile's X(0) for every supported kind
        c.T = (net.calc_mode == 0) ? profile_eval(en.pkind[b], en.pparams + (size_t)b * 16, -1.0) : 0.0;
        tile_rates<MB, NW>(tl, net, c.T, true, -1, w);     // k(initial conditions), methods.jl:668
        tile_process_stop<MB, NW>(tl, net, en, c, true, w);
        tile_hinit<MB, NW>(tl, net, en, c, true, su, red, w);
        plan_attempt(en, c, b);
        store_ctl(tl, en, c, 0, w);
        __syncthreads();
    }
}

// W = I/(hs*gamma) - J(u) assembled and factorised
template <int MB>
__global__ void __launch_bounds__(32) k_step_lu(DevNet net, DevPlan pl, DevEns en, int ntiles, int data_bytes)
{
    extern __shared__ double smem[];
    const BulkChan ch = chan_setup<MB>(smem, data_bytes);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        WTile<MB> tl(tile, net, pl, en, ch);
        const Ctl *c = en.ctl + tl.b;
        if (!__any_sync(FULL, c->active)) continue;
        if (en.continuous)      // k(T(t)) at the start of the step (see k_step_jac)
            tile_rates(tl, net, profile_eval(en.pkind[tl.b], en.pparams + (size_t)tl.b * 16, c->t), c->active != 0, -1);
        tile_assemble_w(tl, net, pl, tl.u, 1.0 / (c->hs * kGamma), en.u_smem ? smem : nullptr);
        tile_lu(tl, pl, smem);
    }
}

// stage s: argument ua = u + sum_{q<s} a_sq K_q (stage 6 shares stage 5's coefficients plus K5) and,
// from the same loads, the stage combination rv = sum_{q<s} (C_sq/h) K_q; then rv += f(ua).
// NW warps of one CTA share a tile: the passes are streams with short dependent chains, so what
// they need is many warps in flight per SM (NW = 4: 28 resident warps instead of 7 on C3).
// RS: the tile's rate table (R*MB doubles) lives in shared memory next to the state vector, so the
// rates never travel to HBM and back between the per-reaction pass and the gather (networks whose
// table fits: one 16-warp CTA per SM).
constexpr int RHS_NW = 4, RHS_NW_RS = 16;
template <int MB, int NW, bool RS>
__global__ void __launch_bounds__(NW * 32, RS ? 1 : KB2_RHS_MINB) k_stage_rhs(DevNet net, DevPlan pl, DevEns en, int ntiles, int s)
{
    extern __shared__ double smem[];
    constexpr int LN = 32 / MB, VL = LN * NW;
    BulkChan ch; ch.bar = 0; ch.par = nullptr;
    const int w = threadIdx.x >> 5;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        WTile<MB> tl(tile, net, pl, en, ch);
        if (RS) tl.rate = smem + (size_t)net.S * MB;
        const Ctl *c = en.ctl + tl.b;
        if (!__syncthreads_or(c->active)) continue;
        const int m = tl.m, vl = w * LN + tl.ln;
        const double *Us = tl.u;
        if (en.continuous) {
            // k at the stage time t + c_s h from the member's own profile; at the first stage also
            // dk/dt and from it df/dt = f(u; dk/dt) (mass action is linear in k), which enters the
            // first four stages with the weights h*d_s
            const double ts = c->t + cT[s] * c->hs;
            const int kind = en.pkind[tl.b];
            const double *pp = en.pparams + (size_t)tl.b * 16;
            const double Ts = profile_eval(kind, pp, ts);
            double *kd = en.kdot + (size_t)tile * net.R * MB, *ft = en.ft + (size_t)tile * net.S * MB;
            if (s == 0) {
                tile_rates<MB, NW>(tl, net, Ts, c->active != 0, -1, w, kd, profile_grad(kind, pp, ts));
                WTile<MB> tk = tl;
                tk.k = kd;
                tile_rhs<MB, NW>(tk, net, tl.u, ft, false, en.u_smem ? smem : nullptr, w);
            } else if (s < 5) {
                tile_rates<MB, NW>(tl, net, Ts, c->active != 0, -1, w);
            }
        }
        if (s > 0) {
            const double ih = 1.0 / c->hs;
            const double a0 = cA[s][0], a1 = cA[s][1], a2 = cA[s][2], a3 = cA[s][3], a4 = cA[s][4];
            const double c0 = cC[s][0] * ih, c1 = cC[s][1] * ih, c2 = cC[s][2] * ih, c3 = cC[s][3] * ih, c4 = cC[s][4] * ih;
#pragma unroll 4
            for (int i = vl; i < net.S; i += VL) {
                const int o = i * MB + m;
                const double k0 = tl.K[0][o];
                double a = tl.u[o] + a0 * k0, r = c0 * k0;
                if (s > 1) { const double kq = tl.K[1][o]; a += a1 * kq; r += c1 * kq; }
                if (s > 2) { const double kq = tl.K[2][o]; a += a2 * kq; r += c2 * kq; }
                if (s > 3) { const double kq = tl.K[3][o]; a += a3 * kq; r += c3 * kq; }
                if (s > 4) { const double kq = tl.K[4][o]; a += a4 * kq; r += c4 * kq; }
                tl.ua[o] = a;
                tl.rv[o] = r;
            }
            __syncthreads();
            Us = tl.ua;
        }
        tile_rhs<MB, NW>(tl, net, Us, tl.rv, s > 0, en.u_smem ? smem : nullptr, w);
        if (en.continuous && s < 4) {
            const double hd = c->hs * cD[s];
            const double *ft = en.ft + (size_t)tile * net.S * MB;
            for (int i = vl; i < net.S; i += VL) tl.rv[i * MB + m] += hd * ft[i * MB + m];
            __syncthreads();
        }
    }
}

// stage s: K_s = W^-1 rv
template <int MB>
__global__ void __launch_bounds__(32) k_stage_sweep(DevNet net, DevPlan pl, DevEns en, int ntiles, int data_bytes, int s)
{
    extern __shared__ double smem[];
    const BulkChan ch = chan_setup<MB>(smem, data_bytes);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        WTile<MB> tl(tile, net, pl, en, ch);
        if (!__any_sync(FULL, en.ctl[tl.b].active)) continue;
        tile_trisolve(tl, net, pl, tl.rv, tl.K[s], en.u_smem ? smem : nullptr);
    }
}

// error estimate (= K6), controller, commit, stop handling (rate update, saves), and the plan of
// the next attempt
template <int MB, int NW>
__global__ void __launch_bounds__(NW * 32, KB2_RHS_MINB) k_step_end(DevNet net, DevPlan pl, DevEns en, int ntiles, int slot)
{
    extern __shared__ double smem[];
    constexpr int LN = 32 / MB, VL = LN * NW;
    BulkChan ch; ch.bar = 0; ch.par = nullptr;
    double *su = en.u_smem ? smem : nullptr;
    double *red = smem + (en.u_smem ? (size_t)net.S * MB : 0);
    const int w = threadIdx.x >> 5;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        WTile<MB> tl(tile, net, pl, en, ch);
        if (!__syncthreads_or(en.ctl[tl.b].active)) continue;
        Ctl c = en.ctl[tl.b];
        const int m = tl.m, ln = tl.ln, vl = w * LN + ln;
        const double hs = c.hs;
        const size_t sb = (size_t)tl.b * en.nstops;
        double e2 = 0.0;
        int neg = 0;
        for (int i = vl; i < net.S; i += VL) {
            const int o = i * MB + m;
            const double k6 = tl.K[5][o], un = tl.ua[o] + k6;
            const double sc = c.atol + c.rtol * fmax(fabs(tl.u[o]), fabs(un));
            e2 += (k6 / sc) * (k6 / sc);
            neg |= (un < 0.0);
        }
        double err = sqrt(member_sum_cta<MB, NW>(e2, red, w, m, ln) / net.S);
        const double nneg = en.ban_neg ? member_sum_cta<MB, NW>((double)neg, red, w, m, ln) : 0.0;
        if (!(err < INFINITY)) err = INFINITY;          // NaN/Inf (singular pivot, overflow) -> reject
        if (nneg > 0.0) err = fmax(err, 1e4);            // isoutofdomain, methods.jl:169-171
        if (c.active) {
            double fac = (err < INFINITY) ? fmax(1.0 / 6.0, fmin(5.0, pow(err, 0.25) / 0.9)) : 5.0;
            double hnew = hs / fac;
            c.nlu++; c.nrhs += 6;
            if (err <= 1.0) {
                c.nacc++;
                if (!c.firstacc) {
                    double facgus = (c.hold / hs) * pow(err * err / c.errold, 0.25) / 0.9;
                    facgus = fmax(1.0 / 6.0, fmin(5.0, facgus));
                    fac = fmax(fac, facgus);
                    hnew = hs / fac;
                }
                if (c.fresh) { c.hfirst = hs / fmax(1.0 / 6.0, fmin(5.0, pow(err, 0.25) / 0.9)); c.fresh = 0; }
                c.firstacc = 0;
                c.hold = hs; c.errold = fmax(1e-2, err);
                if (c.rejlast) hnew = fmin(hnew, hs);
                c.rejlast = 0;
                c.accept = 1;
                if (c.hit) { c.t = en.stop_t[sb + c.si]; c.h = fmax(hnew, c.h); }
                else { c.t += hs; c.h = hnew; }
            } else {
                c.nrej++; c.rejlast = 1; c.h = hnew;
                if (hnew < en.dtmin) c.status = (err < INFINITY) ? 2 : 3;      // 3: the last error norm was not finite (singular pivot, overflow)
            }
        }
        if (c.accept)
            for (int i = vl; i < net.S; i += VL) {
                const int o = i * MB + m;
                tl.u[o] = tl.ua[o] + tl.K[5][o];
            }
        tile_sync<NW>();
        tile_process_stop<MB, NW>(tl, net, en, c, false, w);
        if (__any_sync(FULL, c.upd || c.chunk)) tile_restart_h<MB, NW>(tl, net, en, c, su, red, w);
        plan_attempt(en, c, tl.b);
        if (en.chunk_retry) tile_chunk_retry<MB, NW>(tl, net, en, c, su, red, w);
        store_ctl(tl, en, c, slot, w);
        __syncthreads();
    }
}

// per-species maxima over the save points (what identify_next_seeds consumes, explore_utils.jl:344-349)
template <int MB>
__global__ void k_umax(DevEns en, int S, size_t n)
{
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        // idx = (tile * S + i) * MB + m
        const size_t tile = idx / ((size_t)S * MB), rem = idx % ((size_t)S * MB);
        const double *src = en.out_u + tile * en.Ns * S * MB + rem;
        double mx = src[0];
        for (int s = 1; s < en.Ns; ++s) mx = fmax(mx, src[(size_t)s * S * MB]);
        en.out_umax[idx] = mx;
    }
}

// members the loop gave up on (host-side round limit): status 5
__global__ void k_mark_unfinished(DevEns en)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= en.B) return;
    const Ctl &c = en.ctl[b];
    if (c.status == ST_RUNNING) {
        en.status[b] = 5;
        long long *st = en.stats + (size_t)b * 8;
        st[0] = c.nacc; st[1] = c.nrej; st[2] = c.nlu; st[3] = c.nrhs; st[4] = c.isave; st[5] = c.si; st[6] = c.iters; st[7] = 0;
    }
}

}  // namespace kb2
