// Kernels that compose the tile primitives: stand-alone entry points and the fused solve.
#pragma once
#include "kb2_kernels.cuh"
#include "kb2_panel.cuh"

namespace kb2 {

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

template <int MB, int MINB>
__global__ void __launch_bounds__(32 * MB, MINB) k_factor(DevNet net, DevPlan pl, DevEns en, const double *hg_inv, int ntiles, int mode)
{
    extern __shared__ double smem[];
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        Tile<MB> tl(tile, en.Bp);
        if (mode & 1) tile_assemble_w(tl, net, en.u, en.k, hg_inv[tl.b], en.lu);
        __syncthreads();
        if (mode & 2) tile_lu_panels(tl, pl, en.lu, en.invd, smem);
        __syncthreads();
    }
}

template <int MB, int MINB>
__global__ void __launch_bounds__(32 * MB, MINB) k_trisolve(DevNet net, DevPlan pl, DevEns en, int ntiles)
{
    extern __shared__ double smem[];
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        Tile<MB> tl(tile, en.Bp);
        tile_trisolve_panels(tl, net, pl, en.lu, en.invd, en.rv, en.y, en.ua, smem);
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
__global__ void k_pack_bs(int S, int B, int Bp, const double *src, double *dst)
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
        else if (c.upd[m] && c.status[m] == ST_RUNNING) { c.h[m] = fmin(c.h[m], 0.1 * hn); c.nrhs[m] += 2; }   // 0.1: see DESIGN.md §3
    }
    __syncthreads();
}

template <int MB>
__device__ void solve_tile(int tile, const DevNet &net, const DevPlan &pl, const DevEns &en, Ctl<MB> &c, double *lbuf, double *red)
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
        tile_lu_panels(tl, pl, en.lu, en.invd, lbuf);
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
            tile_trisolve_panels(tl, net, pl, en.lu, en.invd, en.rv, en.y, en.K[s], red);
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

template <int MB, int MINB>
__global__ void __launch_bounds__(32 * MB, MINB) k_solve(DevNet net, DevPlan pl, DevEns en, int ntiles, int *tile_counter)
{
    extern __shared__ double smem[];
    __shared__ Ctl<MB> c;
    __shared__ int s_tile;
    double *lbuf = smem;
    double *red = smem + lu_smem_doubles(32 * MB, MB);
    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1);
        __syncthreads();
        const int tile = s_tile;
        __syncthreads();
        if (tile >= ntiles) break;
        solve_tile<MB>(tile, net, pl, en, c, lbuf, red);
    }
}

}  // namespace kb2
