/* CPU ORACLE (plain C) — test infrastructure only, never the product path.
 *
 * Restates, for one ensemble member at a time, the kinetic-solve hot path of
 * Kinetica.jl v0.7.2 behind solve_network(::VariableODESolve, Val(:complete),
 * Val(:discrete)) (reference src/solving/methods.jl:655-714):
 *   - Arrhenius rate constants      reference src/solving/calculator.jl:223-232
 *   - mass-action RHS / Jacobian    reference src/solving/solve_utils.jl:318-334
 *                                   (Catalyst, combinatoric_ratelaws=false)
 *   - zero-order-hold rate updates  reference src/solving/solve_utils.jl:445-450
 *   - step control knobs            reference src/solving/methods.jl:160-171
 * integrated with the same algorithm the CUDA path uses (Rodas4, Hairer &
 * Wanner's RODAS coefficients, sparse LU over a shared symbolic factorisation)
 * so the two can be compared step for step.  The reference delegates the
 * implicit step to `pars.solver` (a user-chosen DifferentialEquations.jl
 * algorithm), so there is no reference step sequence to mirror: PARITY
 * UNPINNED at this boundary; trajectories are additionally checked against
 * scipy Radau/BDF in oracle/kinetica_oracle.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline /
 * --impl reference) may load this library.
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -shared -fPIC)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define R_GAS 8.314462618      /* reference src/constants.jl:4 */
#define N_AVO 6.02214076e23    /* reference src/constants.jl:5 */

/* k_r = A*exp(-Ea/(R*T))*N_A*t_mult ; optional harmonic cap with k_max
 * (reference src/solving/calculator.jl:223-232, same operation order). */
void ko_arrhenius(int64_t R, const double *A, const double *Ea, double T, double k_max,
                  double t_mult, double *k)
{
    for (int64_t r = 0; r < R; ++r) {
        double kr = A[r] * exp(-Ea[r] / (R_GAS * T)) * N_AVO * t_mult;
        k[r] = isnan(k_max) ? kr : 1.0 / ((1.0 / k_max) + (1.0 / kr));
    }
}

static double ipow(double x, int64_t n)
{
    double r = 1.0;
    for (int64_t i = 0; i < n; ++i) r *= x;
    return r;
}

typedef struct {
    int64_t S, R;
    const int64_t *rp, *ri, *rn;   /* reactants CSR  */
    const int64_t *pp, *pi, *pn;   /* products  CSR  */
    int64_t *np_, *ni, *nn;        /* net stoichiometry CSR (species ascending, zeros dropped) */
} net_t;

/* net[i,j] = sum(stoic_prods) - sum(stoic_reacs); a species on both sides nets out
 * (SURVEY.md §8a R5). */
static void net_build(net_t *n)
{
    int64_t cap = n->rp[n->R] + n->pp[n->R];
    n->np_ = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n->R + 1));
    n->ni = (int64_t *)malloc(sizeof(int64_t) * (size_t)(cap + 1));
    n->nn = (int64_t *)malloc(sizeof(int64_t) * (size_t)(cap + 1));
    int64_t w = 0;
    n->np_[0] = 0;
    for (int64_t j = 0; j < n->R; ++j) {
        int64_t start = w;
        for (int side = 0; side < 2; ++side) {
            const int64_t *p = side ? n->pp : n->rp, *ix = side ? n->pi : n->ri, *nu = side ? n->pn : n->rn;
            for (int64_t e = p[j]; e < p[j + 1]; ++e) {
                int64_t q = start;
                while (q < w && n->ni[q] != ix[e]) ++q;
                if (q == w) { n->ni[w] = ix[e]; n->nn[w] = 0; ++w; }
                n->nn[q] += side ? nu[e] : -nu[e];
            }
        }
        /* drop zeros, sort ascending by species */
        int64_t m = start;
        for (int64_t q = start; q < w; ++q)
            if (n->nn[q] != 0) { n->ni[m] = n->ni[q]; n->nn[m] = n->nn[q]; ++m; }
        w = m;
        for (int64_t a = start + 1; a < w; ++a)
            for (int64_t c = a; c > start && n->ni[c - 1] > n->ni[c]; --c) {
                int64_t t1 = n->ni[c]; n->ni[c] = n->ni[c - 1]; n->ni[c - 1] = t1;
                int64_t t2 = n->nn[c]; n->nn[c] = n->nn[c - 1]; n->nn[c - 1] = t2;
            }
        n->np_[j + 1] = w;
    }
}
static void net_free(net_t *n) { free(n->np_); free(n->ni); free(n->nn); }

/* neg_guard: rate laws are evaluated on max(u, 0) — the Lipschitz extension of mass action outside
 * the non-negative orthant used by the B200 solver (identical wherever the true solution lives). */
static int g_neg_guard = 0;
void ko_set_neg_guard(int on) { g_neg_guard = on; }
static inline double gu(double x) { return (g_neg_guard && x < 0.0) ? 0.0 : x; }

static double rate_of(const net_t *n, int64_t j, const double *u, const double *k)
{
    double r = k[j];
    for (int64_t e = n->rp[j]; e < n->rp[j + 1]; ++e) r *= ipow(gu(u[n->ri[e]]), n->rn[e]);
    return r;
}

/* du_i = sum_j net[i,j]*rate_j, accumulated in ascending j. */
static void rhs_eval(const net_t *n, const double *u, const double *k, double *du)
{
    memset(du, 0, sizeof(double) * (size_t)n->S);
    for (int64_t j = 0; j < n->R; ++j) {
        double r = rate_of(n, j, u, k);
        for (int64_t e = n->np_[j]; e < n->np_[j + 1]; ++e) du[n->ni[e]] += (double)n->nn[e] * r;
    }
}

void ko_rhs(int64_t S, int64_t R, const int64_t *rp, const int64_t *ri, const int64_t *rn,
            const int64_t *pp, const int64_t *pi, const int64_t *pn, const double *u,
            const double *k, double *du)
{
    net_t n = {S, R, rp, ri, rn, pp, pi, pn, 0, 0, 0};
    net_build(&n);
    rhs_eval(&n, u, k, du);
    net_free(&n);
}

/* Dense column-major-free Jacobian accumulate: J[i,l] += net[i,j]*d(rate_j)/du_l,
 * written through `put(i,l,val)`; species listed twice on one side are merged. */
typedef void (*put_fn)(void *ctx, int64_t i, int64_t l, double v);

static void jac_accumulate(const net_t *n, const double *u, const double *k, put_fn put, void *ctx)
{
    for (int64_t j = 0; j < n->R; ++j) {
        for (int64_t a = n->rp[j]; a < n->rp[j + 1]; ++a) {
            int64_t l = n->ri[a];
            int dup = 0;                      /* merged duplicate listing of l: handle at first hit */
            int64_t nu_l = 0;
            for (int64_t b = n->rp[j]; b < n->rp[j + 1]; ++b)
                if (n->ri[b] == l) { if (b < a) dup = 1; nu_l += n->rn[b]; }
            if (dup) continue;
            double d = k[j] * (double)nu_l * ipow(gu(u[l]), nu_l - 1);
            if (g_neg_guard && u[l] < 0.0) d = 0.0;
            for (int64_t b = n->rp[j]; b < n->rp[j + 1]; ++b)
                if (n->ri[b] != l) d *= ipow(gu(u[n->ri[b]]), n->rn[b]);
            for (int64_t e = n->np_[j]; e < n->np_[j + 1]; ++e) put(ctx, n->ni[e], l, (double)n->nn[e] * d);
        }
    }
}

typedef struct { const int64_t *colptr, *rowval; double *val; } csc_ctx;
static void put_csc(void *c, int64_t i, int64_t l, double v)
{
    csc_ctx *x = (csc_ctx *)c;
    for (int64_t p = x->colptr[l]; p < x->colptr[l + 1]; ++p)
        if (x->rowval[p] == i) { x->val[p] += v; return; }
}

void ko_jac_csc(int64_t S, int64_t R, const int64_t *rp, const int64_t *ri, const int64_t *rn,
                const int64_t *pp, const int64_t *pi, const int64_t *pn, const int64_t *colptr,
                const int64_t *rowval, const double *u, const double *k, double *Jval)
{
    net_t n = {S, R, rp, ri, rn, pp, pi, pn, 0, 0, 0};
    net_build(&n);
    memset(Jval, 0, sizeof(double) * (size_t)colptr[S]);
    csc_ctx c = {colptr, rowval, Jval};
    jac_accumulate(&n, u, k, put_csc, &c);
    net_free(&n);
}

/* ---- sparse LU over a given symbolic factorisation (row CSR of L\U, permuted) ---- */
typedef struct {
    int64_t S;
    const int64_t *perm;      /* perm[a] = species at pivot position a */
    int64_t *iperm;
    const int64_t *rowptr, *colidx, *diagpos;
    double *val;              /* nnzLU */
    double *work;             /* S dense work row */
    double hg_inv;            /* 1/(h*gamma) */
} lu_t;

static void put_lu(void *c, int64_t i, int64_t l, double v)
{
    lu_t *x = (lu_t *)c;
    int64_t a = x->iperm[i], b = x->iperm[l];
    int64_t lo = x->rowptr[a], hi = x->rowptr[a + 1] - 1;
    while (lo <= hi) {
        int64_t mid = (lo + hi) >> 1;
        if (x->colidx[mid] == b) { x->val[mid] -= v; return; }   /* W = I/(h g) - J */
        if (x->colidx[mid] < b) lo = mid + 1; else hi = mid - 1;
    }
}

static int lu_factor(lu_t *x)
{
    const int64_t S = x->S;
    double *w = x->work;
    for (int64_t i = 0; i < S; ++i) {
        for (int64_t p = x->rowptr[i]; p < x->rowptr[i + 1]; ++p) w[x->colidx[p]] = x->val[p];
        for (int64_t p = x->rowptr[i]; p < x->diagpos[i]; ++p) {
            int64_t kk = x->colidx[p];
            double l = w[kk] / x->val[x->diagpos[kk]];
            w[kk] = l;
            for (int64_t q = x->diagpos[kk] + 1; q < x->rowptr[kk + 1]; ++q)
                w[x->colidx[q]] -= l * x->val[q];
        }
        for (int64_t p = x->rowptr[i]; p < x->rowptr[i + 1]; ++p) x->val[p] = w[x->colidx[p]];
        double d = x->val[x->diagpos[i]];
        if (!(fabs(d) > 0.0) || !isfinite(d)) return 1;
    }
    return 0;
}

/* solve W x = b (b, x in species order) */
static void lu_solve(const lu_t *x, const double *b, double *out, double *y)
{
    const int64_t S = x->S;
    for (int64_t i = 0; i < S; ++i) {
        double s = b[x->perm[i]];
        for (int64_t p = x->rowptr[i]; p < x->diagpos[i]; ++p) s -= x->val[p] * y[x->colidx[p]];
        y[i] = s;
    }
    for (int64_t i = S - 1; i >= 0; --i) {
        double s = y[i];
        for (int64_t p = x->diagpos[i] + 1; p < x->rowptr[i + 1]; ++p) s -= x->val[p] * y[x->colidx[p]];
        y[i] = s / x->val[x->diagpos[i]];
    }
    for (int64_t i = 0; i < S; ++i) out[x->perm[i]] = y[i];
}

/* ---- Rodas4 tableau (Hairer & Wanner, RODAS), transformed K-form ---- */
static const double RG = 0.25;
static const double RA[6][6] = {
    {0},
    {0.1544000000000000e+01},
    {0.9466785280815826e+00, 0.2557011698983284e+00},
    {0.3314825187068521e+01, 0.2896124015972201e+01, 0.9986419139977817e+00},
    {0.1221224509226641e+01, 0.6019134481288629e+01, 0.1253708332932087e+02, -0.6878860361058950e+00},
    {0.1221224509226641e+01, 0.6019134481288629e+01, 0.1253708332932087e+02, -0.6878860361058950e+00, 1.0}};
static const double RC[6][6] = {
    {0},
    {-0.5668800000000000e+01},
    {-0.2430093356833875e+01, -0.2063599157091915e+00},
    {-0.1073529058151375e+00, -0.9594562251023355e+01, -0.2047028614809616e+02},
    {0.7496443313967647e+01, -0.1024680431464352e+02, -0.3399990352819905e+02, 0.1170890893206160e+02},
    {0.8083246795921522e+01, -0.7981132988064893e+01, -0.3152159432874371e+02, 0.1631930543123136e+02,
     -0.6058818238834054e+01}};

static double wrms(int64_t S, const double *e, const double *u0, const double *u1, double atol, double rtol)
{
    double s = 0.0;
    for (int64_t i = 0; i < S; ++i) {
        double sc = atol + rtol * fmax(fabs(u0[i]), fabs(u1[i]));
        double q = e[i] / sc;
        s += q * q;
    }
    return sqrt(s / (double)S);
}

/* starting step size of Hairer/Nørsett/Wanner II.4 for an order-4 method; also used to restart the
 * step size after every discrete rate update (the RHS jumps there). f, un, tmp: scratch. */
static double hinit(const net_t *n, const double *u, const double *k, double *f, double *un, double *tmp,
                    double abstol, double reltol)
{
    const int64_t S = n->S;
    rhs_eval(n, u, k, f);
    double d0 = 0, d1 = 0;
    for (int64_t i = 0; i < S; ++i) {
        double sc = abstol + reltol * fabs(u[i]);
        d0 += (u[i] / sc) * (u[i] / sc); d1 += (f[i] / sc) * (f[i] / sc);
    }
    d0 = sqrt(d0 / S); d1 = sqrt(d1 / S);
    double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
    for (int64_t i = 0; i < S; ++i) un[i] = u[i] + h0 * f[i];
    rhs_eval(n, un, k, tmp);
    double d2 = 0;
    for (int64_t i = 0; i < S; ++i) {
        double sc = abstol + reltol * fabs(u[i]);
        double q = (tmp[i] - f[i]) / sc; d2 += q * q;
    }
    d2 = sqrt(d2 / S) / h0;
    double dm = fmax(d1, d2);
    double h1 = (dm <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(0.01 / dm, 0.2);
    return fmin(100.0 * h0, h1);
}

static double g_facmin = 1.0 / 6.0, g_safe = 0.9, g_restart = 0.1;   /* restart = fraction of the fresh step estimate used after a rate update */
void ko_set_controller(double facmin, double safe, double restart) { g_facmin = facmin; g_safe = safe; g_restart = restart; }

enum { ST_OK = 0, ST_MAXITERS = 1, ST_DTMIN = 2, ST_SINGULAR = 3, ST_NAN = 4 };

/* Integrate every member.  Layouts: T_stop[b*nstops + s], u0[b*S + i] (u0_stride = 0 broadcasts
 * one vector), out_u[(b*Ns + s)*S + i].  stop_flags bit0 = rate update, bit1 = save.
 * stats[b*4] = {accepted, rejected, lu_factorisations, rhs_evals}.  Returns #failed members. */
int64_t ko_solve_rodas4(int64_t S, int64_t R, const int64_t *rp, const int64_t *ri, const int64_t *rn,
                        const int64_t *pp, const int64_t *pi, const int64_t *pn, const int64_t *perm,
                        const int64_t *rowptr, const int64_t *colidx, const int64_t *diagpos,
                        const double *A, const double *Ea, double k_max, double t_mult, int64_t B,
                        const double *T_init, int64_t nstops, const double *stop_t,
                        const int32_t *stop_flags, const double *T_stop, const double *u0,
                        int64_t u0_stride, double t0, double abstol, double reltol, double dtmin,
                        int64_t maxiters, int32_t ban_negatives, int64_t Ns, double *out_u,
                        int32_t *status, int64_t *stats, int32_t nthreads)
{
    net_t n = {S, R, rp, ri, rn, pp, pi, pn, 0, 0, 0};
    net_build(&n);
    int64_t nfail = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : nfail)
    for (int64_t b = 0; b < B; ++b) {
        const int64_t nnz = rowptr[S];
        lu_t lu;
        lu.S = S; lu.perm = perm; lu.rowptr = rowptr; lu.colidx = colidx; lu.diagpos = diagpos;
        lu.iperm = (int64_t *)malloc(sizeof(int64_t) * (size_t)S);
        for (int64_t a = 0; a < S; ++a) lu.iperm[perm[a]] = a;
        lu.val = (double *)malloc(sizeof(double) * (size_t)nnz);
        lu.work = (double *)calloc((size_t)S, sizeof(double));
        double *u = (double *)malloc(sizeof(double) * (size_t)S * 12);
        double *un = u + S, *f = un + S, *rhs = f + S, *y = rhs + S, *K[6];
        for (int s = 0; s < 6; ++s) K[s] = y + S + s * S;
        double *tmp = K[5] + S;
        double *k = (double *)malloc(sizeof(double) * (size_t)R);
        memcpy(u, u0 + b * u0_stride, sizeof(double) * (size_t)S);
        ko_arrhenius(R, A, Ea, T_init[b], k_max, t_mult, k);
        double t = t0, h;
        int64_t si = 0, isave = 0, nacc = 0, nrej = 0, nlu = 0, nrhs = 0;
        int st = ST_OK;
        /* stops at t0 fire before the first step (PresetTimeCallback at t0; harmless) */
        while (si < nstops && stop_t[si] <= t0) {
            if (stop_flags[si] & 1) ko_arrhenius(R, A, Ea, T_stop[b * nstops + si], k_max, t_mult, k);
            if (stop_flags[si] & 2) { memcpy(out_u + (b * Ns + isave) * S, u, sizeof(double) * (size_t)S); ++isave; }
            ++si;
        }
        /* initial step (Hairer/Nørsett/Wanner II.4 starting step, order 4) */
        h = hinit(&n, u, k, f, un, tmp, abstol, reltol); nrhs += 2;
        double err_old = 1.0, h_old = h;
        int rejected_last = 0, first_acc = 1;
        int64_t iters = 0;
        while (si < nstops) {
            if (++iters > maxiters) { st = ST_MAXITERS; break; }
            double tstop = stop_t[si];
            double hs = h;
            int hit = 0;
            if (t + 1.01 * hs >= tstop) { hs = tstop - t; hit = 1; }
            if (hs < dtmin && !hit) { st = ST_DTMIN; break; }
            /* W = I/(hs*gamma) - J(u) */
            lu.hg_inv = 1.0 / (hs * RG);
            memset(lu.val, 0, sizeof(double) * (size_t)nnz);
            jac_accumulate(&n, u, k, put_lu, &lu);
            for (int64_t a = 0; a < S; ++a) lu.val[diagpos[a]] += lu.hg_inv;
            ++nlu;
            int bad = lu_factor(&lu);
            double err = INFINITY;
            if (!bad) {
                for (int s = 0; s < 6; ++s) {
                    const double *Us = u;
                    if (s > 0) {
                        for (int64_t i = 0; i < S; ++i) {
                            double a = u[i];
                            for (int q = 0; q < s; ++q) a += RA[s][q] * K[q][i];
                            un[i] = a;
                        }
                        Us = un;
                    }
                    rhs_eval(&n, Us, k, f); ++nrhs;
                    for (int64_t i = 0; i < S; ++i) {
                        double a = f[i];
                        for (int q = 0; q < s; ++q) a += (RC[s][q] / hs) * K[q][i];
                        rhs[i] = a;
                    }
                    lu_solve(&lu, rhs, K[s], y);
                }
                /* un currently = u + sum a6j Kj (stage-6 argument); new solution = un + K6 */
                for (int64_t i = 0; i < S; ++i) tmp[i] = un[i] + K[5][i];
                err = wrms(S, K[5], u, tmp, abstol, reltol);
                if (!isfinite(err)) err = INFINITY;
                if (ban_negatives)
                    for (int64_t i = 0; i < S; ++i) if (tmp[i] < 0.0) { err = fmax(err, 1e4); break; }
            }
            /* step-size controller (RODAS: fac in [1/6, 5], safety 0.9, Gustafsson predictive) */
            double fac = isfinite(err) ? fmax(g_facmin, fmin(5.0, pow(err, 0.25) / g_safe)) : 5.0;
            double hnew = hs / fac;
            if (getenv("KO_DEBUG")) {
                int64_t im = 0; double qm = 0;
                for (int64_t i = 0; i < S; ++i) { double q = fabs(K[5][i]) / (abstol + reltol * fmax(fabs(u[i]), fabs(tmp[i]))); if (q > qm) { qm = q; im = i; } }
                fprintf(stderr, "it=%ld t=%.17g hs=%.6g err=%.6g hit=%d imax=%ld u=%.6g unew=%.6g K6=%.6g\n", (long)iters, t, hs, err, hit, (long)im, u[im], tmp[im], K[5][im]);
            }
            if (err <= 1.0) {
                ++nacc;
                if (!first_acc) {
                    double facgus = (h_old / hs) * pow(err * err / err_old, 0.25) / g_safe;
                    facgus = fmax(g_facmin, fmin(5.0, facgus));
                    fac = fmax(fac, facgus);
                    hnew = hs / fac;
                }
                first_acc = 0;
                h_old = hs; err_old = fmax(1e-2, err);
                if (rejected_last) hnew = fmin(hnew, hs);
                rejected_last = 0;
                memcpy(u, tmp, sizeof(double) * (size_t)S);
                if (hit) {
                    t = tstop;
                    h = fmax(hnew, h);            /* keep the pre-truncation proposal */
                    int updated = 0;
                    while (si < nstops && stop_t[si] <= t) {
                        if (stop_flags[si] & 1) { ko_arrhenius(R, A, Ea, T_stop[b * nstops + si], k_max, t_mult, k); updated = 1; }
                        if (stop_flags[si] & 2) { memcpy(out_u + (b * Ns + isave) * S, u, sizeof(double) * (size_t)S); ++isave; }
                        ++si;
                    }
                    if (updated && si < nstops) {   /* the RHS jumped: restart the step size */
                        h = fmin(h, g_restart * hinit(&n, u, k, f, un, tmp, abstol, reltol)); nrhs += 2;
                    }
                } else {
                    t += hs;
                    h = hnew;
                }
            } else {
                ++nrej;
                rejected_last = 1;
                h = hnew;
                if (h < dtmin) { st = ST_DTMIN; break; }
            }
        }
        status[b] = st;
        stats[b * 4 + 0] = nacc; stats[b * 4 + 1] = nrej; stats[b * 4 + 2] = nlu; stats[b * 4 + 3] = nrhs;
        if (st != ST_OK) ++nfail;
        free(lu.iperm); free(lu.val); free(lu.work); free(u); free(k);
    }
    net_free(&n);
    return nfail;
}
