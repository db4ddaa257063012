/* kinetica_b200.h — C ABI of libkinetica_b200.so
 *
 * B200-native (sm_100a) replacement for the kinetic-solve hot path of Kinetica.jl
 * v0.7.2.  The reference has no FFI for this path: everything below
 * `solve_network(method, sd, rd)` is Julia calling Julia (Catalyst / ModelingToolkit /
 * OrdinaryDiffEq).  Each entry point therefore names the reference code it replaces
 * (file:line under the reference checkout); the Julia-side `ccall` bindings are in
 * julia/KineticaB200.jl and INTEGRATION.md.
 *
 * Conventions
 *   - all functions return int32 status, 0 = OK; kb2_last_error(h) describes the last failure
 *   - indices are int64 and 0-based (the Julia shim subtracts 1)
 *   - reals are FP64
 *   - host arrays belong to the caller and are only read/written during the call
 *   - ensemble arrays are species-major x member-minor: x[i*B + b]
 *   - one handle = one GPU = one stream; a handle is not re-entrant
 */
#ifndef KINETICA_B200_H
#define KINETICA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kb2_ctx *kb2_handle;

/* per-member status words written by kb2_solve (reference: sol.retcode, and the
 * ErrorException("ODE solution failed.") of src/solving/solve_utils.jl:405-411) */
enum {
    KB2_OK = 0,
    KB2_MAXITERS = 1,      /* reference maxiters, src/solving/methods.jl:165 */
    KB2_DTMIN = 2,         /* reference dtmin = eps(tspan[end]), src/solving/methods.jl:164 */
    KB2_UNSTABLE = 3,      /* step size below dtmin with a non-finite error norm (singular pivot, overflow) */
    KB2_UNFINISHED = 5
};

/* condition-profile kinds understood by the device (reference src/conditions/*.jl) */
enum {
    KB2_PROFILE_STATIC = 0,          /* static.jl:7-9              params: [value] */
    KB2_PROFILE_NULL = 1,            /* direct_variable.jl:49-92 / gradient_variable.jl:70-114   [X_start] */
    KB2_PROFILE_LINEAR_DIRECT = 2,   /* direct_variable.jl:98-155   [rate, X_start, X_end, t_end] */
    KB2_PROFILE_LINEAR_GRADIENT = 3, /* gradient_variable.jl:120-175 [rate, X_start, X_end, t_end] */
    KB2_PROFILE_DOUBLE_RAMP = 4      /* gradient_variable.jl:181-310 [X_start, rate1, rate2, t_startr1,
                                        t_endr1, t_startr2, t_endr2, t_blend] */
};
#define KB2_PROFILE_NPARAMS 16

/* stop flags */
#define KB2_STOP_RATE 1   /* discrete rate update: CompleteRateUpdateAffect, solve_utils.jl:445-450 */
#define KB2_STOP_SAVE 2   /* saveat point, methods.jl:166 */
#define KB2_STOP_CHUNK 4  /* start of a new chunk of a chunkwise solve (methods.jl:185-303, 717-865): the integrator is
                             re-initialised there — maxiters counts per chunk, the step size restarts, tolerances
                             tightened by a retry are reset */

/* ---- lifetime ---- */
int32_t kb2_create(int32_t device, kb2_handle *out);
int32_t kb2_destroy(kb2_handle h);
const char *kb2_last_error(kb2_handle h);
/* number of kernels this handle has launched so far (bench.py's gpu_launches) */
int64_t kb2_launch_count(kb2_handle h);

/* ---- network: replaces make_rs (solve_utils.jl:318-334) + Catalyst/MTK codegen of f and J.
 * Input schema = RxData id/stoichiometry arrays flattened to CSR (network.jl:193-203). ---- */
int32_t kb2_set_network(kb2_handle h, int64_t S, int64_t R,
                        const int64_t *reac_ptr, const int64_t *reac_idx, const int64_t *reac_nu,
                        const int64_t *prod_ptr, const int64_t *prod_idx, const int64_t *prod_nu);

/* ---- symbolic analysis: replaces MTK `jac=true, sparse=true` pattern detection
 * (methods.jl:157-158) and KLU's symbolic phase.  ordering: 0 = minimum degree,
 * 1 = natural, 2 = caller-supplied via kb2_set_ordering, 3 = natural with hub species last,
 * 5 = reverse Cuthill-McKee and 6 / 7 = Sloan's profile reduction (weights 1:2 / 2:1) on the graph
 * without the hub species, hubs last, 4 = auto (3, unless one of 0, 5, 6, 7 has a modelled cost of the
 * factorisation and the triangular sweeps more than 3 % below it; see kb2_symbolic in kb2_api.cu). ---- */
int32_t kb2_set_ordering(kb2_handle h, const int64_t *perm);
int32_t kb2_symbolic(kb2_handle h, int32_t ordering, int64_t *nnzJ, int64_t *nnzLU, int64_t *n_fma);
int32_t kb2_get_pattern(kb2_handle h, int64_t *colptr, int64_t *rowval);          /* CSC of P_J */
int32_t kb2_get_ordering(kb2_handle h, int64_t *perm);
int32_t kb2_get_lu_pattern(kb2_handle h, int64_t *rowptr, int64_t *colidx, int64_t *diagpos);
/* block plan of the numeric factorisation (replaces KLU's numeric phase bookkeeping):
 * out[8] = {padded storage slots, panels, units (panel x column chunk), source-block tasks,
 * FMAs incl. padding, widest panel, target-map entries, the ordering in use (what `auto` chose)} */
int32_t kb2_get_plan_stats(kb2_handle h, int64_t *out);
/* raw plan tables for host-side verification; which: 0 p_row0, 1 p_nrows, 2 p_width, 3 p_next,
 * 4 p_base, 5 p_cptr, 6 cols, 7 u_info (12 per unit), 8 t_info (12 per task), 9 map, 10 slot_of,
 * 11 jslot, 12 diag_slot; gather tables of the right-hand side and the Jacobian: 13 rhs_ptr,
 * 14 rhs_rxn, 15 rhs_coef, 16 rhs_order, 17 rate_pos, 18 ell_ptr, 19 ell, 20 jt_ptr, 21 jt_rxn,
 * 22 jt_pack, 23 j_order, 24 drate_pos, 25 jell_ptr, 26 jell, 27 jt_pk,
 * 28 {rhs_nlong, j_nlong, jslots, ELL group size}; front plan of the window LU: 29 f_info (12 per
 * front), 30 lists, 31 init, 32 {fronts, window rows, window columns, max L rows, max U columns,
 * max init entries}.
 * Returns the length (copies when cap is large enough), -1 on error. */
int64_t kb2_get_plan_array(kb2_handle h, int32_t which, int32_t *out, int64_t cap);

/* ---- calculators: PrecalculatedArrheniusCalculator (calculator.jl:164-238);
 * k = A*T^n*exp(-Ea/(R*T))*N_A*t_mult, harmonic cap with k_max unless k_max is NaN;
 * n may be NULL (= the reference formula, which has no T^n term). ---- */
int32_t kb2_set_arrhenius(kb2_handle h, const double *A, const double *Ea, const double *n,
                          double k_max, double t_mult);
/* any other calculator: host-precomputed table k[s*R + r] for the rate-update stops, in stop
 * order (calculate_discrete_rates, solve_utils.jl:91-109); k_init[R] = calc(initial conditions)
 * (methods.jl:668).  Shared by all members. */
int32_t kb2_set_rate_table(kb2_handle h, int64_t n_rate_stops, const double *k_table,
                           const double *k_init);

/* ---- conditions: per-member profile of the :T condition (ConditionSet, condition_set.jl:1-58) ---- */
int32_t kb2_set_profiles(kb2_handle h, int64_t B, const int32_t *kind, const double *params);
/* optional override: condition value per member at every stop, T[b*nstops + s] (NaN = use the
 * profile) — the reference reads the interpolated profile solution, solve_utils.jl:101-104 */
int32_t kb2_set_T_table(kb2_handle h, int64_t B, int64_t nstops, const double *T);
/* merged stop list shared by all members: sorted, last entry = tspan[2]
 * (get_tstops condition_set.jl:172-176, tstops kwarg methods.jl:697, saveat :166) */
int32_t kb2_set_stops(kb2_handle h, int64_t nstops, const double *stop_t, const int32_t *flags);
/* per-member stop lists (members whose ConditionSets have different tstops):
 * stop_t[b*nstops_max + s], flags likewise, counts[b] valid entries per member; every member must
 * carry the same save points */
int32_t kb2_set_member_stops(kb2_handle h, int64_t B, int64_t nstops_max, const int32_t *counts,
                             const double *stop_t, const int32_t *flags);

/* ---- the solve: replaces init/solve!/adaptive_solve! (methods.jl:174-180, 705-711,
 * solve_utils.jl:376-424) for B members at once.
 *   u0[u0_stride*b + i]  (u0_stride = 0 broadcasts one vector; else u0_stride = S)
 *   out_u[(s*S + i)*B + b]  s = save index;  out_umax[i*B + b] = max over saves (NULL to skip)
 *   status[b];  stats[b*8] = {accepted, rejected, lu, rhs, saves, stops passed, attempts in the last chunk, chunk retries} ---- */
int32_t kb2_solve(kb2_handle h, int64_t B, const double *u0, int64_t u0_stride, double t0,
                  double abstol, double reltol, double dtmin, int64_t maxiters,
                  int32_t ban_negatives, int64_t Ns, double *out_u, double *out_umax,
                  int32_t *status, int64_t *stats);
/* Batch tiling: kb2_solve walks an ensemble that does not fit the device memory in chunks of
 * b_tile members (members are independent; network tables, plans and calculator stay resident).
 * kb2_memory_plan reports the device bytes one member needs for Ns save points, the chunk size the
 * free memory allows (or the kb2_set_batch_tile override; 0 = automatic) and the free bytes;
 * kb2_last_batch_tiles = chunks the last kb2_solve used. */
int32_t kb2_memory_plan(kb2_handle h, int64_t Ns, int64_t *bytes_per_member, int64_t *b_tile, int64_t *free_bytes);
int32_t kb2_set_batch_tile(kb2_handle h, int64_t b_tile);
int64_t kb2_last_batch_tiles(kb2_handle h);
/* continuous rate updates (methods.jl:363-458: k follows the condition profile inside the step instead
 * of being held between tstops): every Rodas4 stage evaluates k(T_b(t + c_s h)) from the member's
 * profile on the device, and the non-autonomous term h d_s df/dt (df/dt = f(u; dk/dT dT/dt)) enters
 * the stages.  Needs kb2_set_arrhenius + kb2_set_profiles; stops flagged KB2_STOP_RATE are then only
 * forced step ends (the kinks of the profiles). */
int32_t kb2_set_continuous(kb2_handle h, int32_t continuous);
/* chunkwise solves: adaptive_solve! per chunk (solve_utils.jl:376-424 inside the chunk loops) on the
 * device — a member whose chunk fails (maxiters, dtmin) repeats it from the chunk's start state with
 * abstol / reltol x0.1, at most five attempts; update_tols keeps the tightened tolerances for the
 * following chunks (params.jl update_tols).  Without chunk stops the whole solve is one chunk. */
int32_t kb2_set_chunking(kb2_handle h, int32_t retry_failed_chunks, int32_t update_tols);
/* the same in three phases (bench.py times `run` alone with inputs resident in HBM); the whole
 * ensemble must fit the device memory */
int32_t kb2_solve_prepare(kb2_handle h, int64_t B, const double *u0, int64_t u0_stride, double t0,
                          double abstol, double reltol, double dtmin, int64_t maxiters,
                          int32_t ban_negatives, int64_t Ns);
/* `run` drives the phase kernels of the solve from a host loop: per attempted
 * step  W assembly + LU | 6 x (stage right-hand side | triangular sweeps) | error control + stop
 * handling, launched back to back on the handle's stream in batches of 16 rounds (one CUDA graph per
 * batch; KB2_GRAPH=0: plain launches); the loop reads a "members still running" word back per batch.  ms_device = CUDA-event time from the first to the last launch. */
int32_t kb2_solve_run(kb2_handle h, float *ms_device);
int32_t kb2_solve_fetch(kb2_handle h, double *out_u, double *out_umax, int32_t *status, int64_t *stats);
/* phase timing of the last kb2_solve_run, from CUDA events around every kernel of the sampled
 * rounds (one round in 64): ms_avg[5] = average launch duration of {LU, stage right-hand side,
 * stage sweeps, step end, Jacobian values}, launches_sampled[5], rounds = rounds the loop ran */
int32_t kb2_get_phase_times(kb2_handle h, double *ms_avg, int64_t *launches_sampled, int64_t *rounds);
/* device-side results for the multi-GPU allgather: packs final concentrations and per-species
 * maxima member-major into caller-provided DEVICE buffers final_bs[b*S+i], umax_bs[b*S+i] */
int32_t kb2_pack_results_device(kb2_handle h, double *final_bs_dev, double *umax_bs_dev);

/* ---- multi-GPU: members are independent, each GPU solves a contiguous slice
 * of the ensemble with its own handle and no communication; the one exchange is an all-gather of
 * the final concentrations and the per-species maxima (what identify_next_seeds consumes,
 * explore_utils.jl:344-349) over NCCL / NVLink.  NCCL is bound at run time (dlopen).
 *   process per GPU : rank 0 calls kb2_comm_unique_id, ships the 128 bytes to the other ranks by
 *                     any means, every rank calls kb2_comm_init_rank
 *   single process  : kb2_comm_init_all over the handles of all devices (ncclCommInitAll)
 * kb2_allgather_results packs the results of the last solve member-major on the device and gathers
 * them; n = local handles (1 per process, or all of them: the calls are grouped).  Host outputs
 * final_all[i][(r*B + b)*S + s] (r = rank) may be NULL to leave the data on the device. ---- */
int32_t kb2_comm_unique_id(uint8_t *id128);
int32_t kb2_comm_init_rank(kb2_handle h, int32_t nranks, int32_t rank, const uint8_t *id128);
int32_t kb2_comm_init_all(int32_t ndev, kb2_handle *handles);
int32_t kb2_allgather_results(kb2_handle *handles, int32_t n, double **final_all, double **umax_all);
int32_t kb2_gathered_device(kb2_handle h, double **final_all_dev, double **umax_all_dev, float *gather_ms,
                            int32_t *rank, int32_t *nranks);

/* ---- kernel-level entry points (parity tests + per-kernel roofline) ---- */
int32_t kb2_eval_k(kb2_handle h, int64_t B, const double *T, double *k_out);
int32_t kb2_eval_profile(kb2_handle h, int64_t B, int64_t nt, const double *t, double *X_out);
int32_t kb2_eval_rhs(kb2_handle h, int64_t B, const double *u, const double *k, double *du);
int32_t kb2_eval_jac(kb2_handle h, int64_t B, const double *u, const double *k, double *Jval);
int32_t kb2_factor(kb2_handle h, int64_t B, const double *u, const double *k,
                   const double *hg_inv, double *lu_out);
int32_t kb2_trisolve(kb2_handle h, int64_t B, const double *rhs, double *x);
/* time `iters` launches of one standalone kernel on resident data; which: 0 arrhenius, 1 rhs,
 * 2 jacobian values, 3 factorisation as the solve runs it (Jacobian values + window LU, or the
 * block-plan assembly + LU when the window does not fit), 4 trisolve, 5 block-plan W assembly,
 * 6 block-plan LU, 7 window LU alone, 8 block-plan assembly + LU */
int32_t kb2_time_kernel(kb2_handle h, int32_t which, int64_t B, int32_t iters, float *ms_avg);

/* measured FP64 FMA peak of the handle's device in TFLOP/s (dependency-free DFMA chains on every
 * SM, CUDA-event timed): the denominator of the factorisation's FP64 fraction */
int32_t kb2_measure_fp64_peak(kb2_handle h, double *tflops);

/* tuning: members per warp tile (1, 2 or 4; 0 = auto); the second argument is reserved (pass 0) */
int32_t kb2_set_tiling(kb2_handle h, int32_t members_per_tile, int32_t reserved);
/* members per warp tile and resident solve warps per SM of the last allocation / solve launch */
int32_t kb2_get_launch_info(kb2_handle h, int32_t *members_per_tile, int32_t *ctas_per_sm);

#ifdef __cplusplus
}
#endif
#endif
