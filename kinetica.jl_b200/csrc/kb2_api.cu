// C ABI of libkinetica_b200.so (see include/kinetica_b200.h).  Host plumbing only: owns device
// memory behind the opaque handle, copies caller arrays in/out, launches the kernels of
// kb2_kernels.cuh.  There is no CPU fallback: without a usable CUDA device every compute entry
// point fails with a non-zero status.
#include "kb2_solve.cuh"
#include "kb2_front.cuh"
#include "kb2_internal.h"
#include "../../include/kinetica_b200.h"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>

using namespace kb2;

struct kb2_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int sm_count = 148;
    size_t smem_optin = 0;
    std::string err;
    int64_t launches = 0;
    Network net;
    Symbolic sym;
    bool net_on_device = false;
    // calculator (host copies)
    int calc_mode = -1;
    std::vector<double> A, Ea, nexp, ktab, kinit;
    bool has_n = false;
    double k_max = NAN, t_mult = 1.0;
    int64_t n_rate_stops = 0;
    // conditions (host copies)
    std::vector<int32_t> pkind;
    std::vector<double> pparams, Ttab, stop_t;
    std::vector<int32_t> stop_flags, stop_cnt;     // stop_cnt empty = one list shared by all members
    int64_t Bprof = 0, Btab = 0, ns_row = 0, Bstops = 0;
    // device
    std::vector<void *> net_allocs, calc_allocs, ens_allocs;
    DevNet dn{};
    DevPlan dp{};
    DevFront df{};
    bool window_ok = false;       // the front plan's window fits the shared memory of an SM for the current tile size
    int chunk_retry = 0, chunk_update_tols = 0;     // kb2_set_chunking
    int continuous = 0;                             // kb2_set_continuous
    bool has_chunk_stops = false;
    int64_t b_tile_user = 0;      // batch tile override (0: sized from the free device memory)
    int64_t last_tiles = 0;       // batch tiles of the last kb2_solve
    int window_mw = 0;            // members per CTA of the window LU
    int last_window_ctas = 0;     // resident window CTAs per SM
    int window_stagger_ns = 0;    // start delay of the second CTA of an SM
    DevEns de{};
    int64_t ens_B = -1, ens_Ns = -1;
    size_t ens_fixed = 0;
    int mb_user = 0, last_ctas_per_sm = 0;
    int ens_mb = 0;               // members per warp tile of the current ensemble allocation
    int auto_ordering = -1;       // the ordering kb2_symbolic ended up with (its choice under `auto`)
    double *stage = nullptr;      // device staging for layout conversion (caller rows <-> tiles)
    size_t stage_cap = 0;
    double *scal = nullptr;       // [Bp] per-member scalars of the kernel-level entry points
    bool prepared = false;
    // phase timing of the last solve (sampled rounds), and its host loop
    std::vector<cudaEvent_t> phase_ev;
    double phase_ms[5] = {0, 0, 0, 0, 0};
    long long phase_n[5] = {0, 0, 0, 0, 0};
    long long rounds = 0;
    int *h_flag = nullptr;        // pinned
    // multi-GPU: one NCCL communicator per handle (kb2_comm_*), gather buffers on the device
    ncclComm_t comm = nullptr;
    int comm_nranks = 1, comm_rank = 0;
    double *g_send = nullptr, *g_recv = nullptr;     // [2][B][S] packed results, [2][nranks][B][S] gathered
    size_t g_send_cap = 0, g_recv_cap = 0;
    float gather_ms = 0.f;
};

#define FAIL(h, msg) do { (h)->err = (msg); return 1; } while (0)
#define CU(h, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    (h)->err = std::string(#call) + ": " + cudaGetErrorString(e_); return 2; } } while (0)

template <class T>
static int dev_upload(kb2_ctx *h, std::vector<void *> &pool, const T *src, size_t n, const T **out)
{
    void *p = nullptr;
    CU(h, cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    pool.push_back(p);
    if (n) CU(h, cudaMemcpyAsync(p, src, n * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    *out = (const T *)p;
    return 0;
}

template <class T>
static int dev_alloc(kb2_ctx *h, std::vector<void *> &pool, size_t n, T **out, bool zero = true)
{
    void *p = nullptr;
    CU(h, cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    pool.push_back(p);
    if (zero) CU(h, cudaMemsetAsync(p, 0, std::max<size_t>(n, 1) * sizeof(T), h->stream));
    *out = (T *)p;
    return 0;
}

static void free_pool(std::vector<void *> &pool)
{
    for (void *p : pool) cudaFree(p);
    pool.clear();
}

static void kb2_comm_release(kb2_ctx *h);

extern "C" int32_t kb2_create(int32_t device, kb2_handle *out)
{
    if (!out) return 1;
    *out = nullptr;
    if (device < 0) {       // host-only handle: symbolic analysis works, every compute call fails
        kb2_ctx *h = new kb2_ctx();
        h->device = -1;
        *out = h;
        return 0;
    }
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device >= n) return 3;   // no GPU: no fallback
    kb2_ctx *h = new kb2_ctx();
    h->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete h; return 3; }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    h->sm_count = prop.multiProcessorCount;
    h->smem_optin = prop.sharedMemPerBlockOptin;
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return 3; }
    cudaEventCreate(&h->ev0);
    cudaEventCreate(&h->ev1);
    if (cudaMallocHost((void **)&h->h_flag, KB2_FLAG_SLOTS * sizeof(int)) != cudaSuccess) {
        cudaStreamDestroy(h->stream);
        delete h;
        return 3;
    }
    *out = h;
    return 0;
}

extern "C" int32_t kb2_destroy(kb2_handle h)
{
    if (!h) return 0;
    if (h->device < 0) { delete h; return 0; }
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    free_pool(h->net_allocs);
    free_pool(h->calc_allocs);
    free_pool(h->ens_allocs);
    cudaFree(h->stage);
    cudaFree(h->g_send);
    cudaFree(h->g_recv);
    kb2_comm_release(h);
    cudaFreeHost(h->h_flag);
    for (auto &ev : h->phase_ev) cudaEventDestroy(ev);
    cudaEventDestroy(h->ev0);
    cudaEventDestroy(h->ev1);
    cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

extern "C" const char *kb2_last_error(kb2_handle h) { return h ? h->err.c_str() : "null handle"; }
extern "C" int64_t kb2_launch_count(kb2_handle h) { return h ? h->launches : 0; }

extern "C" int32_t kb2_set_tiling(kb2_handle h, int32_t mb, int32_t reserved)
{
    if (!h) return 1;
    (void)reserved;
    if (mb != 0 && mb != 1 && mb != 2 && mb != 4) FAIL(h, "members_per_tile must be 0 (auto), 1, 2 or 4");
    h->mb_user = mb;
    h->ens_B = -1;            // the tile size is baked into the device layout
    h->prepared = false;
    return 0;
}

extern "C" int32_t kb2_set_network(kb2_handle h, int64_t S, int64_t R, const int64_t *reac_ptr,
                                   const int64_t *reac_idx, const int64_t *reac_nu, const int64_t *prod_ptr,
                                   const int64_t *prod_idx, const int64_t *prod_nu)
{
    if (!h) return 1;
    if (S <= 0 || R < 0 || S >= (1 << 30) || R >= (1 << 30)) FAIL(h, "bad network size");
    Network n;
    n.S = S; n.R = R;
    n.rp.assign(reac_ptr, reac_ptr + R + 1);
    n.pp.assign(prod_ptr, prod_ptr + R + 1);
    if (n.rp[0] != 0 || n.pp[0] != 0) FAIL(h, "CSR pointers must start at 0");
    for (int64_t j = 0; j < R; ++j) if (n.rp[j + 1] < n.rp[j] || n.pp[j + 1] < n.pp[j]) FAIL(h, "CSR pointers must be non-decreasing");
    n.ri.assign(reac_idx, reac_idx + n.rp[R]); n.rn.assign(reac_nu, reac_nu + n.rp[R]);
    n.pi.assign(prod_idx, prod_idx + n.pp[R]); n.pn.assign(prod_nu, prod_nu + n.pp[R]);
    std::string e = build_network(n);
    if (!e.empty()) FAIL(h, e);
    h->net = std::move(n);
    h->sym = Symbolic();
    h->net_on_device = false;
    h->calc_mode = -1;
    h->prepared = false;
    return 0;
}

extern "C" int32_t kb2_set_ordering(kb2_handle h, const int64_t *perm)
{
    if (!h) return 1;
    if (h->net.S <= 0) FAIL(h, "set the network first");
    h->sym.perm.assign(perm, perm + h->net.S);
    return 0;
}

static int upload_network(kb2_ctx *h)
{
    if (h->device < 0) return 0;     // host-only handle
    CU(h, cudaSetDevice(h->device));
    free_pool(h->net_allocs);
    const Symbolic &s = h->sym;
    const PanelPlan &pp = s.panels;
    DevNet &d = h->dn;
    d = DevNet{};
    d.S = (int)h->net.S; d.R = (int)h->net.R; d.nnzJ = (int)s.nnzJ;
    int rc = 0;
    auto &P = h->net_allocs;
    const int32_t *rd = nullptr;
    std::vector<int32_t> perm32(s.perm.begin(), s.perm.end());
    rc |= dev_upload(h, P, s.rhs_ptr.data(), s.rhs_ptr.size(), &d.rhs_ptr);
    rc |= dev_upload(h, P, s.rhs_src.data(), s.rhs_src.size(), &d.rhs_rxn);      // positions in the first-touch rate table
    rc |= dev_upload(h, P, s.rate_pos.data(), s.rate_pos.size(), &d.rate_pos);
    rc |= dev_upload(h, P, s.drate_pos.data(), s.drate_pos.size(), &d.drate_pos);
    rc |= dev_upload(h, P, s.rhs_coef.data(), s.rhs_coef.size(), &d.rhs_coef);
    rc |= dev_upload(h, P, s.rdesc.data(), s.rdesc.size(), &rd);
    d.rdesc = (const int4 *)rd;
    rc |= dev_upload(h, P, s.jt_ptr.data(), s.jt_ptr.size(), &d.jt_ptr);
    rc |= dev_upload(h, P, s.jt_pk.data(), s.jt_pk.size(), &d.jt_pk);
    rc |= dev_upload(h, P, s.jell_ptr.data(), s.jell_ptr.size(), &d.jell_ptr);
    rc |= dev_upload(h, P, s.jell.data(), s.jell.size(), &d.jell);
    d.jell_ngroups = (int)s.jell_ptr.size() - 1;
    d.jslots = s.jslots;
    rc |= dev_upload(h, P, s.rhs_order.data(), s.rhs_order.size(), &d.rhs_order);
    rc |= dev_upload(h, P, s.j_order.data(), s.j_order.size(), &d.j_order);
    d.rhs_nlong = s.rhs_nlong; d.j_nlong = s.j_nlong;
    rc |= dev_upload(h, P, s.ell_ptr.data(), s.ell_ptr.size(), &d.ell_ptr);
    rc |= dev_upload(h, P, s.ell.data(), s.ell.size(), &d.ell);
    d.ell_ngroups = (int)s.ell_ptr.size() - 1;
    rc |= dev_upload(h, P, pp.jslot.data(), pp.jslot.size(), &d.jslot);
    rc |= dev_upload(h, P, pp.diag_slot.data(), pp.diag_slot.size(), &d.diag_slot);
    rc |= dev_upload(h, P, perm32.data(), perm32.size(), &d.perm);
    // block plan: the LU value storage is the padded panel layout
    {
        DevPlan &q = h->dp;
        q = DevPlan{};
        q.npanels = (int)pp.p_row0.size(); q.nunits = (int)pp.n_units(); q.padded = (int)pp.padded;
        const int32_t *ui = nullptr, *ti = nullptr;
        rc |= dev_upload(h, P, pp.p_row0.data(), pp.p_row0.size(), &q.p_row0);
        rc |= dev_upload(h, P, pp.p_nrows.data(), pp.p_nrows.size(), &q.p_nrows);
        rc |= dev_upload(h, P, pp.p_width.data(), pp.p_width.size(), &q.p_width);
        rc |= dev_upload(h, P, pp.p_next.data(), pp.p_next.size(), &q.p_next);
        rc |= dev_upload(h, P, pp.p_base.data(), pp.p_base.size(), &q.p_base);
        rc |= dev_upload(h, P, pp.p_cptr.data(), pp.p_cptr.size(), &q.p_cptr);
        rc |= dev_upload(h, P, pp.cols.data(), pp.cols.size(), &q.cols);
        rc |= dev_upload(h, P, pp.u_info.data(), pp.u_info.size(), &ui);
        rc |= dev_upload(h, P, pp.t_info.data(), pp.t_info.size(), &ti);
        rc |= dev_upload(h, P, pp.map.data(), pp.map.size(), &q.map);
        q.u_info = (const int4 *)ui; q.t_info = (const int4 *)ti;
    }
    {
        const FrontPlan &f = s.fronts;
        DevFront &q = h->df;
        q = DevFront{};
        q.NF = f.NF; q.Wr = f.Wr; q.Wc = f.Wc; q.max_nl = f.max_nl; q.max_nu = f.max_nu;
        const int32_t *ini = nullptr;
        rc |= dev_upload(h, P, f.f_info.data(), f.f_info.size(), &q.f_info);
        rc |= dev_upload(h, P, f.lists.data(), f.lists.size(), &q.lists);
        rc |= dev_upload(h, P, f.init.data(), f.init.size(), &ini);
        q.init = (const int2 *)ini;
    }
    if (rc) return rc;
    CU(h, cudaStreamSynchronize(h->stream));
    free_pool(h->calc_allocs);
    h->net_on_device = true;
    h->calc_mode = -1;          // a new network needs its calculator set again
    return 0;
}

extern "C" int32_t kb2_symbolic(kb2_handle h, int32_t ordering, int64_t *nnzJ, int64_t *nnzLU, int64_t *n_fma)
{
    if (!h) return 1;
    if (h->net.S <= 0) FAIL(h, "set the network first");
    std::string e;
    if (ordering == 4) {
        // auto: the candidate ordering with the smallest modelled cost of the numeric phases that
        // depend on it, per tile of four members in SM cycles (calibrated on C3):
        // the window LU's update (8 x 4 register blocks), its strips (one task per L row / U column),
        // the per-front overhead (more for a front whose pivot block cannot be factorised ahead), and
        // the six triangular sweeps of a step, which stream the padded panel storage.  A window that
        // does not fit shared memory with four members per CTA costs the factorisation a factor
        // (fewer members per CTA, or the left-looking block plan).
        const size_t cap = h->smem_optin ? h->smem_optin : (size_t)227 * 1024;
        auto cost = [&](const Symbolic &s) {
            const FrontPlan &f = s.fronts;
            double blocks = 0, tasks = 0, serial = 0;
            for (int32_t P = 0; P < f.NF; ++P) {
                const int32_t *r = f.f_info.data() + (size_t)P * FrontPlan::FREC;
                blocks += (double)((r[3] + 7) / 8) * ((r[2] + WL_CB - 1) / WL_CB);
                tasks += r[2] + r[3];
                serial += r[10] ? 0 : 1;       // pivot block factorised in a phase of its own (no look-ahead)
            }
            double lu = 66.3 * blocks + 38.9 * tasks + 4523.0 * f.NF + 2000.0 * serial;
            int mw = 0;
            for (int c = 4; c >= 1 && !mw; c >>= 1) if (wl_smem_bytes(c, f.Wr, f.Wc, f.max_nl, f.max_nu) <= cap) mw = c;
            lu *= mw ? 4.0 / mw : 6.0;
            return lu + 5.25 * (double)s.panels.padded;       // slope of the sweeps in the same-box A/B: 0.772 -> 0.732 ms for 120 620 -> 107 787 slots
        };
        // The natural order with hub species last is the baseline (every full-size run of configs 3-5
        // was first made with it); another candidate replaces it only if the model puts it more than
        // 3 % ahead (the model's error: for the LU of C3 it says -9 % where the same-box A/B measured -7 %).
        Symbolic best, base;
        double best_cost = 0.0, base_cost = 0.0;
        int best_cand = -1;
        std::string first_error;
        // minimum degree goes first: it is the robust one, and its fill bounds what the banded
        // candidates may spend — on a network without locality their fill explodes: a candidate whose
        // envelope (an upper bound of a banded order's fill) is beyond 2 x the minimum-degree nnzLU is
        // dropped before its symbolic LU, and a symbolic LU is abandoned at 16 x the minimum-degree FMAs
        int64_t fma_limit = INT64_MAX, env_limit = INT64_MAX;
        for (int cand : {0, 3, 5, 6, 7}) {
            Symbolic c;
            e = build_symbolic(h->net, cand, c, fma_limit, env_limit);
            if (e.empty() && cand == 0) { fma_limit = 16 * c.n_fma + 1000000; env_limit = 2 * c.nnzLU + 100000; }
            if (e.empty()) e = build_panels(c, h->net.S);
            if (e.empty()) e = build_fronts(c, h->net.S);
            if (!e.empty()) {       // a candidate may be out of reach (e.g. the fill of a banded order on a network without locality)
                if (first_error.empty()) first_error = e;
                continue;
            }
            const double cc = cost(c);
            if (cand == 3) { base = std::move(c); base_cost = cc; continue; }
            if (best_cand < 0 || cc < best_cost) { best = std::move(c); best_cost = cc; best_cand = cand; }
        }
        if (best_cand < 0 && !base.fronts.ready) FAIL(h, first_error);
        if (base.fronts.ready && (best_cand < 0 || best_cost >= 0.97 * base_cost)) { h->sym = std::move(base); h->auto_ordering = 3; }
        else { h->sym = std::move(best); h->auto_ordering = best_cand; }
    } else {
        e = build_symbolic(h->net, ordering, h->sym);
        if (!e.empty()) FAIL(h, e);
        e = build_panels(h->sym, h->net.S);
        if (!e.empty()) FAIL(h, e);
        e = build_fronts(h->sym, h->net.S);
        if (!e.empty()) FAIL(h, e);
        h->auto_ordering = ordering;
    }
    if (nnzJ) *nnzJ = h->sym.nnzJ;
    if (nnzLU) *nnzLU = h->sym.nnzLU;
    if (n_fma) *n_fma = h->sym.n_fma;
    h->prepared = false;
    h->ens_B = -1;
    return upload_network(h);
}

extern "C" int32_t kb2_get_plan_stats(kb2_handle h, int64_t *out)
{
    if (!h || !h->sym.ready) return 1;
    const PanelPlan &pp = h->sym.panels;
    out[0] = pp.padded; out[1] = (int64_t)pp.p_row0.size(); out[2] = pp.n_units();
    out[3] = pp.n_tasks(); out[4] = pp.n_fma_padded; out[5] = pp.max_width;
    out[6] = (int64_t)pp.map.size(); out[7] = h->auto_ordering;
    return 0;
}

extern "C" int64_t kb2_get_plan_array(kb2_handle h, int32_t which, int32_t *out, int64_t cap)
{
    if (!h || !h->sym.ready) return -1;
    const PanelPlan &pp = h->sym.panels;
    const std::vector<int32_t> *v = nullptr;
    switch (which) {
    case 0: v = &pp.p_row0; break;
    case 1: v = &pp.p_nrows; break;
    case 2: v = &pp.p_width; break;
    case 3: v = &pp.p_next; break;
    case 4: v = &pp.p_base; break;
    case 5: v = &pp.p_cptr; break;
    case 6: v = &pp.cols; break;
    case 7: v = &pp.u_info; break;
    case 8: v = &pp.t_info; break;
    case 9: v = &pp.map; break;
    case 10: v = &pp.slot_of; break;
    case 11: v = &pp.jslot; break;
    case 12: v = &pp.diag_slot; break;
    // gather tables of the right-hand side and the Jacobian (kb2_symbolic.cpp)
    case 13: v = &h->sym.rhs_ptr; break;
    case 14: v = &h->sym.rhs_rxn; break;
    case 15: v = &h->sym.rhs_coef; break;
    case 16: v = &h->sym.rhs_order; break;
    case 17: v = &h->sym.rate_pos; break;
    case 18: v = &h->sym.ell_ptr; break;
    case 19: v = &h->sym.ell; break;
    case 20: v = &h->sym.jt_ptr; break;
    case 21: v = &h->sym.jt_rxn; break;
    case 22: v = &h->sym.jt_pack; break;
    case 23: v = &h->sym.j_order; break;
    case 24: v = &h->sym.drate_pos; break;
    case 25: v = &h->sym.jell_ptr; break;
    case 26: v = &h->sym.jell; break;
    case 27: v = &h->sym.jt_pk; break;
    case 28: {
        const int32_t meta[4] = {h->sym.rhs_nlong, h->sym.j_nlong, h->sym.jslots, Symbolic::ELL_G};
        if (out && cap >= 4) std::copy(meta, meta + 4, out);
        return 4;
    }
    // front plan of the window LU (kb2_front.cpp)
    case 29: v = &h->sym.fronts.f_info; break;
    case 30: v = &h->sym.fronts.lists; break;
    case 31: v = &h->sym.fronts.init; break;
    case 32: {
        const FrontPlan &f = h->sym.fronts;
        const int32_t meta[6] = {f.NF, f.Wr, f.Wc, f.max_nl, f.max_nu, f.max_init};
        if (out && cap >= 6) std::copy(meta, meta + 6, out);
        return 6;
    }
    default: return -1;
    }
    if (out && cap >= (int64_t)v->size()) std::copy(v->begin(), v->end(), out);
    return (int64_t)v->size();
}

extern "C" int32_t kb2_get_pattern(kb2_handle h, int64_t *colptr, int64_t *rowval)
{
    if (!h || !h->sym.ready) return 1;
    std::copy(h->sym.colptr.begin(), h->sym.colptr.end(), colptr);
    std::copy(h->sym.rowval.begin(), h->sym.rowval.end(), rowval);
    return 0;
}

extern "C" int32_t kb2_get_ordering(kb2_handle h, int64_t *perm)
{
    if (!h || !h->sym.ready) return 1;
    std::copy(h->sym.perm.begin(), h->sym.perm.end(), perm);
    return 0;
}

extern "C" int32_t kb2_get_lu_pattern(kb2_handle h, int64_t *rowptr, int64_t *colidx, int64_t *diagpos)
{
    if (!h || !h->sym.ready) return 1;
    std::copy(h->sym.rowptr.begin(), h->sym.rowptr.end(), rowptr);
    std::copy(h->sym.colidx.begin(), h->sym.colidx.end(), colidx);
    std::copy(h->sym.diagpos.begin(), h->sym.diagpos.end(), diagpos);
    return 0;
}

extern "C" int32_t kb2_set_arrhenius(kb2_handle h, const double *A, const double *Ea, const double *n,
                                     double k_max, double t_mult)
{
    if (!h) return 1;
    if (!h->net_on_device) FAIL(h, "run kb2_symbolic before setting the calculator");
    const int64_t R = h->net.R;
    h->A.assign(A, A + R); h->Ea.assign(Ea, Ea + R);
    h->has_n = n != nullptr;
    if (n) h->nexp.assign(n, n + R);
    h->k_max = k_max; h->t_mult = t_mult;
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaStreamSynchronize(h->stream));
    free_pool(h->calc_allocs);          // the previous calculator's tables
    h->prepared = false;
    int rc = dev_upload(h, h->calc_allocs, h->A.data(), (size_t)R, &h->dn.A);
    rc |= dev_upload(h, h->calc_allocs, h->Ea.data(), (size_t)R, &h->dn.Ea);
    h->dn.n = nullptr;
    if (n) rc |= dev_upload(h, h->calc_allocs, h->nexp.data(), (size_t)R, &h->dn.n);
    if (rc) return rc;
    h->dn.k_max = k_max; h->dn.t_mult = t_mult; h->dn.calc_mode = 0;
    h->calc_mode = 0;
    return 0;
}

extern "C" int32_t kb2_set_rate_table(kb2_handle h, int64_t n_rate_stops, const double *k_table, const double *k_init)
{
    if (!h) return 1;
    if (!h->net_on_device) FAIL(h, "run kb2_symbolic before setting the calculator");
    const int64_t R = h->net.R;
    h->n_rate_stops = n_rate_stops;
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaStreamSynchronize(h->stream));
    free_pool(h->calc_allocs);          // the previous calculator's tables
    h->prepared = false;
    int rc = dev_upload(h, h->calc_allocs, k_table, (size_t)(n_rate_stops * R), &h->dn.ktab);
    rc |= dev_upload(h, h->calc_allocs, k_init, (size_t)R, &h->dn.kinit);
    if (rc) return rc;
    CU(h, cudaStreamSynchronize(h->stream));
    h->dn.calc_mode = 1;
    h->calc_mode = 1;
    return 0;
}

extern "C" int32_t kb2_set_profiles(kb2_handle h, int64_t B, const int32_t *kind, const double *params)
{
    if (!h) return 1;
    if (B <= 0) FAIL(h, "B must be positive");
    for (int64_t b = 0; b < B; ++b) if (kind[b] < 0 || kind[b] > 4) FAIL(h, "unknown profile kind");
    h->pkind.assign(kind, kind + B);
    h->pparams.assign(params, params + B * KB2_PROFILE_NPARAMS);
    h->Bprof = B;
    h->prepared = false;
    return 0;
}

extern "C" int32_t kb2_set_T_table(kb2_handle h, int64_t B, int64_t nstops, const double *T)
{
    if (!h) return 1;
    if (!T) { h->Ttab.clear(); h->Btab = 0; return 0; }
    h->Ttab.assign(T, T + B * nstops);
    h->Btab = B;
    h->prepared = false;
    return 0;
}

extern "C" int32_t kb2_set_stops(kb2_handle h, int64_t nstops, const double *stop_t, const int32_t *flags)
{
    if (!h) return 1;
    if (nstops <= 0) FAIL(h, "need at least one stop (the end of tspan)");
    for (int64_t s = 1; s < nstops; ++s) if (!(stop_t[s] > stop_t[s - 1])) FAIL(h, "stops must be strictly increasing");
    h->stop_t.assign(stop_t, stop_t + nstops);
    h->stop_flags.assign(flags, flags + nstops);
    h->stop_cnt.clear();
    h->ns_row = nstops; h->Bstops = 0;
    h->prepared = false;
    return 0;
}

extern "C" int32_t kb2_set_member_stops(kb2_handle h, int64_t B, int64_t nstops_max, const int32_t *counts,
                                        const double *stop_t, const int32_t *flags)
{
    if (!h) return 1;
    if (B <= 0 || nstops_max <= 0) FAIL(h, "need at least one member and one stop");
    for (int64_t b = 0; b < B; ++b) {
        if (counts[b] <= 0 || counts[b] > nstops_max) FAIL(h, "bad per-member stop count");
        for (int64_t s = 1; s < counts[b]; ++s)
            if (!(stop_t[b * nstops_max + s] > stop_t[b * nstops_max + s - 1])) FAIL(h, "stops must be strictly increasing");
    }
    h->stop_t.assign(stop_t, stop_t + B * nstops_max);
    h->stop_flags.assign(flags, flags + B * nstops_max);
    h->stop_cnt.assign(counts, counts + B);
    h->ns_row = nstops_max; h->Bstops = B;
    h->prepared = false;
    return 0;
}

// ---- ensemble state -------------------------------------------------------------------------
static int pick_mb(kb2_ctx *h, int64_t B)
{
    if (h->mb_user) return h->mb_user;
    if (const char *ev = getenv("KB2_MB")) {       // testing / tuning knob, same meaning as kb2_set_tiling
        const int v = atoi(ev);
        if (v == 1 || v == 2 || v == 4) return v;
    }
    // four members per warp tile (32-byte sectors fully used) unless the ensemble is too small to
    // give every SM two tiles (the sweeps run one warp per tile).  Measured on C4 (1024 members):
    // 91 / 76 / 77 ms per attempted step with 1 / 2 / 4 members per tile.
    int mb = 4;
    while (mb > 1 && (B + mb - 1) / mb < (int64_t)2 * h->sm_count) mb >>= 1;
    return mb;
}

static int ensure_ensemble(kb2_ctx *h, int64_t B, int64_t Ns)
{
    if (h->device < 0) FAIL(h, "host-only handle: no CUDA device, and there is no CPU fallback");
    if (!h->net_on_device) FAIL(h, "run kb2_symbolic first");
    if (B <= 0 || B >= (1 << 30)) FAIL(h, "bad ensemble size");
    if (h->ens_B == B && h->ens_Ns >= Ns) return 0;
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaStreamSynchronize(h->stream));
    free_pool(h->ens_allocs);
    h->ens_B = -1;
    h->prepared = false;          // the stop / profile tables of a prepared solve lived in this pool
    DevEns &e = h->de;
    e = DevEns{};
    e.B = (int)B;
    e.Bp = (int)((B + 31) / 32 * 32);       // member count of the per-member tables (multiple of every tile size)
    e.MB = pick_mb(h, B);
    const size_t Bt = (size_t)(B + e.MB - 1) / e.MB * e.MB, S = h->net.S, R = h->net.R, Bp = e.Bp;
    auto &P = h->ens_allocs;
    int rc = 0;
    rc |= dev_alloc(h, P, S * Bt, &e.u);
    rc |= dev_alloc(h, P, S * Bt, &e.ua);
    rc |= dev_alloc(h, P, S * Bt, &e.rv);
    rc |= dev_alloc(h, P, S * Bt, &e.y);
    for (int q = 0; q < 6; ++q) rc |= dev_alloc(h, P, S * Bt, &e.K[q]);
    rc |= dev_alloc(h, P, R * Bt, &e.k);
    rc |= dev_alloc(h, P, R * Bt, &e.rate);
    rc |= dev_alloc(h, P, R * Bt * (size_t)h->sym.jslots, &e.drate);
    rc |= dev_alloc(h, P, (size_t)h->sym.panels.padded * Bt, &e.lu);
    rc |= dev_alloc(h, P, (size_t)std::max<int64_t>(h->sym.nnzJ, 1) * Bt, &e.jv);
    rc |= dev_alloc(h, P, S * Bt, &e.uc);
    if (h->continuous) {
        rc |= dev_alloc(h, P, R * Bt, &e.kdot);
        rc |= dev_alloc(h, P, S * Bt, &e.ft);
    }
    rc |= dev_alloc(h, P, S * Bt, &e.invd);
    rc |= dev_alloc(h, P, (size_t)std::max<int64_t>(Ns, 1) * S * Bt, &e.out_u);
    rc |= dev_alloc(h, P, S * Bt, &e.out_umax);
    rc |= dev_alloc(h, P, Bp, &e.status);
    rc |= dev_alloc(h, P, Bp * 8, &e.stats);
    rc |= dev_alloc(h, P, Bp, &h->scal);
    rc |= dev_alloc(h, P, Bp, &e.ctl);
    rc |= dev_alloc(h, P, (size_t)KB2_FLAG_SLOTS, &e.flags);
    if (rc) { free_pool(h->ens_allocs); return rc; }
    e.Ns = (int)Ns;
    {
        bool fits = false;
        warp_smem_bytes(e.MB, h->net.S, &fits);
        e.u_smem = fits ? 1 : 0;
#ifdef KB2_NO_USMEM
        e.u_smem = 0;
#endif
    }
    {
        // window LU: one CTA per tile with the active submatrix in shared memory, if it fits
        // members per CTA (a divisor of the tile size): the largest whose window fits an SM (measured
        // on C3: 4 members x 1 CTA per SM 8.5 ms, 2 x 2 CTAs 8.7 ms, 1 x 3 CTAs 13 ms: the update is
        // bound by shared-memory bandwidth, which more resident warps do not add to)
        const FrontPlan &f = h->sym.fronts;
        const size_t cap = h->smem_optin;
        int mw = 0;
        for (int c = e.MB; c >= 1 && !mw; c >>= 1) if (wl_smem_bytes(c, f.Wr, f.Wc, f.max_nl, f.max_nu) <= cap) mw = c;
        if (const char *ev = getenv("KB2_MW")) {       // tuning knob
            const int v = atoi(ev);
            if ((v == 1 || v == 2 || v == 4) && v <= e.MB && wl_smem_bytes(v, f.Wr, f.Wc, f.max_nl, f.max_nu) <= cap) mw = v;
        }
        h->window_ok = f.ready && mw > 0;
        h->window_mw = mw;
        if (const char *ev = getenv("KB2_LU")) if (!strcmp(ev, "panel")) h->window_ok = false;   // A/B switch: left-looking block plan
    }
    h->ens_B = B; h->ens_Ns = Ns; h->ens_mb = e.MB;
    h->ens_fixed = P.size();
    return 0;
}

#define DISPATCH_MW(mb, mw, ...)                                                       \
    switch ((mb) * 8 + (mw)) {                                                         \
    case 4 * 8 + 4: { constexpr int MB = 4, MW = 4; __VA_ARGS__; } break;              \
    case 4 * 8 + 2: { constexpr int MB = 4, MW = 2; __VA_ARGS__; } break;              \
    case 4 * 8 + 1: { constexpr int MB = 4, MW = 1; __VA_ARGS__; } break;              \
    case 2 * 8 + 2: { constexpr int MB = 2, MW = 2; __VA_ARGS__; } break;              \
    case 2 * 8 + 1: { constexpr int MB = 2, MW = 1; __VA_ARGS__; } break;              \
    default: { constexpr int MB = 1, MW = 1; __VA_ARGS__; } break;                     \
    }

#define DISPATCH_MB(mb, ...)                                          \
    switch (mb) {                                                     \
    case 1: { constexpr int MB = 1; __VA_ARGS__; } break;             \
    case 2: { constexpr int MB = 2; __VA_ARGS__; } break;             \
    default: { constexpr int MB = 4; __VA_ARGS__; } break;            \
    }

static int ensure_stage(kb2_ctx *h, size_t doubles)
{
    if (h->stage_cap >= doubles) return 0;
    if (h->stage) { CU(h, cudaStreamSynchronize(h->stream)); cudaFree(h->stage); h->stage = nullptr; h->stage_cap = 0; }
    CU(h, cudaMalloc((void **)&h->stage, std::max<size_t>(doubles, 1) * 8));
    h->stage_cap = doubles;
    return 0;
}

static int conv_grid(kb2_ctx *h, size_t n) { return (int)std::min<size_t>((n + 255) / 256, (size_t)h->sm_count * 16); }

// host rows [nrows][B] -> device tiles [tile][nrows][MB]
static int up_tiles(kb2_ctx *h, double *tiles, const double *src, size_t nrows, size_t B)
{
    const DevEns &e = h->de;
    int rc = ensure_stage(h, nrows * B);
    if (rc) return rc;
    CU(h, cudaMemcpyAsync(h->stage, src, nrows * B * 8, cudaMemcpyHostToDevice, h->stream));
    const int Bt = (int)((B + e.MB - 1) / e.MB * e.MB);
    DISPATCH_MB(e.MB, (k_rows_to_tiles<MB><<<conv_grid(h, nrows * Bt), 256, 0, h->stream>>>(h->stage, tiles, nrows, (int)B, Bt, B)));
    h->launches++;
    CU(h, cudaGetLastError());
    return 0;
}

// device tiles -> host rows [nrows][B]
static int down_tiles(kb2_ctx *h, double *dst, const double *tiles, size_t nrows, size_t B)
{
    const DevEns &e = h->de;
    int rc = ensure_stage(h, nrows * B);
    if (rc) return rc;
    DISPATCH_MB(e.MB, (k_tiles_to_rows<MB><<<conv_grid(h, nrows * B), 256, 0, h->stream>>>(tiles, h->stage, nrows, (int)B)));
    h->launches++;
    CU(h, cudaGetLastError());
    CU(h, cudaMemcpyAsync(dst, h->stage, nrows * B * 8, cudaMemcpyDeviceToHost, h->stream));
    return 0;
}

static int up_scal(kb2_ctx *h, const double *v, int64_t B, double fill)
{
    std::vector<double> p(h->de.Bp, fill);
    std::copy(v, v + B, p.begin());
    CU(h, cudaMemcpyAsync(h->scal, p.data(), p.size() * 8, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));     // p goes out of scope
    return 0;
}

static int n_tiles(const DevEns &e) { return (e.B + e.MB - 1) / e.MB; }

template <class K>
static int set_smem(kb2_ctx *h, K kern, size_t smem)
{
    if (smem > 48 * 1024) CU(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return 0;
}

// ---- kernel-level entry points --------------------------------------------------------------
extern "C" int32_t kb2_eval_k(kb2_handle h, int64_t B, const double *T, double *k_out)
{
    if (!h) return 1;
    if (h->calc_mode != 0) FAIL(h, "kb2_eval_k needs an Arrhenius calculator");
    int rc = ensure_ensemble(h, B, 1);
    if (rc) return rc;
    DevEns &e = h->de;
    if ((rc = up_scal(h, T, B, 300.0))) return rc;
    const int ntiles = n_tiles(e);
    DISPATCH_MB(e.MB, (k_rates<MB><<<std::min(ntiles, 32 * h->sm_count), 32, 0, h->stream>>>(h->dn, h->dp, e, h->scal, ntiles)));
    h->launches++;
    CU(h, cudaGetLastError());
    if ((rc = down_tiles(h, k_out, e.k, h->net.R, B))) return rc;
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int32_t kb2_eval_profile(kb2_handle h, int64_t B, int64_t nt, const double *t, double *X_out)
{
    if (!h) return 1;
    if (h->Bprof != B) FAIL(h, "kb2_set_profiles must be called with the same B first");
    if (h->device < 0) FAIL(h, "host-only handle: no CUDA device, and there is no CPU fallback");
    CU(h, cudaSetDevice(h->device));
    std::vector<void *> pool;
    const int32_t *dk; const double *dp, *dt; double *dx;
    int rc = dev_upload(h, pool, h->pkind.data(), (size_t)B, &dk);
    rc |= dev_upload(h, pool, h->pparams.data(), (size_t)B * 16, &dp);
    rc |= dev_upload(h, pool, t, (size_t)nt, &dt);
    rc |= dev_alloc(h, pool, (size_t)(B * nt), &dx);
    if (rc) { free_pool(pool); return rc; }
    const int n = (int)(B * nt);
    k_profile<<<(n + 255) / 256, 256, 0, h->stream>>>((int)B, (int)nt, dk, dp, dt, dx);
    h->launches++;
    cudaError_t ce = cudaMemcpyAsync(X_out, dx, (size_t)n * 8, cudaMemcpyDeviceToHost, h->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(h->stream);
    free_pool(pool);
    if (ce != cudaSuccess) FAIL(h, cudaGetErrorString(ce));
    return 0;
}

static int upload_uk(kb2_ctx *h, int64_t B, const double *u, const double *k)
{
    DevEns &e = h->de;
    int rc = 0;
    h->prepared = false;          // the kernel-level entry points overwrite the state of a prepared solve
    if (u) rc |= up_tiles(h, e.u, u, h->net.S, B);
    if (k) rc |= up_tiles(h, e.k, k, h->net.R, B);
    return rc;
}

// shared memory of a streaming-phase CTA: the tile's state vector when it is staged
static size_t stream_smem(kb2_ctx *h) { return h->de.u_smem ? (size_t)h->net.S * h->de.MB * 8 : 0; }

template <class K>
static int stream_grid(kb2_ctx *h, K kern, int threads, size_t smem, int ntiles, int *grid)
{
    int r = set_smem(h, kern, smem);
    if (r) return r;
    int per_sm = 0;
    CU(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    if (per_sm < 1) FAIL(h, "a phase kernel does not fit on an SM");
    *grid = std::min(ntiles, per_sm * h->sm_count);
    return 0;
}

static int launch_rhs(kb2_ctx *h)
{
    DevEns &e = h->de;
    const int ntiles = n_tiles(e);
    const size_t smem = stream_smem(h);
    DISPATCH_MB(e.MB, {
        int grid = 0;
        int r = stream_grid(h, k_rhs<MB, RHS_NW>, RHS_NW * 32, smem, ntiles, &grid);
        if (r) return r;
        k_rhs<MB, RHS_NW><<<grid, RHS_NW * 32, smem, h->stream>>>(h->dn, h->dp, e, ntiles);
    });
    h->launches++;
    CU(h, cudaGetLastError());
    return 0;
}

extern "C" int32_t kb2_eval_rhs(kb2_handle h, int64_t B, const double *u, const double *k, double *du)
{
    if (!h) return 1;
    int rc = ensure_ensemble(h, B, 1);
    if (rc) return rc;
    if ((rc = upload_uk(h, B, u, k))) return rc;
    DevEns &e = h->de;
    const int ntiles = n_tiles(e);
    if ((rc = launch_rhs(h))) return rc;
    if ((rc = down_tiles(h, du, e.rv, h->net.S, B))) return rc;
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}

static int launch_jac(kb2_ctx *h, int use_ctl)
{
    DevEns &e = h->de;
    const int ntiles = n_tiles(e);
    const size_t smem = stream_smem(h);
    DISPATCH_MB(e.MB, {
        int grid = 0;
        int r = stream_grid(h, k_step_jac<MB, RHS_NW>, RHS_NW * 32, smem, ntiles, &grid);
        if (r) return r;
        k_step_jac<MB, RHS_NW><<<grid, RHS_NW * 32, smem, h->stream>>>(h->dn, h->dp, e, ntiles, use_ctl);
    });
    h->launches++;
    CU(h, cudaGetLastError());
    return 0;
}

// hg: per-member 1/(h*gamma) on the device, or null (taken from the control state of the solve)
static int window_launch_shape(kb2_ctx *h, size_t *smem, int *grid)
{
    DevEns &e = h->de;
    const int mw = h->window_mw, nwork = n_tiles(e) * (e.MB / mw);
    *smem = wl_smem_bytes(mw, h->df.Wr, h->df.Wc, h->df.max_nl, h->df.max_nu);
    int per_sm = 0;
    DISPATCH_MW(e.MB, mw, {
        int r = set_smem(h, k_lu_window<MB, MW>, *smem);
        if (r) return r;
        CU(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_lu_window<MB, MW>, WL_NT, *smem));
    });
    if (per_sm < 1) FAIL(h, "the window LU kernel does not fit on an SM");
    *grid = std::min(nwork, per_sm * h->sm_count);
    h->last_window_ctas = per_sm;
    h->window_stagger_ns = 0;      // start delay of the second CTA of an SM: measured, no effect (KB2_WL_STAGGER to experiment)
    if (const char *ev = getenv("KB2_WL_STAGGER")) h->window_stagger_ns = atoi(ev);
    return 0;
}

// hg: per-member 1/(h*gamma) on the device, or null (taken from the control state of the solve)
static int launch_window_lu(kb2_ctx *h, const double *d_hg)
{
    DevEns &e = h->de;
    size_t smem = 0;
    int grid = 0;
    int rc = window_launch_shape(h, &smem, &grid);
    if (rc) return rc;
    DISPATCH_MW(e.MB, h->window_mw, (k_lu_window<MB, MW><<<grid, WL_NT, smem, h->stream>>>(h->dn, h->dp, h->df, e, d_hg, n_tiles(e), h->window_stagger_ns)));
    h->launches++;
    CU(h, cudaGetLastError());
    return 0;
}

extern "C" int32_t kb2_eval_jac(kb2_handle h, int64_t B, const double *u, const double *k, double *Jval)
{
    if (!h) return 1;
    int rc = ensure_ensemble(h, B, 1);
    if (rc) return rc;
    if ((rc = upload_uk(h, B, u, k))) return rc;
    DevEns &e = h->de;
    if ((rc = launch_jac(h, 0))) return rc;
    if ((rc = down_tiles(h, Jval, e.jv, (size_t)h->sym.nnzJ, B))) return rc;
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}

static int launch_factor(kb2_ctx *h, const double *d_hg, int mode = 3)
{
    DevEns &e = h->de;
    const int ntiles = n_tiles(e);
    const size_t smem = warp_smem_bytes(e.MB, h->net.S, nullptr) + 16;
    DISPATCH_MB(e.MB, {
        int r = set_smem(h, k_factor<MB>, smem);
        if (r) return r;
        k_factor<MB><<<std::min(ntiles, 32 * h->sm_count), 32, smem, h->stream>>>(h->dn, h->dp, e, d_hg, ntiles, mode, (int)smem - 16);
    });
    h->launches++;
    CU(h, cudaGetLastError());
    return 0;
}

static int launch_trisolve(kb2_ctx *h)
{
    DevEns &e = h->de;
    const int ntiles = n_tiles(e);
    const size_t smem = warp_smem_bytes(e.MB, h->net.S, nullptr) + 16;
    DISPATCH_MB(e.MB, {
        int r = set_smem(h, k_trisolve<MB>, smem);
        if (r) return r;
        k_trisolve<MB><<<std::min(ntiles, 32 * h->sm_count), 32, smem, h->stream>>>(h->dn, h->dp, e, ntiles, (int)smem - 16);
    });
    h->launches++;
    CU(h, cudaGetLastError());
    return 0;
}

extern "C" int32_t kb2_factor(kb2_handle h, int64_t B, const double *u, const double *k,
                              const double *hg_inv, double *lu_out)
{
    if (!h) return 1;
    int rc = ensure_ensemble(h, B, 1);
    if (rc) return rc;
    if ((rc = upload_uk(h, B, u, k))) return rc;
    DevEns &e = h->de;
    if ((rc = up_scal(h, hg_inv, B, 1.0))) return rc;
    if (h->window_ok) {
        if ((rc = launch_jac(h, 0))) return rc;
        if ((rc = launch_window_lu(h, h->scal))) return rc;
    } else if ((rc = launch_factor(h, h->scal))) return rc;
    if (lu_out) {
        // device storage is the padded block layout: bring it back and gather the exact pattern
        const PanelPlan &pp = h->sym.panels;
        std::vector<double> tmp((size_t)pp.padded * B);
        if ((rc = down_tiles(h, tmp.data(), e.lu, (size_t)pp.padded, B))) return rc;
        CU(h, cudaStreamSynchronize(h->stream));
        for (int64_t q = 0; q < h->sym.nnzLU; ++q)
            std::copy(tmp.begin() + (size_t)pp.slot_of[q] * B, tmp.begin() + ((size_t)pp.slot_of[q] + 1) * B, lu_out + q * B);
    }
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int32_t kb2_trisolve(kb2_handle h, int64_t B, const double *rhs, double *x)
{
    if (!h) return 1;
    if (h->ens_B != B) FAIL(h, "kb2_trisolve must follow kb2_factor with the same B");
    DevEns &e = h->de;
    int rc = up_tiles(h, e.rv, rhs, h->net.S, B);
    if (rc) return rc;
    if ((rc = launch_trisolve(h))) return rc;
    if ((rc = down_tiles(h, x, e.ua, h->net.S, B))) return rc;
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int32_t kb2_time_kernel(kb2_handle h, int32_t which, int64_t B, int32_t iters, float *ms_avg)
{
    if (!h) return 1;
    if (h->ens_B != B) FAIL(h, "call an eval/solve entry point with this B first so data is resident");
    DevEns &e = h->de;
    const int ntiles = n_tiles(e);
    const int grid = std::min(ntiles, 32 * h->sm_count);
    if (which == 0 || which == 3 || (which >= 5 && which <= 8)) {
        std::vector<double> c(e.Bp, which == 0 ? 1000.0 : 1.0e6);
        int rc = up_scal(h, c.data(), e.Bp, 1.0);
        if (rc) return rc;
    }
    for (int it = -2; it < iters; ++it) {
        if (it == 0) CU(h, cudaEventRecord(h->ev0, h->stream));
        int rc = 0;
        switch (which) {
        case 0: DISPATCH_MB(e.MB, (k_rates<MB><<<grid, 32, 0, h->stream>>>(h->dn, h->dp, e, h->scal, ntiles))); h->launches++; break;
        case 1: rc = launch_rhs(h); break;
        case 2: rc = launch_jac(h, 0); break;
        case 3: rc = h->window_ok ? (launch_jac(h, 0) || launch_window_lu(h, h->scal)) : launch_factor(h, h->scal); break;
        case 4: rc = launch_trisolve(h); break;
        case 5: rc = launch_factor(h, h->scal, 1); break;   // W assembly only
        case 6: rc = launch_factor(h, h->scal, 2); break;   // block-plan LU only (on whatever the storage holds)
        case 7: if (!h->window_ok) FAIL(h, "the window LU does not fit"); rc = launch_window_lu(h, h->scal); break;   // window LU only
        case 8: rc = launch_factor(h, h->scal); break;      // block-plan W assembly + LU (the fallback path)
        default: FAIL(h, "unknown kernel id");
        }
        if (rc) return rc;
    }
    CU(h, cudaEventRecord(h->ev1, h->stream));
    CU(h, cudaEventSynchronize(h->ev1));
    float ms = 0;
    CU(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    CU(h, cudaGetLastError());
    if (ms_avg) *ms_avg = ms / (float)iters;
    return 0;
}

// ---- the solve ------------------------------------------------------------------------------
// Prepare members [b0, b0 + B) of an ensemble of Btot members whose per-member tables (profiles,
// stops, T table) were set for all Btot (batch tiling: kb2_solve walks the ensemble in chunks that
// fit the device memory; the three-phase API prepares everything at once, b0 = 0, B = Btot).
static int prepare_range(kb2_ctx *h, int64_t b0, int64_t B, int64_t Btot, const double *u0, int64_t u0_stride, double t0,
                         double abstol, double reltol, double dtmin, int64_t maxiters, int32_t ban_negatives, int64_t Ns)
{
    h->prepared = false;
    if (h->calc_mode < 0) FAIL(h, "no calculator set");
    if (h->stop_t.empty()) FAIL(h, "no stops set");
    if (h->calc_mode == 0 && h->Bprof != Btot) FAIL(h, "kb2_set_profiles must be called with the same B");
    const int64_t nstops = h->ns_row;
    const bool shared = h->stop_cnt.empty();
    if (!shared && h->Bstops != Btot) FAIL(h, "kb2_set_member_stops was called with a different B");
    if (!h->Ttab.empty() && (h->Btab != Btot || (int64_t)h->Ttab.size() != Btot * nstops))
        FAIL(h, "T table does not match B x nstops");
    if (!(abstol > 0) || !(reltol > 0)) FAIL(h, "tolerances must be positive");
    const int64_t Bp64 = (B + 31) / 32 * 32;
    // expand to per-member tables [b][nstops]
    std::vector<double> st((size_t)(Bp64 * nstops), 0.0);
    std::vector<int32_t> sf((size_t)(Bp64 * nstops), 0), ridx((size_t)(Bp64 * nstops), -1), cnt((size_t)Bp64, 0);
    bool chunked = false;
    for (int64_t b = 0; b < B; ++b) {
        const double *ts = shared ? h->stop_t.data() : h->stop_t.data() + (b0 + b) * nstops;
        const int32_t *fl = shared ? h->stop_flags.data() : h->stop_flags.data() + (b0 + b) * nstops;
        const int64_t n = shared ? nstops : h->stop_cnt[b0 + b];
        int64_t nsave = 0, nrate = 0;
        for (int64_t s = 0; s < n; ++s) {
            if (ts[s] < t0) FAIL(h, "stops must not precede t0");
            st[b * nstops + s] = ts[s];
            sf[b * nstops + s] = fl[s];
            if (fl[s] & KB2_STOP_SAVE) ++nsave;
            if (fl[s] & KB2_STOP_CHUNK) chunked = true;
            if (fl[s] & KB2_STOP_RATE) ridx[b * nstops + s] = (int32_t)nrate++;
        }
        cnt[b] = (int32_t)n;
        if (nsave != Ns) FAIL(h, "Ns does not match the number of save stops");
        if (h->calc_mode == 1 && nrate != h->n_rate_stops) FAIL(h, "rate table length does not match the rate-update stops");
        if (shared && b == 0 && B > 1) {       // replicate member 0
            for (int64_t bb = 1; bb < B; ++bb) {
                std::copy(st.begin(), st.begin() + nstops, st.begin() + bb * nstops);
                std::copy(sf.begin(), sf.begin() + nstops, sf.begin() + bb * nstops);
                std::copy(ridx.begin(), ridx.begin() + nstops, ridx.begin() + bb * nstops);
                cnt[bb] = cnt[0];
            }
            break;
        }
    }
    int rc = ensure_ensemble(h, B, Ns);
    if (rc) return rc;
    DevEns &e = h->de;
    const size_t Bp = e.Bp, S = h->net.S;
    // u0 -> tile-major [tile][S][MB]
    {
        const size_t MB = e.MB, Bt = (size_t)(B + e.MB - 1) / e.MB * e.MB;
        std::vector<double> up(S * Bt, 0.0);
        for (int64_t b = 0; b < B; ++b) {
            const double *src = u0 + (size_t)(u0_stride ? (b0 + b) * u0_stride : 0);
            double *dst = up.data() + (size_t)(b / MB) * S * MB + b % MB;
            for (size_t i = 0; i < S; ++i) dst[i * MB] = src[i];
        }
        CU(h, cudaMemcpyAsync(e.u, up.data(), S * Bt * 8, cudaMemcpyHostToDevice, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
    }
    // per-solve condition tables live at the tail of the ensemble pool: drop the previous ones
    while (h->ens_allocs.size() > h->ens_fixed) { cudaFree(h->ens_allocs.back()); h->ens_allocs.pop_back(); }
    auto &P = h->ens_allocs;
    std::vector<int32_t> kind(Bp, 0);
    std::vector<double> par(Bp * 16, 0.0);
    if (h->calc_mode == 0) {
        std::copy(h->pkind.begin() + b0, h->pkind.begin() + b0 + B, kind.begin());
        std::copy(h->pparams.begin() + b0 * 16, h->pparams.begin() + (b0 + B) * 16, par.begin());
    }
    rc |= dev_upload(h, P, kind.data(), kind.size(), &e.pkind);
    rc |= dev_upload(h, P, par.data(), par.size(), &e.pparams);
    rc |= dev_upload(h, P, st.data(), st.size(), &e.stop_t);
    rc |= dev_upload(h, P, sf.data(), sf.size(), &e.stop_flags);
    rc |= dev_upload(h, P, ridx.data(), ridx.size(), &e.stop_ridx);
    rc |= dev_upload(h, P, cnt.data(), cnt.size(), &e.stop_cnt);
    e.Ttab = nullptr;
    if (!h->Ttab.empty()) {
        std::vector<double> tp(Bp * nstops, NAN);
        std::copy(h->Ttab.begin() + b0 * nstops, h->Ttab.begin() + (b0 + B) * nstops, tp.begin());
        rc |= dev_upload(h, P, tp.data(), tp.size(), &e.Ttab);
    }
    if (rc) return rc;
    CU(h, cudaStreamSynchronize(h->stream));
    e.nstops = (int)nstops; e.Ns = (int)Ns;
    e.t0 = t0; e.abstol = abstol; e.reltol = reltol; e.dtmin = dtmin; e.maxiters = maxiters;
    e.ban_neg = ban_negatives;
    e.chunk_retry = h->chunk_retry; e.update_tols = h->chunk_update_tols;
    e.continuous = h->continuous;
    if (h->continuous && h->calc_mode != 0) FAIL(h, "continuous rate updates need a calculator with a device form (kb2_set_arrhenius)");
    h->has_chunk_stops = chunked;
    h->prepared = true;
    return 0;
}

extern "C" int32_t kb2_set_continuous(kb2_handle h, int32_t continuous)
{
    if (!h) return 1;
    const int v = continuous ? 1 : 0;
    if (v != h->continuous) { h->continuous = v; h->ens_B = -1; }      // the ensemble carries two more arrays in continuous mode
    h->prepared = false;
    return 0;
}

extern "C" int32_t kb2_set_chunking(kb2_handle h, int32_t retry_failed_chunks, int32_t update_tols)
{
    if (!h) return 1;
    h->chunk_retry = retry_failed_chunks ? 1 : 0;
    h->chunk_update_tols = update_tols ? 1 : 0;
    h->prepared = false;
    return 0;
}

extern "C" int32_t kb2_solve_prepare(kb2_handle h, int64_t B, const double *u0, int64_t u0_stride, double t0,
                                     double abstol, double reltol, double dtmin, int64_t maxiters,
                                     int32_t ban_negatives, int64_t Ns)
{
    if (!h) return 1;
    return prepare_range(h, 0, B, B, u0, u0_stride, t0, abstol, reltol, dtmin, maxiters, ban_negatives, Ns);
}

// Device bytes one member of the current network needs during a solve with Ns save points, and how
// many members fit the free device memory at once (batch tile; 0 = kb2_set_batch_tile override unset).
extern "C" int32_t kb2_memory_plan(kb2_handle h, int64_t Ns, int64_t *bytes_per_member, int64_t *b_tile, int64_t *free_bytes)
{
    if (!h) return 1;
    if (h->device < 0) FAIL(h, "host-only handle: no CUDA device");
    if (!h->net_on_device) FAIL(h, "run kb2_symbolic first");
    const int64_t S = h->net.S, R = h->net.R;
    // LU values, Jacobian values, derivative table, k + rate, 11 state-sized vectors + maxima, saves
    // (twice: the layout conversion of the fetch stages them), control state and tables
    const int64_t per = 8 * (h->sym.panels.padded + std::max<int64_t>(h->sym.nnzJ, 1) + (int64_t)h->sym.jslots * R + 2 * R + 13 * S +
                             2 * std::max<int64_t>(Ns, 1) * S) + (int64_t)sizeof(Ctl) + 2048;
    CU(h, cudaSetDevice(h->device));
    size_t fr = 0, tot = 0;
    CU(h, cudaMemGetInfo(&fr, &tot));
    // memory this handle already holds for an ensemble is reusable
    int64_t tile = (int64_t)((double)fr * 0.9 / (double)per);
    tile = tile / 128 * 128;
    if (h->b_tile_user > 0) tile = h->b_tile_user;
    if (bytes_per_member) *bytes_per_member = per;
    if (b_tile) *b_tile = tile;
    if (free_bytes) *free_bytes = (int64_t)fr;
    return 0;
}

extern "C" int32_t kb2_set_batch_tile(kb2_handle h, int64_t b_tile)
{
    if (!h) return 1;
    if (b_tile < 0) FAIL(h, "batch tile must be >= 0 (0 = sized from the free device memory)");
    h->b_tile_user = b_tile;
    return 0;
}

// launch shape of a phase kernel: one warp per CTA, every tile resident when the grid allows it
template <class K>
static int phase_grid(kb2_ctx *h, K kern, size_t smem, int ntiles, int *grid)
{
    int r = set_smem(h, kern, smem);
    if (r) return r;
    int per_sm = 0;
    CU(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem));
    if (per_sm < 1) FAIL(h, "a phase kernel does not fit on an SM");
    *grid = std::min(ntiles, per_sm * h->sm_count);
    return 0;
}

enum { PH_LU = 0, PH_RHS = 1, PH_SWEEP = 2, PH_END = 3, PH_JAC = 4, PH_COUNT = 5 };

extern "C" int32_t kb2_solve_run(kb2_handle h, float *ms_device)
{
    if (!h) return 1;
    if (!h->prepared) FAIL(h, "kb2_solve_prepare has not succeeded");
    CU(h, cudaSetDevice(h->device));
    DevEns &e = h->de;
    const int ntiles = n_tiles(e);
    const size_t smem = warp_smem_bytes(e.MB, h->net.S, nullptr) + 16;
    const int data_bytes = (int)smem - 16;
    int g_init = 0, g_lu = 0, g_rhs = 0, g_sweep = 0, g_end = 0, g_jac = 0, g_wl = 0;
    const size_t smem_st = stream_smem(h);
    const size_t smem_end = smem_st + (size_t)END_NW * e.MB * 8;      // + the cross-warp reduction buffer
    // right-hand side with the tile's rate table in shared memory, when state vector + table fit an SM
    const size_t smem_rs = (size_t)(h->net.S + h->net.R) * e.MB * 8;
    bool rhs_rs = e.u_smem && smem_rs + 1024 <= h->smem_optin;
    if (const char *ev = getenv("KB2_RHS_RS")) rhs_rs = rhs_rs && atoi(ev) != 0;      // A/B switch
    const bool window = h->window_ok;
    size_t smem_wl = 0;
    if (window) { int r = window_launch_shape(h, &smem_wl, &g_wl); if (r) return r; }
    DISPATCH_MB(e.MB, {
        int r = stream_grid(h, k_solve_init<MB, END_NW>, END_NW * 32, smem_end, ntiles, &g_init);
        if (!r) r = phase_grid(h, k_step_lu<MB>, smem, ntiles, &g_lu);
        if (!r) r = stream_grid(h, k_step_jac<MB, RHS_NW>, RHS_NW * 32, smem_st, ntiles, &g_jac);

        if (!r && rhs_rs) r = stream_grid(h, k_stage_rhs<MB, RHS_NW_RS, true>, RHS_NW_RS * 32, smem_rs, ntiles, &g_rhs);
        if (!r && !rhs_rs) r = stream_grid(h, k_stage_rhs<MB, RHS_NW, false>, RHS_NW * 32, smem_st, ntiles, &g_rhs);
        if (!r) r = phase_grid(h, k_stage_sweep<MB>, smem, ntiles, &g_sweep);
        if (!r) r = stream_grid(h, k_step_end<MB, END_NW>, END_NW * 32, smem_end, ntiles, &g_end);
        if (r) return r;
    });
    h->last_ctas_per_sm = (g_lu + h->sm_count - 1) / h->sm_count;
    cudaStream_t st = h->stream;
    // phase timing: one round in 64 is bracketed with events, kernel by kernel
    if (h->phase_ev.empty()) {
        h->phase_ev.resize(16);
        for (auto &ev : h->phase_ev) CU(h, cudaEventCreate(&ev));
    }
    for (int q = 0; q < PH_COUNT; ++q) { h->phase_ms[q] = 0.0; h->phase_n[q] = 0; }
    h->rounds = 0;
    CU(h, cudaMemsetAsync(e.flags, 0, KB2_FLAG_SLOTS * sizeof(int), st));
    CU(h, cudaEventRecord(h->ev0, st));
    DISPATCH_MB(e.MB, (k_solve_init<MB, END_NW><<<g_init, END_NW * 32, smem_end, st>>>(h->dn, h->dp, e, ntiles)));
    h->launches++;
    CU(h, cudaGetLastError());
    // rounds are launched in batches; flags[j] of a batch = some member is still running after
    // round j of it.  A round on a finished ensemble is thirteen empty kernels.
    const int NB = 16;
    // every running member counts each round against maxiters; chunkwise solves count per chunk
    // (and per repeat of a chunk), so the loop's own limit is per stop
    const long long max_rounds = (e.maxiters + 2) * (h->has_chunk_stops ? (long long)e.nstops * 5 : 1);
    int *hflag = h->h_flag;
    bool running = true;
    {
        CU(h, cudaMemcpyAsync(hflag, e.flags, sizeof(int), cudaMemcpyDeviceToHost, st));
        CU(h, cudaStreamSynchronize(st));
        running = hflag[0] != 0;
    }
    // one round = the phase kernels of one attempted step; `tm`: bracket every kernel with an event
    auto launch_round = [&](int j, bool tm) {
        int evi = 0;
        if (tm) cudaEventRecord(h->phase_ev[evi++], st);
        DISPATCH_MB(e.MB, {
            if (window) {
                k_step_jac<MB, RHS_NW><<<g_jac, RHS_NW * 32, smem_st, st>>>(h->dn, h->dp, e, ntiles, 1);
                if (tm) cudaEventRecord(h->phase_ev[evi++], st);
                DISPATCH_MW(MB, h->window_mw, (k_lu_window<MB, MW><<<g_wl, WL_NT, smem_wl, st>>>(h->dn, h->dp, h->df, e, (const double *)nullptr, ntiles, h->window_stagger_ns)));
            } else {
                if (tm) cudaEventRecord(h->phase_ev[evi++], st);
                k_step_lu<MB><<<g_lu, 32, smem, st>>>(h->dn, h->dp, e, ntiles, data_bytes);
            }
            if (tm) cudaEventRecord(h->phase_ev[evi++], st);
            for (int s = 0; s < 6; ++s) {
                if (rhs_rs) k_stage_rhs<MB, RHS_NW_RS, true><<<g_rhs, RHS_NW_RS * 32, smem_rs, st>>>(h->dn, h->dp, e, ntiles, s);
                else k_stage_rhs<MB, RHS_NW, false><<<g_rhs, RHS_NW * 32, smem_st, st>>>(h->dn, h->dp, e, ntiles, s);
                if (tm) cudaEventRecord(h->phase_ev[evi++], st);
                k_stage_sweep<MB><<<g_sweep, 32, smem, st>>>(h->dn, h->dp, e, ntiles, data_bytes, s);
                if (tm) cudaEventRecord(h->phase_ev[evi++], st);
            }
            k_step_end<MB, END_NW><<<g_end, END_NW * 32, smem_end, st>>>(h->dn, h->dp, e, ntiles, j);
            if (tm) cudaEventRecord(h->phase_ev[evi++], st);
        });
    };
    // A batch of NB rounds as ONE CUDA graph (memset of the flags, NB x 15 kernels, read-back of the
    // flags): a single small solve is launch-bound (C2: 14 000 chunks, ~2.7 M launches).  One batch in
    // four is launched kernel by kernel instead, with the phase-timing events around one of its rounds.
    cudaGraphExec_t gexec = nullptr;
    {
        bool use_graph = true;
        if (const char *ev = getenv("KB2_GRAPH")) use_graph = atoi(ev) != 0;
        cudaGraph_t g = nullptr;
        if (use_graph && cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            cudaMemsetAsync(e.flags, 0, KB2_FLAG_SLOTS * sizeof(int), st);
            for (int j = 0; j < NB; ++j) launch_round(j, false);
            cudaMemcpyAsync(hflag, e.flags, NB * sizeof(int), cudaMemcpyDeviceToHost, st);
            if (cudaStreamEndCapture(st, &g) != cudaSuccess || !g || cudaGraphInstantiate(&gexec, g, 0) != cudaSuccess) gexec = nullptr;
            if (g) cudaGraphDestroy(g);
        }
        cudaGetLastError();       // a failed capture falls back to plain launches
    }
    long long nbatch = 0;
    while (running && h->rounds < max_rounds) {
        const bool direct = !gexec || (nbatch++ % 4 == 1);
        if (direct) {
            CU(h, cudaMemsetAsync(e.flags, 0, KB2_FLAG_SLOTS * sizeof(int), st));
            for (int j = 0; j < NB; ++j) launch_round(j, j == NB / 2);
            CU(h, cudaGetLastError());
            CU(h, cudaMemcpyAsync(hflag, e.flags, NB * sizeof(int), cudaMemcpyDeviceToHost, st));
        } else {
            CU(h, cudaGraphLaunch(gexec, st));
        }
        h->launches += (long long)NB * (window ? 15 : 14);
        CU(h, cudaStreamSynchronize(st));
        int used = NB;
        for (int j = 0; j < NB; ++j) if (!hflag[j]) { used = j + 1; break; }
        h->rounds += used;
        running = hflag[NB - 1] != 0 && used == NB;
        if (direct) {
            float ms = 0;
            const int map[15] = {PH_JAC, PH_LU, PH_RHS, PH_SWEEP, PH_RHS, PH_SWEEP, PH_RHS, PH_SWEEP, PH_RHS, PH_SWEEP,
                                 PH_RHS, PH_SWEEP, PH_RHS, PH_SWEEP, PH_END};
            for (int q = window ? 0 : 1; q < 15; ++q) {
                CU(h, cudaEventElapsedTime(&ms, h->phase_ev[q], h->phase_ev[q + 1]));
                h->phase_ms[map[q]] += ms;
                h->phase_n[map[q]]++;
            }
        }
    }
    if (gexec) cudaGraphExecDestroy(gexec);
    if (running) {
        k_mark_unfinished<<<(e.B + 255) / 256, 256, 0, st>>>(e);
        h->launches++;
    }
    {
        const size_t n = (size_t)ntiles * h->net.S * e.MB;
        DISPATCH_MB(e.MB, (k_umax<MB><<<conv_grid(h, n), 256, 0, st>>>(e, (int)h->net.S, n)));
        h->launches++;
    }
    CU(h, cudaEventRecord(h->ev1, st));
    CU(h, cudaGetLastError());
    CU(h, cudaEventSynchronize(h->ev1));
    float ms = 0;
    CU(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    if (ms_device) *ms_device = ms;
    h->prepared = false;      // u has been advanced in place: prepare again before another run
    return 0;
}

extern "C" int32_t kb2_get_phase_times(kb2_handle h, double *ms_avg, int64_t *launches_sampled, int64_t *rounds)
{
    if (!h) return 1;
    for (int q = 0; q < PH_COUNT; ++q) {
        if (ms_avg) ms_avg[q] = h->phase_n[q] ? h->phase_ms[q] / (double)h->phase_n[q] : 0.0;
        if (launches_sampled) launches_sampled[q] = h->phase_n[q];
    }
    if (rounds) *rounds = h->rounds;
    return 0;
}

extern "C" int32_t kb2_solve_fetch(kb2_handle h, double *out_u, double *out_umax, int32_t *status, int64_t *stats)
{
    if (!h) return 1;
    if (h->ens_B <= 0) FAIL(h, "nothing to fetch");
    DevEns &e = h->de;
    const size_t B = e.B, S = h->net.S;
    int rc = 0;
    if (out_u) {
        if ((rc = down_tiles(h, out_u, e.out_u, (size_t)e.Ns * S, B))) return rc;
        CU(h, cudaStreamSynchronize(h->stream));      // the staging buffer is reused below
    }
    if (out_umax && (rc = down_tiles(h, out_umax, e.out_umax, S, B))) return rc;
    if (status) CU(h, cudaMemcpyAsync(status, e.status, B * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    if (stats) CU(h, cudaMemcpyAsync(stats, e.stats, B * 8 * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int32_t kb2_solve(kb2_handle h, int64_t B, const double *u0, int64_t u0_stride, double t0,
                             double abstol, double reltol, double dtmin, int64_t maxiters, int32_t ban_negatives,
                             int64_t Ns, double *out_u, double *out_umax, int32_t *status, int64_t *stats)
{
    if (!h) return 1;
    // batch tiling: an ensemble larger than the device memory is solved in chunks of b_tile members
    // (members are independent; the network tables, plan and calculator stay resident)
    int64_t per = 0, tile = 0;
    if (h->device < 0) FAIL(h, "host-only handle: no CUDA device, and there is no CPU fallback");
    if (h->ens_B != B) {          // an allocation of the right size is kept as it is
        CU(h, cudaSetDevice(h->device));
        CU(h, cudaStreamSynchronize(h->stream));
        free_pool(h->ens_allocs);
        h->ens_B = -1; h->prepared = false;
    }
    int rc = kb2_memory_plan(h, Ns, &per, &tile, nullptr);
    if (rc) return rc;
    if (h->ens_B == B && h->b_tile_user <= 0) tile = std::max(tile, B);      // resident already: it fits
    if (tile <= 0) FAIL(h, "not enough device memory for a single member of this network");
    h->last_tiles = (B + tile - 1) / tile;
    if (B <= tile) {
        rc = prepare_range(h, 0, B, B, u0, u0_stride, t0, abstol, reltol, dtmin, maxiters, ban_negatives, Ns);
        if (rc) return rc;
        if ((rc = kb2_solve_run(h, nullptr))) return rc;
        return kb2_solve_fetch(h, out_u, out_umax, status, stats);
    }
    const size_t S = h->net.S;
    std::vector<double> cu, cm;
    for (int64_t b0 = 0; b0 < B; b0 += tile) {
        const int64_t Bc = std::min(tile, B - b0);
        rc = prepare_range(h, b0, Bc, B, u0, u0_stride, t0, abstol, reltol, dtmin, maxiters, ban_negatives, Ns);
        if (rc) return rc;
        if ((rc = kb2_solve_run(h, nullptr))) return rc;
        if (out_u) cu.resize((size_t)Ns * S * Bc);
        if (out_umax) cm.resize(S * Bc);
        rc = kb2_solve_fetch(h, out_u ? cu.data() : nullptr, out_umax ? cm.data() : nullptr, status ? status + b0 : nullptr,
                             stats ? stats + b0 * 8 : nullptr);
        if (rc) return rc;
        if (out_u)
            for (size_t r = 0; r < (size_t)Ns * S; ++r) std::copy(cu.begin() + r * Bc, cu.begin() + (r + 1) * Bc, out_u + r * B + b0);
        if (out_umax)
            for (size_t r = 0; r < S; ++r) std::copy(cm.begin() + r * Bc, cm.begin() + (r + 1) * Bc, out_umax + r * B + b0);
    }
    return 0;
}

extern "C" int64_t kb2_last_batch_tiles(kb2_handle h) { return h ? h->last_tiles : 0; }

extern "C" int32_t kb2_pack_results_device(kb2_handle h, double *final_bs_dev, double *umax_bs_dev)
{
    if (!h) return 1;
    if (h->ens_B <= 0) FAIL(h, "nothing to pack");
    DevEns &e = h->de;
    const int S = (int)h->net.S;
    const int grid = conv_grid(h, (size_t)S * e.B);
    if (final_bs_dev) { DISPATCH_MB(e.MB, (k_pack_bs<MB><<<grid, 256, 0, h->stream>>>(S, e.B, e.u, final_bs_dev))); h->launches++; }
    if (umax_bs_dev) { DISPATCH_MB(e.MB, (k_pack_bs<MB><<<grid, 256, 0, h->stream>>>(S, e.B, e.out_umax, umax_bs_dev))); h->launches++; }
    CU(h, cudaGetLastError());
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}

// FP64 FMA peak of the device, measured: 16 independent DFMA chains per thread, all SMs full
// (the denominator of the factorisation's FP64 fraction; SURVEY.md section 8d asks for a measured one)
__global__ void __launch_bounds__(256) k_fp64_peak(double *out, int iters, double a, double b)
{
    double v[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = a * (threadIdx.x + q);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = fma(v[q], b, a);
    }
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 16; ++q) s += v[q];
    if (s == 123.456) out[0] = s;        // never true: keeps the chains alive
}

extern "C" int32_t kb2_measure_fp64_peak(kb2_handle h, double *tflops)
{
    if (!h || !tflops) return 1;
    if (h->device < 0) FAIL(h, "host-only handle: no CUDA device");
    CU(h, cudaSetDevice(h->device));
    double *d = nullptr;
    CU(h, cudaMalloc((void **)&d, 8));
    const int iters = 4096, grid = h->sm_count * 8;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CU(h, cudaEventRecord(h->ev0, h->stream));
        k_fp64_peak<<<grid, 256, 0, h->stream>>>(d, iters, 1e-3, 0.999);
        h->launches++;
        CU(h, cudaEventRecord(h->ev1, h->stream));
        CU(h, cudaEventSynchronize(h->ev1));
        float ms = 0;
        CU(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        if (rep > 0) best = std::max(best, 2.0 * 16.0 * iters * 256.0 * grid / (ms * 1e-3) / 1e12);
    }
    cudaFree(d);
    CU(h, cudaGetLastError());
    *tflops = best;
    return 0;
}

extern "C" int32_t kb2_get_launch_info(kb2_handle h, int32_t *members_per_tile, int32_t *ctas_per_sm)
{
    if (!h) return 1;
    if (members_per_tile) *members_per_tile = h->ens_mb;
    if (ctas_per_sm) *ctas_per_sm = h->last_ctas_per_sm;
    return 0;
}

// ---- multi-GPU: members are independent, so ranks never talk during the solve; the only
// exchange is one all-gather of the packed results (final concentrations and per-species maxima,
// member-major) over NCCL / NVLink at the end.  NCCL is bound at run time
// (dlopen of libnccl.so.2: the copy a host process has already loaded, e.g. PyTorch's, or the
// system one), so the library itself has no link-time dependency on it. ----
namespace {
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
};
NcclApi g_nccl;

bool nccl_load()
{
    if (g_nccl.lib) return true;
    const char *names[] = {getenv("KB2_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        if (!n || !*n) continue;
        g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) { g_nccl.err = "libnccl.so.2 not found (set KB2_NCCL_LIB)"; return false; }
    bool ok = true;
    auto sym = [&](const char *n) { void *p = dlsym(g_nccl.lib, n); if (!p) ok = false; return p; };
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))sym("ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))sym("ncclCommInitRank");
    g_nccl.CommInitAll = (decltype(g_nccl.CommInitAll))sym("ncclCommInitAll");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))sym("ncclCommDestroy");
    g_nccl.AllGather = (decltype(g_nccl.AllGather))sym("ncclAllGather");
    g_nccl.GroupStart = (decltype(g_nccl.GroupStart))sym("ncclGroupStart");
    g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))sym("ncclGroupEnd");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))sym("ncclGetErrorString");
    if (!ok) { g_nccl.err = "libnccl is missing a required symbol"; dlclose(g_nccl.lib); g_nccl.lib = nullptr; }
    return ok;
}
}  // namespace

#define NC(h, call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) { \
    (h)->err = std::string(#call) + ": " + g_nccl.GetErrorString(r_); return 4; } } while (0)

static void kb2_comm_release(kb2_ctx *h)
{
    if (h->comm && g_nccl.lib) g_nccl.CommDestroy(h->comm);
    h->comm = nullptr; h->comm_nranks = 1; h->comm_rank = 0;
}

extern "C" int32_t kb2_comm_unique_id(uint8_t *id128)
{
    if (!id128 || !nccl_load()) return 4;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return 4;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id128, &id, 128);
    return 0;
}

extern "C" int32_t kb2_comm_init_rank(kb2_handle h, int32_t nranks, int32_t rank, const uint8_t *id128)
{
    if (!h || !id128) return 1;
    if (h->device < 0) FAIL(h, "host-only handle: no CUDA device");
    if (nranks < 1 || rank < 0 || rank >= nranks) FAIL(h, "bad rank / world size");
    if (!nccl_load()) FAIL(h, g_nccl.err);
    kb2_comm_release(h);
    CU(h, cudaSetDevice(h->device));
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    NC(h, g_nccl.CommInitRank(&h->comm, nranks, id, rank));
    h->comm_nranks = nranks; h->comm_rank = rank;
    return 0;
}

extern "C" int32_t kb2_comm_init_all(int32_t ndev, kb2_handle *handles)
{
    if (ndev < 1 || !handles) return 1;
    kb2_ctx *h0 = handles[0];
    if (!h0) return 1;
    if (!nccl_load()) FAIL(h0, g_nccl.err);
    std::vector<int> devs(ndev);
    std::vector<ncclComm_t> comms(ndev);
    for (int i = 0; i < ndev; ++i) {
        if (!handles[i] || handles[i]->device < 0) FAIL(h0, "every handle needs a CUDA device");
        kb2_comm_release(handles[i]);
        devs[i] = handles[i]->device;
    }
    NC(h0, g_nccl.CommInitAll(comms.data(), ndev, devs.data()));
    for (int i = 0; i < ndev; ++i) { handles[i]->comm = comms[i]; handles[i]->comm_nranks = ndev; handles[i]->comm_rank = i; }
    return 0;
}

// All-gather of the packed results of the last solve over the handles' communicator.  `n` local
// handles (1 in a process-per-GPU job, all of them in a single-process job: the calls are grouped).
// Every rank must hold the same B.  final_all[i] / umax_all[i]: host buffers of nranks*B*S doubles
// (member-major, rank-major: [rank][b][i]) or NULL to leave the result on the device
// (kb2_gathered_device).
extern "C" int32_t kb2_allgather_results(kb2_handle *handles, int32_t n, double **final_all, double **umax_all)
{
    if (!handles || n < 1 || !handles[0]) return 1;
    kb2_ctx *h0 = handles[0];
    if (!nccl_load()) FAIL(h0, g_nccl.err);
    for (int i = 0; i < n; ++i) {
        kb2_ctx *h = handles[i];
        if (!h || !h->comm) FAIL(h0, "kb2_comm_init_rank / kb2_comm_init_all first");
        if (h->ens_B <= 0) FAIL(h0, "nothing to gather: run a solve first");
        CU(h, cudaSetDevice(h->device));
        const size_t per = (size_t)h->ens_B * h->net.S;
        if (h->g_send_cap < 2 * per) {
            cudaFree(h->g_send); h->g_send = nullptr; h->g_send_cap = 0;
            CU(h, cudaMalloc((void **)&h->g_send, 2 * per * 8));
            h->g_send_cap = 2 * per;
        }
        if (h->g_recv_cap < 2 * per * h->comm_nranks) {
            cudaFree(h->g_recv); h->g_recv = nullptr; h->g_recv_cap = 0;
            CU(h, cudaMalloc((void **)&h->g_recv, 2 * per * h->comm_nranks * 8));
            h->g_recv_cap = 2 * per * h->comm_nranks;
        }
        DevEns &e = h->de;
        const int S = (int)h->net.S, grid = conv_grid(h, per);
        CU(h, cudaEventRecord(h->ev0, h->stream));
        DISPATCH_MB(e.MB, (k_pack_bs<MB><<<grid, 256, 0, h->stream>>>(S, e.B, e.u, h->g_send)));
        DISPATCH_MB(e.MB, (k_pack_bs<MB><<<grid, 256, 0, h->stream>>>(S, e.B, e.out_umax, h->g_send + per)));
        h->launches += 2;
        CU(h, cudaGetLastError());
    }
    NC(h0, g_nccl.GroupStart());
    for (int i = 0; i < n; ++i) {
        kb2_ctx *h = handles[i];
        const size_t per = (size_t)h->ens_B * h->net.S;
        cudaSetDevice(h->device);
        NC(h, g_nccl.AllGather(h->g_send, h->g_recv, per, ncclDouble, h->comm, h->stream));
        NC(h, g_nccl.AllGather(h->g_send + per, h->g_recv + per * h->comm_nranks, per, ncclDouble, h->comm, h->stream));
    }
    NC(h0, g_nccl.GroupEnd());
    for (int i = 0; i < n; ++i) {
        kb2_ctx *h = handles[i];
        const size_t tot = (size_t)h->ens_B * h->net.S * h->comm_nranks;
        CU(h, cudaSetDevice(h->device));
        CU(h, cudaEventRecord(h->ev1, h->stream));
        if (final_all && final_all[i]) CU(h, cudaMemcpyAsync(final_all[i], h->g_recv, tot * 8, cudaMemcpyDeviceToHost, h->stream));
        if (umax_all && umax_all[i]) CU(h, cudaMemcpyAsync(umax_all[i], h->g_recv + tot, tot * 8, cudaMemcpyDeviceToHost, h->stream));
    }
    for (int i = 0; i < n; ++i) {
        kb2_ctx *h = handles[i];
        CU(h, cudaSetDevice(h->device));
        CU(h, cudaStreamSynchronize(h->stream));
        CU(h, cudaEventElapsedTime(&h->gather_ms, h->ev0, h->ev1));
    }
    return 0;
}

// device pointers of the gathered results ([nranks][B][S] each), device time of the last gather
// (pack kernels + the two all-gathers), rank and world size of the handle's communicator
extern "C" int32_t kb2_gathered_device(kb2_handle h, double **final_all_dev, double **umax_all_dev, float *gather_ms,
                                       int32_t *rank, int32_t *nranks)
{
    if (!h) return 1;
    const size_t tot = (size_t)std::max<int64_t>(h->ens_B, 0) * h->net.S * h->comm_nranks;
    if (final_all_dev) *final_all_dev = h->g_recv;
    if (umax_all_dev) *umax_all_dev = h->g_recv ? h->g_recv + tot : nullptr;
    if (gather_ms) *gather_ms = h->gather_ms;
    if (rank) *rank = h->comm_rank;
    if (nranks) *nranks = h->comm_nranks;
    return 0;
}
