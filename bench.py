#!/usr/bin/env python
"""bench.py — ensemble CRN solves/sec on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c3|c4|c5]
                    [--budget-s SECONDS]

One "step" = one pass of the hot path over one batch: the full stiff solve (t0 -> tf, discrete
rate updates at every tstop, saves) of B ensemble members of the named synthetic CRN.  At N > 1
(launched under torch.distributed.run, one rank per GPU) every rank solves its own B members
(weak scaling, no communication during the solve), then the final concentrations and per-species
maxima are all-gathered over NCCL by libkinetica_b200.so itself (kb2_allgather_results).

A full C3 step costs tens of seconds, so `--budget-s` (default 780 s wall clock for the whole
process) bounds the run: if `(warmup + steps) x step time` does not fit, fewer repetitions are
executed and the JSON line reports the EXECUTED `steps` / `warmup` (the requested ones are kept
under `config.requested`).

The JSON line carries `value` (device-timed, inputs resident in HBM), `e2e` (same metric through
the host-buffer API: H2D of inputs + solve + D2H of results inside the timed region), `roofline`
for the dominant phase kernel (average launch duration measured live with CUDA events around its
launches inside the timed solves), per-kernel roofline numbers for all phase kernels, and
`cpu_baseline` (the plain-C oracle on the host cores; the Julia reference cannot be run: no julia
binary in this image).
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

T_START = time.perf_counter()
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

WORKLOADS = {
    # name: (S, R, members per GPU, config id for the seed, description)
    "c3": (1000, 5000, 4096, 3, "C3: synthetic 1k-species/5k-reaction stiff mass-action CRN, 4096-member temperature-ramp ensemble per GPU"),
    "c4": (10000, 50000, 1024, 4, "C4: synthetic 10k-species/50k-reaction CRN, 1024-member ensemble per GPU"),
    "c5": (5000, 25000, 8192, 5, "C5: synthetic 5k-species/25k-reaction CRN, 65536-member sweep sharded as 8192 members per GPU"),
}
PARITY_RTOL = 1e-4      # GPU vs plain-C twin at the reference-default tolerances (the solver's own global error level)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


KERNEL_FILES = ("kb2_kernels.cuh", "kb2_solve.cuh", "kb2_front.cuh")
# what the DRAM traffic of a phase kernel depends on: the kernel sources it is compiled from and, for the
# phases that stream the factors, the block plan (padded panel storage) of the ordering in use
TRAFFIC_DEPENDS = {
    "jacobian": (("kb2_kernels.cuh", "kb2_solve.cuh"), False),
    "stage_rhs": (("kb2_kernels.cuh", "kb2_solve.cuh"), False),
    "step_end": (("kb2_kernels.cuh", "kb2_solve.cuh"), False),
    "stage_sweeps": (("kb2_kernels.cuh", "kb2_solve.cuh"), True),
    "lu": (KERNEL_FILES, True),
}


def code_hash(text):
    """Hash of a kernel source: code only (comments and blank lines do not count)."""
    h = hashlib.sha256()
    for line in text.splitlines():
        code = line.split("//")[0].strip()
        if code:
            h.update(code.encode() + b"\n")
    return h.hexdigest()[:16]


def kernel_source_hashes():
    return {f: code_hash(open(os.path.join(ROOT, "kinetica.jl_b200", "csrc", f), "r").read()) for f in KERNEL_FILES}


def load_traffic(workload, B, padded):
    """DRAM bytes per launch of the phase kernels from an ncu capture (profiles/r02_traffic.json,
    written by scripts/ncu_traffic.py).  A phase's figure is only used while the kernel sources it was
    captured on are unchanged and, for the factorisation and the sweeps, the plan has the same padded
    storage (a different ordering moves different bytes); otherwise it is null."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if not os.path.exists(p):
        return {}
    d = json.load(open(p))
    if d.get("workload") != workload or d.get("members") != B:
        return {}
    now, then = kernel_source_hashes(), d.get("source_hashes", {})
    out = {}
    for ph, v in d.get("dram_bytes_per_launch", {}).items():
        files, plan = TRAFFIC_DEPENDS.get(ph, (KERNEL_FILES, True))
        if all(now[f] == then.get(f) for f in files) and (not plan or d.get("plan", {}).get("padded") == padded):
            out[ph] = v
    return out


TOLS = {"default": (1e-10, 1e-8), "throughput": (1e-8, 1e-6)}


def build_problem(name, B, rank=0, world=1, tol="default", members=None):
    """members: explicit global member indices (CPU sample / parity spot check) instead of a rank's slice."""
    import kinetica_b200 as kb
    from kinetica_b200.synthetic import synthetic_crn, synthetic_u0, SEED_BASE
    S, R, Bw, cid, _ = WORKLOADS[name]
    sd, rd, Ea, A = synthetic_crn(S, R, SEED_BASE + cid)
    calc = kb.PrecalculatedArrheniusCalculator(Ea, A, k_max=1e12)
    pars = kb.ODESimulationParams(tspan=(0.0, 1.0), u0=synthetic_u0(S), save_interval=0.1,
                                  low_k_cutoff="none", solve_chunks=False, abstol=TOLS[tol][0], reltol=TOLS[tol][1])
    Btot = B * world
    # a rank's members are strided over the global sweep (b = rank, rank + world, ...): every rank gets
    # the same mix of temperatures, hence about the same number of steps (load balance, parallel.py)
    idx = list(members) if members is not None else list(range(rank, Btot, world))
    # member b of the whole job: ramp from 600 + 600*b/(Btot-1) K, +100 K at 100 K/s, ts_update 1e-2
    conds = []
    for b in idx:
        T0 = 600.0 + 600.0 * b / max(Btot - 1, 1)
        conds.append(kb.ConditionSet({"T": kb.LinearDirectProfile(rate=100.0, X_start=T0, X_end=T0 + 100.0)},
                                     ts_update=1e-2))
    for cs in conds:
        cs.solve_variable_conditions(pars)
    return sd, rd, Ea, A, calc, pars, conds


def phase_bytes(S, R, nnzJ, nnzLU):
    """Algorithmic bytes per member and launch of each phase kernel."""
    return {
        "jacobian": 8 * (R + S + nnzJ),
        "lu": 8 * (nnzJ + nnzLU),
        "stage_rhs": 8 * (R + 2 * S) + 8 * S * 25 / 6.0,      # + stage argument / combination vectors (avg over the 6 stages)
        "stage_sweeps": 8 * (nnzLU + 2 * S),
        "step_end": 8 * S * 4,
    }


ORDERING_NAMES = {0: "minimum degree", 1: "natural", 2: "caller-supplied", 3: "natural, hub species last",
                  5: "reverse Cuthill-McKee, hub species last", 6: "Sloan 1:2, hub species last", 7: "Sloan 2:1, hub species last"}
PHASE_LAUNCHES = {"jacobian": 1, "lu": 1, "stage_rhs": 6, "stage_sweeps": 6, "step_end": 1}


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in self.rows)]
        busy = [x for x in sm if x > 0.5 * (max(mx) if mx else 1)] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_solve(name, members, Btot, nthreads, tol="default"):
    """Plain-C oracle (same Rodas4 + sparse LU, OpenMP over members) on the given global members of
    the workload's sweep -> (trajectories [n, Ns, S], seconds, mean attempted steps)."""
    from oracle import c_oracle as co, kinetica_oracle as ko
    S, R, B, cid, desc = WORKLOADS[name]
    sd, rd, Ea, A, calc, pars, conds = build_problem(name, Btot, tol=tol, members=members)
    net = ko.Network(S, rd.id_reacs, rd.id_prods, rd.stoic_reacs, rd.stoic_prods)
    colptr, rowval = net.pattern_csc()
    perm = ko.min_degree_order(S, colptr, rowval)
    rowptr, colidx, diagpos, _ = ko.symbolic_lu(S, colptr, rowval, perm)
    Ts = [cs.get_profile("T").X_start for cs in conds]
    ts = conds[0].get_tstops()
    save = np.arange(11) / 10.0
    u0 = np.zeros(S); u0[8:18] = 0.1
    t = time.perf_counter()
    out, st, stats, _ = co.solve_rodas4(net, A, Ea, 1e12, 1.0, Ts, ts, lambda b, tt: Ts[b] + 100.0 * min(tt, 1.0),
                                        u0, (0.0, 1.0), save, nthreads=nthreads, abstol=pars.abstol,
                                        reltol=pars.reltol, symbolic=(perm, rowptr, colidx, diagpos))
    dt = time.perf_counter() - t
    assert np.all(st == 0)
    return out, dt, float(stats[:, 2].mean())


def sample_members(n, Btot):
    return [int(round(i * (Btot - 1) / max(n - 1, 1))) for i in range(n)]


def reference_arm(args, B, desc, ncores):
    """`--impl reference`: the CPU arm — the plain-C port of the same algorithm on all host threads,
    one member per thread, each step a bounded sample of the workload."""
    n = args.cpu_sample
    used = min(ncores, n)
    warm = min(args.warmup, 1)              # one warm-up pass is enough for a CPU loop; reported as executed
    times = []
    budget_left = lambda: args.budget_s - (time.perf_counter() - T_START)
    steps_done = 0
    for it in range(warm + max(args.steps, 1)):
        _, dt, attempts = cpu_solve(args.workload, sample_members(n, B), B, ncores, args.tol)
        if it >= warm:
            times.append(dt)
            steps_done += 1
        if budget_left() < 1.5 * dt and steps_done >= 1:
            break
    sec = float(np.mean(times))
    sps = n / sec
    sample = (f"{n} members evenly spaced over the {B}-member sweep, {used} of {ncores} host threads busy "
              f"(OpenMP over members), full t0->tf solve each")
    print(json.dumps({
        "impl": "reference", "metric": "ensemble_crn_solves_per_sec", "value": sps, "unit": "solves/s",
        "n_gpus": args.gpus, "steps": steps_done, "warmup": warm, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "S": WORKLOADS[args.workload][0], "R": WORKLOADS[args.workload][1],
                   "members_per_gpu": B, "tspan": [0.0, 1.0], "ts_update": 1e-2, "saves": 11,
                   "abstol": TOLS[args.tol][0], "reltol": TOLS[args.tol][1], "integrator": "Rodas4",
                   "requested": {"steps": args.steps, "warmup": args.warmup},
                   "note": "vs own CPU port: CPU restatement of reference semantics (plain-C oracle, Rodas4 + sparse "
                           "LU, same tableau / controller as the CUDA path); the Julia reference cannot run here "
                           "(no julia binary), so this is NOT Kinetica.jl with CVODE_BDF/KLU"},
        "cpu_baseline": {"value": sps, "unit": "solves/s", "cores": used, "kind": "port", "sample": sample,
                         "per_core": sps / used},
        "e2e": {"value": sps, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--members", type=int, default=0, help="override members per GPU (parity/debug only)")
    ap.add_argument("--tol", default="default", choices=sorted(TOLS),
                    help="default = the reference's abstol 1e-10 / reltol 1e-8; throughput = 1e-8 / 1e-6")
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="members of the CPU sample (default: one per host thread, at least 8)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--budget-s", type=float, default=780.0,
                    help="wall-clock budget of the whole process; repetitions are cut to fit and reported as executed")
    ap.add_argument("--parity-members", type=int, default=4, help="members spot-checked against the plain-C twin (N=1, outside the timed region)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    S, R, B, cid, desc = WORKLOADS[args.workload]
    if args.members:
        B = args.members
    ncores = os.cpu_count() or 1
    if args.cpu_sample <= 0:
        args.cpu_sample = max(8, ncores)       # OpenMP over members: one member per host thread keeps every core busy
    used = min(ncores, args.cpu_sample)        # threads that actually had work

    if args.impl == "reference":
        if rank == 0:
            reference_arm(args, B, desc, ncores)
        return

    import torch
    import kinetica_b200 as kb
    from kinetica_b200 import parallel
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank if world > 1 else 0
    torch.cuda.set_device(dev)

    sd, rd, Ea, A, calc, pars, conds = build_problem(args.workload, B, rank, world, args.tol)
    u0 = pars.u0
    t_sym = time.perf_counter()
    es = kb.EnsembleSolver(sd, rd, calc, device=dev)
    t_sym = time.perf_counter() - t_sym
    if dist is not None:
        parallel.init_comm(es.h, rank, world)          # NCCL communicator inside libkinetica_b200.so
    Ns = 11
    # pinned host buffers for the e2e path
    out_u = torch.empty((Ns, S, B), dtype=torch.float64).pin_memory().numpy()
    out_umax = torch.empty((S, B), dtype=torch.float64).pin_memory().numpy()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def one_step():
        """host-buffer API: H2D of inputs, solve, D2H of results, (N > 1) all-gather of the summaries"""
        barrier()
        t = time.perf_counter()
        es.prepare(conds, pars, u0)              # H2D: u0, profiles, stop tables
        ms = es.run()                            # CUDA events from the first to the last phase kernel, on its stream
        _, _, status, stats = es.h.solve_fetch(out_u, out_umax)   # D2H: saves, maxima, status, stats
        g_ms = 0.0
        if dist is not None:
            barrier()                            # ranks finish their solves at different times: keep that wait out of the gather's own time
            es.h.allgather_results(to_host=False)
            g_ms = es.h.gathered_device()["gather_ms"]
        barrier()
        return ms, g_ms, time.perf_counter() - t, status, stats

    # ---- how many repetitions fit the wall-clock budget (decided on rank 0's clock for all ranks) ----
    ms, g_ms, e2e, status, stats = one_step()          # first warm-up step doubles as the probe
    t_step = e2e
    reserve = 45.0 + (25.0 if (world == 1 and not args.no_cpu) else 0.0)      # kernel timing, parity spot check, CPU leg
    left = args.budget_s - (time.perf_counter() - T_START) - reserve
    n_fit = max(1, int(left / (t_step * 1.03)))
    want_w, want_k = max(args.warmup - 1, 0), max(args.steps, 1)
    if want_w + want_k <= n_fit:
        ex_w, ex_k = want_w, want_k
    else:
        ex_w = min(want_w, 2, max(n_fit - 1, 0))          # keep 3 warm-up steps (probe included) when there is room
        ex_k = max(1, min(want_k, n_fit - ex_w))
    if dist is not None:
        tt = torch.tensor([ex_w, ex_k], device=f"cuda:{dev}", dtype=torch.int64)
        dist.broadcast(tt, src=0)
        ex_w, ex_k = int(tt[0].item()), int(tt[1].item())
    for _ in range(ex_w):
        one_step()

    # ---- K timed steps.  Each step goes through the host-buffer API: its wall time is the e2e arm,
    # the CUDA-event time of the solve inside it (inputs already resident when that region starts)
    # plus the device time of the gather is the device arm. ----
    sampler = ClockSampler(dev)
    sampler.start()
    launches0 = es.h.launch_count
    dev_ms, e2e_t, gather_ms = [], [], []
    phase_acc = {}
    rounds = 0
    for it in range(ex_k):
        ms, g_ms, e2e, status, stats = one_step()
        dev_ms.append(ms); e2e_t.append(e2e); gather_ms.append(g_ms)
        ph, rounds = es.h.get_phase_times()
        for k, v in ph.items():
            a = phase_acc.setdefault(k, [0.0, 0])
            a[0] += v["ms"] * v["sampled_launches"]; a[1] += v["sampled_launches"]
    clocks = sampler.stop()
    launches = es.h.launch_count - launches0
    ok = int(np.sum(status == 0))
    attempts = int(stats[:, 2].sum())
    step_s = (float(np.mean(dev_ms)) + float(np.mean(gather_ms))) * 1e-3
    e2e_s = float(np.mean(e2e_t))
    if dist is not None:
        tt = torch.tensor([step_s, e2e_s], device=f"cuda:{dev}", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        step_s, e2e_s = float(tt[0].item()), float(tt[1].item())
        oks = torch.tensor([ok], device=f"cuda:{dev}", dtype=torch.int64)
        dist.all_reduce(oks)
        ok_all = int(oks.item())
    else:
        ok_all = ok
    value = ok_all / step_s
    Bp = (B + 31) // 32 * 32
    h2d = 8 * S * Bp + Bp * (4 + 16 * 8) + Bp * 102 * 16
    d2h = 8 * Ns * S * B + 8 * S * B + B * 4 + B * 64

    # ---- per-kernel roofline, from the live phase timing of the timed solves ----
    peak, peak_src = load_peaks()
    pb = phase_bytes(S, R, es.nnzJ, es.nnzLU)
    plan = es.h.get_plan_stats()
    traffic = load_traffic(args.workload, B, plan["padded"])
    kern = {}
    round_ms = 0.0
    for nm, (tot, n) in phase_acc.items():
        if n == 0:
            continue
        ms_k = tot / n
        round_ms += ms_k * PHASE_LAUNCHES[nm]
        gbs = pb[nm] * B / (ms_k * 1e-3) / 1e9
        kern[nm] = {"ms": ms_k, "launches_per_step_attempt": PHASE_LAUNCHES[nm], "alg_bytes_per_launch": pb[nm] * B,
                    "alg_GBps": gbs, "frac_hbm": gbs / peak, "sampled_launches": n,
                    "traffic": traffic.get(nm)}
    for nm in kern:
        kern[nm]["share"] = kern[nm]["ms"] * PHASE_LAUNCHES[nm] / round_ms
    try:
        fp64_peak = es.h.measure_fp64_peak()
        if "lu" in kern:
            tf = 2 * es.n_fma * B / (kern["lu"]["ms"] * 1e-3) / 1e12
            kern["lu"].update({"fp64_tflops": tf, "fp64_peak_tflops": fp64_peak, "frac_fp64": tf / fp64_peak,
                               "fp64_peak_source": "measured (kb2_measure_fp64_peak: dependency-free DFMA chains on all SMs)"})
    except Exception as e:      # pragma: no cover
        kern["fp64_peak_error"] = str(e)
    if "stage_rhs" in kern and "jacobian" in kern:
        t_rj = kern["stage_rhs"]["ms"] + kern["jacobian"]["ms"]
        gbs = B * (pb["stage_rhs"] + pb["jacobian"]) / (t_rj * 1e-3) / 1e9
        kern["rhs+jacobian"] = {"ms": t_rj, "alg_GBps": gbs, "frac_hbm": gbs / peak}
    dom = max((k for k in kern if "share" in kern[k]), key=lambda k: kern[k]["share"])
    kd = kern[dom]
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kd["alg_GBps"], "peak": peak, "unit": "GB/s",
                "frac": kd["frac_hbm"], "traffic": kd["traffic"], "peak_source": peak_src,
                "algorithmic_bytes_per_launch": kd["alg_bytes_per_launch"], "avg_launch_ms": kd["ms"],
                "share_of_step": kd["share"],
                "whole_solve": {"achieved": sum(pb[k] * PHASE_LAUNCHES[k] for k in pb) * attempts / (float(np.mean(dev_ms)) * 1e-3) / 1e9,
                                "algorithmic_bytes_per_member_attempt": sum(pb[k] * PHASE_LAUNCHES[k] for k in pb),
                                "note": "algorithmic bytes of all phase kernels x attempted steps / device time of the solve; the bytes "
                                        "follow nnzLU of the ordering in use (SURVEY 8d: 8(nnzJ+nnzLU) for the LU, 8(nnzLU+2S) per sweep): "
                                        "an ordering with less fill lowers them, so this fraction is only comparable at equal nnzLU"}}
    roofline["whole_solve"]["frac"] = roofline["whole_solve"]["achieved"] / peak
    if dom == "lu" and "frac_fp64" in kd:
        roofline.update({"frac_fp64": kd["frac_fp64"], "fp64_tflops": kd["fp64_tflops"], "fp64_peak_tflops": kd["fp64_peak_tflops"],
                         "note": "the factorisation keeps the active submatrix in shared memory: HBM sees the compact Jacobian values in and "
                                 "the factors out once (traffic ~1.25x the algorithmic bytes); what bounds it is on-chip — the shared-memory "
                                 "bandwidth of the rank-8 update and the latency of the strip / pivot phases with 8 warps per SM (DESIGN.md section 5 K4)"})

    # ---- parity spot check against the plain-C twin, outside the timed region (N = 1) ----
    parity = None
    if world == 1 and args.parity_members > 0:
        try:
            mem = sample_members(args.parity_members, B)
            ref, sec, _ = cpu_solve(args.workload, mem, B, min(ncores, len(mem)), args.tol)
            worst = 0.0
            for q, b in enumerate(mem):
                got = out_u[:, :, b]
                worst = max(worst, float(np.max(np.abs(got - ref[q]) / (PARITY_RTOL * np.abs(ref[q]) + 1e-9))))
            parity = {"members": mem, "worst_over_bound": worst, "ok": bool(worst < 1.0)}
        except Exception as e:      # pragma: no cover
            parity = {"error": str(e)}

    line = {
        "metric": "ensemble_crn_solves_per_sec", "value": value, "unit": "solves/s", "n_gpus": world,
        "steps": ex_k, "warmup": ex_w + 1, "ms_per_step": step_s * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "S": S, "R": R, "members_per_gpu": B, "tspan": [0.0, 1.0], "ts_update": 1e-2,
                   "saves": Ns, "abstol": pars.abstol, "reltol": pars.reltol, "integrator": "Rodas4",
                   "nnzJ": es.nnzJ, "nnzLU": es.nnzLU, "lu_fma_per_member": es.n_fma, "lu_padded_slots": plan["padded"],
                   "ordering": ORDERING_NAMES.get(plan["ordering"], str(plan["ordering"])) + " (chosen by `auto`)",
                   "l2": "inputs larger than L2 (LU values %.1f GB per launch)" % (8 * es.nnzLU * Bp / 1e9),
                   "symbolic_s": t_sym, "members_ok": ok_all, "attempted_steps_per_member": attempts / B,
                   "rounds": rounds, "requested": {"steps": args.steps, "warmup": args.warmup, "budget_s": args.budget_s},
                   "parity_rtol": PARITY_RTOL, "parity_checked_members": parity,
                   "gather_ms": float(np.mean(gather_ms)), "launch": es.h.get_launch_info()},
        "clocks": clocks,
        "e2e": {"value": ok_all / e2e_s, "unit": "solves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s * 1e3},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "kernels": kern,
    }
    if rank == 0 and world == 1 and not args.no_cpu:
        n = args.cpu_sample
        _, sec, _ = cpu_solve(args.workload, sample_members(n, B), B, ncores, args.tol)
        sps = n / sec
        line["cpu_baseline"] = {"value": sps, "unit": "solves/s", "cores": used, "kind": "port", "per_core": sps / used,
                                "sample": f"{n} members evenly spaced over the sweep, {used} of {ncores} host "
                                          f"threads busy, {sec:.1f} s; plain-C oracle (same Rodas4 + sparse LU), not the Julia reference"}
    es.close()
    if dist is not None:
        dist.destroy_process_group()
    if rank == 0:
        sys.stdout.flush()
        print(json.dumps(line), flush=True)      # the last line of stdout (NCCL may print its version line earlier)


if __name__ == "__main__":
    main()
