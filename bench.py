#!/usr/bin/env python
"""bench.py — ensemble CRN solves/sec on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c3|c4|c5]

One "step" = one pass of the hot path over one batch: the full stiff solve (t0 -> tf, discrete
rate updates at every tstop, saves) of B ensemble members of the named synthetic CRN.  At N > 1
(launched under torch.distributed.run, one rank per GPU) every rank solves its own B members
(weak scaling), then the final concentrations and per-species maxima are all-gathered over NCCL.

The JSON line carries `value` (device-timed, inputs resident in HBM), `e2e` (same metric through
the host-buffer API: H2D of inputs + solve + D2H of results inside the timed region), `roofline`
for the dominant kernel (the fused solve kernel), per-kernel roofline numbers for the stand-alone
kernels, and `cpu_baseline` (the plain-C oracle on the host cores; the Julia reference cannot be
run: no julia binary in this image).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

WORKLOADS = {
    # name: (S, R, members per GPU, config id for the seed, description)
    "c3": (1000, 5000, 4096, 3, "C3: synthetic 1k-species/5k-reaction stiff mass-action CRN, 4096-member temperature-ramp ensemble per GPU"),
    "c4": (10000, 50000, 1024, 4, "C4: synthetic 10k-species/50k-reaction CRN, 1024-member ensemble per GPU"),
    "c5": (5000, 25000, 8192, 5, "C5: synthetic 5k-species/25k-reaction CRN, 65536-member sweep sharded as 8192 members per GPU"),
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


TOLS = {"default": (1e-10, 1e-8), "throughput": (1e-8, 1e-6)}


def build_problem(name, B, rank=0, world=1, tol="default"):
    import kinetica_b200 as kb
    from kinetica_b200.synthetic import synthetic_crn, synthetic_u0, SEED_BASE
    S, R, _, cid, _ = WORKLOADS[name]
    sd, rd, Ea, A = synthetic_crn(S, R, SEED_BASE + cid)
    calc = kb.PrecalculatedArrheniusCalculator(Ea, A, k_max=1e12)
    pars = kb.ODESimulationParams(tspan=(0.0, 1.0), u0=synthetic_u0(S), save_interval=0.1,
                                  low_k_cutoff="none", solve_chunks=False, abstol=TOLS[tol][0], reltol=TOLS[tol][1])
    Btot = B * world
    # member b of the whole job: ramp from 600 + 600*b/(Btot-1) K, +100 K at 100 K/s, ts_update 1e-2
    conds = []
    for b in range(rank * B, (rank + 1) * B):
        T0 = 600.0 + 600.0 * b / max(Btot - 1, 1)
        conds.append(kb.ConditionSet({"T": kb.LinearDirectProfile(rate=100.0, X_start=T0, X_end=T0 + 100.0)},
                                     ts_update=1e-2))
    for cs in conds:
        cs.solve_variable_conditions(pars)
    return sd, rd, Ea, A, calc, pars, conds


def algorithmic_bytes_per_attempt(S, R, nnzJ, nnzLU):
    """SURVEY.md §8(d), per member per attempted Rodas4 step (6 stages), FP64."""
    jac = 8 * (R + S + nnzJ)
    lu = 8 * (nnzJ + nnzLU)
    rhs = 8 * (R + 2 * S)
    tri = 8 * (nnzLU + 2 * S)
    combine = 8 * S * (2 + 3 + 4 + 5 + 6 + 7 + 4)      # stage arguments + error norm/commit
    return jac + lu + 6 * (rhs + tri) + combine


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


def cpu_reference(name, n_members, steps, warmup, nthreads, tol="default"):
    """The CPU arm: plain-C oracle (same algorithm, OpenMP over members) on a bounded sample of the
    workload: `n_members` members spread evenly over the temperature sweep."""
    from oracle import c_oracle as co, kinetica_oracle as ko
    S, R, B, cid, desc = WORKLOADS[name]
    sd, rd, Ea, A, calc, pars, conds = build_problem(name, n_members, tol=tol)
    net = ko.Network(S, rd.id_reacs, rd.id_prods, rd.stoic_reacs, rd.stoic_prods)
    colptr, rowval = net.pattern_csc()
    perm = ko.min_degree_order(S, colptr, rowval)
    rowptr, colidx, diagpos, _ = ko.symbolic_lu(S, colptr, rowval, perm)
    Ts = [cs.get_profile("T").X_start for cs in conds]
    ts = conds[0].get_tstops()
    save = np.arange(11) / 10.0
    u0 = np.zeros(S); u0[8:18] = 0.1
    times = []
    for it in range(warmup + steps):
        t = time.perf_counter()
        out, st, stats, _ = co.solve_rodas4(net, A, Ea, 1e12, 1.0, Ts, ts, lambda b, tt: Ts[b] + 100.0 * min(tt, 1.0),
                                            u0, (0.0, 1.0), save, nthreads=nthreads, abstol=pars.abstol,
                                            reltol=pars.reltol, symbolic=(perm, rowptr, colidx, diagpos))
        dt = time.perf_counter() - t
        if it >= warmup:
            times.append(dt)
        assert np.all(st == 0)
    sec = float(np.mean(times))
    return n_members / sec, sec, int(stats[:, 2].mean())


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
        if rank != 0:
            return
        sps, sec, attempts = cpu_reference(args.workload, args.cpu_sample, max(args.steps, 1), min(args.warmup, 1), ncores, args.tol)
        sample = (f"{args.cpu_sample} members evenly spaced over the {B}-member sweep, {used} of {ncores} host threads busy "
                  f"(OpenMP over members), full t0->tf solve each")
        print(json.dumps({
            "impl": "reference", "metric": "ensemble_crn_solves_per_sec", "value": sps, "unit": "solves/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "members_per_gpu": B, "note": "CPU restatement of reference semantics "
                       "(plain-C oracle, Rodas4 + sparse LU); the Julia reference cannot run here (no julia binary)"},
            "cpu_baseline": {"value": sps, "unit": "solves/s", "cores": used, "kind": "port", "sample": sample},
            "e2e": {"value": sps, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    import torch
    import kinetica_b200 as kb
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
    Ns = 11
    # pinned host buffers for the e2e path
    out_u = torch.empty((Ns, S, B), dtype=torch.float64).pin_memory().numpy()
    out_umax = torch.empty((S, B), dtype=torch.float64).pin_memory().numpy()
    fin_loc = torch.empty((B, S), dtype=torch.float64, device=f"cuda:{dev}")
    max_loc = torch.empty((B, S), dtype=torch.float64, device=f"cuda:{dev}")
    fin_all = torch.empty((B * world, S), dtype=torch.float64, device=f"cuda:{dev}") if world > 1 else None
    max_all = torch.empty((B * world, S), dtype=torch.float64, device=f"cuda:{dev}") if world > 1 else None

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def gather():
        if dist is None:
            return
        es.h.pack_results_device(fin_loc.data_ptr(), max_loc.data_ptr())
        dist.all_gather_into_tensor(fin_all, fin_loc)
        dist.all_gather_into_tensor(max_all, max_loc)
        torch.cuda.synchronize()

    # ---- K timed steps.  Each step goes through the host-buffer API (H2D of inputs, solve, D2H of
    # results): its wall time is the e2e arm, the CUDA-event time of the solve kernel inside it
    # (inputs already resident when that region starts) is the device arm. ----
    sampler = ClockSampler(dev)
    launches0 = es.h.launch_count
    dev_ms, e2e_t, gather_t = [], [], []
    for it in range(args.warmup + args.steps):
        if it == args.warmup:
            sampler.start()
            launches0 = es.h.launch_count
        barrier()
        t = time.perf_counter()
        es.prepare(conds, pars, u0)              # H2D: u0, profiles, stop tables
        ms = es.run()                            # CUDA events around the solve kernel, on its stream
        _, _, status, stats = es.h.solve_fetch(out_u, out_umax)   # D2H: saves, maxima, status, stats
        tg = time.perf_counter()
        gather()
        tg = time.perf_counter() - tg
        barrier()
        if it >= args.warmup:
            dev_ms.append(ms)
            e2e_t.append(time.perf_counter() - t)
            gather_t.append(tg)
    clocks = sampler.stop()
    launches = es.h.launch_count - launches0
    ok = int(np.sum(status == 0))
    attempts = int(stats[:, 2].sum())
    step_s = float(np.mean(dev_ms)) * 1e-3 + (float(np.mean(gather_t)) if dist is not None else 0.0)
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

    # ---- roofline of the dominant kernel + per-kernel numbers ----
    peak, peak_src = load_peaks()
    alg_bytes = algorithmic_bytes_per_attempt(S, R, es.nnzJ, es.nnzLU) * attempts
    solve_ms = float(np.mean(dev_ms))
    achieved = alg_bytes / (solve_ms * 1e-3) / 1e9
    # measured DRAM traffic of k_solve: dram__bytes_read.sum + dram__bytes_write.sum of an ncu capture
    # of the same kernel on the same network (C3, 4096 members, 8 attempted steps per member:
    # profiles/r01_k_solve_8attempts_final_dram.csv for this build; the `--set full` capture one
    # commit earlier, profiles/r01_k_solve_8attempts_aligned_full.txt, has 486.7 + 86.4 GB), per
    # member and attempted step, scaled to this launch's attempts; only quoted for the workload
    # it was captured on
    traffic = None
    if args.workload == "c3" and B == 4096:
        traffic = (465.597729e9 + 86.312270e9) / (4096 * 8) * attempts
    kern = {}
    try:
        per = {"arrhenius": 8 * (R + 1), "rhs": 8 * (R + 2 * S), "jacobian": 8 * (R + S + es.nnzJ),
               "w_assembly+lu": 8 * (R + S + es.nnzJ + 2 * es.nnzLU), "trisolve": 8 * (es.nnzLU + 2 * S)}
        for w, nm in enumerate(["arrhenius", "rhs", "jacobian", "w_assembly+lu", "trisolve"]):
            ms = es.h.time_kernel(w, B, 5)
            gbs = per[nm] * B / (ms * 1e-3) / 1e9
            kern[nm] = {"ms": ms, "alg_GBps": gbs, "frac_hbm": gbs / peak}
        t_rj = kern["rhs"]["ms"] + kern["jacobian"]["ms"]
        kern["rhs+jacobian"] = {"ms": t_rj, "alg_GBps": B * 8 * (2 * R + 3 * S + es.nnzJ) / (t_rj * 1e-3) / 1e9}
        kern["rhs+jacobian"]["frac_hbm"] = kern["rhs+jacobian"]["alg_GBps"] / peak
        kern["w_assembly+lu"]["fp64_tflops"] = 2 * es.n_fma * B / (kern["w_assembly+lu"]["ms"] * 1e-3) / 1e12
    except Exception as e:      # pragma: no cover
        kern["error"] = str(e)

    line = {
        "metric": "ensemble_crn_solves_per_sec", "value": value, "unit": "solves/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "S": S, "R": R, "members_per_gpu": B, "tspan": [0.0, 1.0], "ts_update": 1e-2,
                   "saves": Ns, "abstol": pars.abstol, "reltol": pars.reltol, "integrator": "Rodas4",
                   "nnzJ": es.nnzJ, "nnzLU": es.nnzLU, "lu_fma_per_member": es.n_fma,
                   "l2": "inputs larger than L2 (LU values %.1f GB per launch)" % (8 * es.nnzLU * Bp / 1e9),
                   "symbolic_s": t_sym, "members_ok": ok_all, "attempted_steps_per_member": attempts / B},
        "clocks": clocks,
        "e2e": {"value": ok_all / e2e_s, "unit": "solves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s * 1e3},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "k_solve (fused Rodas4 tile kernel)", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes},
        "kernels": kern,
    }
    if rank == 0 and world == 1 and not args.no_cpu:
        sps, sec, _ = cpu_reference(args.workload, args.cpu_sample, 1, 0, ncores, args.tol)
        line["cpu_baseline"] = {"value": sps, "unit": "solves/s", "cores": used, "kind": "port",
                                "sample": f"{args.cpu_sample} members evenly spaced over the sweep, {used} of {ncores} host "
                                          f"threads busy, {sec:.1f} s; plain-C oracle (same Rodas4 + sparse LU), not the Julia reference"}
    if rank == 0:
        print(json.dumps(line))
    es.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
