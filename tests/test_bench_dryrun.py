"""bench.py, GPU arm, dry run on CPU: the JSON line is assembled from a stubbed solver (real symbolic
analysis on a host-only handle, invented timings), so that a slip in the reporting code shows up here
and not as a lost bench line on the GPU box.  Nothing is measured."""
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _FakeTensor:
    def __init__(self, shape):
        self.a = np.zeros(shape)

    def pin_memory(self):
        return self

    def numpy(self):
        return self.a


def _fake_torch():
    t = types.ModuleType("torch")
    t.float64 = "float64"
    t.empty = lambda shape, dtype=None: _FakeTensor(shape)
    t.cuda = types.SimpleNamespace(set_device=lambda d: None, synchronize=lambda: None)
    return t


def test_gpu_arm_line_from_a_stubbed_solver(monkeypatch, capsys):
    sys.path.insert(0, ROOT)
    import bench
    import kinetica_b200 as kb
    from kinetica_b200 import _lib

    class FakeHandle:
        def __init__(self, real):
            self.real, self.launch_count = real, 0

        def solve_fetch(self, out_u, out_umax):
            B = out_u.shape[2]
            stats = np.zeros((B, 8), dtype=np.int64); stats[:, 2] = 100
            return None, None, np.zeros(B, dtype=np.int32), stats

        def get_phase_times(self):
            return {k: {"ms": 1.0 + i, "sampled_launches": 2} for i, k in enumerate(bench.PHASE_LAUNCHES)}, 110

        def measure_fp64_peak(self):
            return 36.7

        def get_plan_stats(self):
            return self.real.get_plan_stats()

        def get_launch_info(self):
            return {"members_per_tile": 4, "ctas_per_sm": 7}

    class FakeSolver:
        def __init__(self, sd, rd, calc, device=0):
            real = _lib.Handle(-1)
            real.set_network(sd.n, *rd.flatten())
            self.nnzJ, self.nnzLU, self.n_fma = real.symbolic(4)
            self.h = FakeHandle(real)

        def prepare(self, conds, pars, u0):
            pass

        def run(self):
            self.h.launch_count += 15 * 110
            return 1500.0

        def close(self):
            self.h.real.close()

    monkeypatch.setitem(sys.modules, "torch", _fake_torch())
    monkeypatch.setattr(kb, "EnsembleSolver", FakeSolver)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--members", "8", "--steps", "2", "--warmup", "2", "--no-cpu", "--parity-members", "0"])
    for v in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        monkeypatch.delenv(v, raising=False)
    bench.main()
    line = [l for l in capsys.readouterr().out.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["metric"] == "ensemble_crn_solves_per_sec" and d["unit"] == "solves/s" and d["n_gpus"] == 1
    assert d["steps"] == 2 and d["warmup"] == 2 and d["dtype"] == "f64" and d["scaling"] == "weak"
    assert abs(d["value"] - 8 / 1.5) < 1e-9 and d["gpu_launches"] == 2 * 15 * 110
    cfg = d["config"]
    assert cfg["workload"].startswith("C3") and cfg["members_per_gpu"] == 8 and cfg["members_ok"] == 8
    assert cfg["ordering"].startswith("Sloan 1:2") and cfg["lu_padded_slots"] == 107787 and cfg["nnzLU"] == 79275
    assert set(d["kernels"]) >= set(bench.PHASE_LAUNCHES) | {"rhs+jacobian"}
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["kernel"] in bench.PHASE_LAUNCHES
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and "whole_solve" in r
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] > 0
    assert d["kernels"]["lu"]["frac_fp64"] > 0 and "cpu_baseline" not in d
    # with another member count than the capture's the ncu traffic figures are not carried over
    assert all(d["kernels"][k]["traffic"] is None for k in bench.PHASE_LAUNCHES)
