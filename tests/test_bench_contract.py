"""bench.py, reference arm (CPU only): the JSON line the driver parses.  The arm times the plain-C
oracle (the Julia reference cannot run in this image) on a bounded sample of the bench workload;
under torchrun only rank 0 runs it."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, *flags):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                           "--warmup", "0", *flags], capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_line():
    r = _run({}, "--cpu-sample", "2")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ensemble_crn_solves_per_sec" and d["unit"] == "solves/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["dtype"] == "f64"
    assert d["value"] > 0 and abs(d["value"] - 2 / (d["ms_per_step"] * 1e-3)) < 1e-9 * d["value"]
    assert d["config"]["workload"].startswith("C3") and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and 1 <= cb["cores"] <= 2 and "2 members" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_do_nothing():
    r = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2")
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_traffic_capture_is_used_per_phase_only_while_sources_and_plan_match():
    """bench.py copies the DRAM bytes of profiles/r02_traffic.json (an ncu capture) into a phase's
    `traffic` only while the kernel sources that phase is compiled from are unchanged and — for the
    factorisation and the sweeps, which stream the factors — the plan has the padded storage of the capture."""
    sys.path.insert(0, ROOT)
    import bench
    d = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
    now, padded = bench.kernel_source_hashes(), d["plan"]["padded"]
    got = bench.load_traffic(d["workload"], d["members"], padded)
    for ph, (files, _plan) in bench.TRAFFIC_DEPENDS.items():
        assert (ph in got) == all(now[f] == d["source_hashes"][f] for f in files), ph
    other = bench.load_traffic(d["workload"], d["members"], padded + 1)
    assert "lu" not in other and "stage_sweeps" not in other
    assert {k: v for k, v in got.items() if not bench.TRAFFIC_DEPENDS[k][1]} == other
    assert bench.load_traffic("c4", d["members"], padded) == {} and bench.load_traffic(d["workload"], 1, padded) == {}
    assert bench.code_hash("a = 1;  // note\n\n// only a comment\nb = 2;") == bench.code_hash("a = 1;\nb = 2;   // other note")
