"""bench.py pieces that run without a GPU: the reference arm's JSON contract and the helpers."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_contract():
    env = dict(os.environ, OMP_NUM_THREADS="1")  # what torchrun exports; the arm must still use all cores
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--qubits", "12",
                           "--steps", "3", "--warmup", "3"], capture_output=True, text=True, env=env, timeout=300)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1  # exactly one JSON line on stdout
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "impl"):
        assert key in d, key
    assert d["impl"] == "reference" and d["vs_baseline"] is None and d["higher_is_better"] is True
    assert d["metric"] == "hpsi_hbm_gbs_in_sweep" and d["unit"] == "GB/s" and d["value"] > 0
    assert "workload" in d["config"] and "12 qubits" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_non_zero_ranks_of_the_reference_arm_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--qubits", "12",
                           "--gpus", "2"], capture_output=True, text=True, env=env, timeout=120)
    assert proc.returncode == 0 and proc.stdout.strip() == ""


def test_clock_sampler_degrades_without_nvml():
    sys.path.insert(0, ROOT)
    import bench

    s = bench.ClockSampler(0)
    s.start()
    out = s.stop()
    assert set(out) >= {"sm_mhz", "sm_max_mhz", "reasons", "samples"}


def test_both_arms_describe_the_same_config():
    """The driver compares the `config` objects of the two arms: they come from one function."""
    import bench

    for n, world in ((25, 1), (26, 2), (28, 8)):
        a, b = bench.make_config(n, world), bench.make_config(n, world)
        assert a == b and a["qubits"] == n and a["amplitudes_per_gpu"] == 1 << 25
        assert set(a) == {"workload", "qubits", "amplitudes_per_gpu", "l2", "parallelism", "stepper"}
