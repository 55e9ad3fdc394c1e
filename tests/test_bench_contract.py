"""CPU test of bench.py's reference arm: it runs without a GPU and prints ONE JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--rows", "20000"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "impl"):
        assert key in d, key
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["unit"] == "queries/s" and d["value"] > 0 and "workload" in d["config"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    # the line's ms_per_step is what was MEASURED (one bounded sample per step); the scale to the full collection is separate
    assert d["steps"] * d["ms_per_step"] / 1e3 < 60 and d["sample_scale"] >= 1.0
    assert abs(d["value"] * d["ms_per_step"] * d["sample_scale"] / 1e3 - 1.0) < 1e-6      # one query per step at batch 1


def test_both_arms_print_the_same_config_keys():
    sys.path.insert(0, ROOT)
    import bench
    for n in (1, 8):
        cfg = bench.workload_config("cfg3", n)
        assert cfg["rows"] == 100_000_000 and cfg["rows_per_gpu"] == 100_000_000 // n and cfg["batch"] == 1 and cfg["k"] == 10
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--rows", "20000"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    d = json.loads(r.stdout.strip().splitlines()[-1])
    w = list(bench.WORKLOADS["cfg3"])
    w[0] = 20000
    bench.WORKLOADS["cfg3"] = tuple(w)
    assert d["config"] == bench.workload_config("cfg3", 1)      # the vrod arm builds its config with the same call


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--rows", "20000"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_uses_all_host_cores_under_torchrun():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm (rank 0 of an N>1 launch) must still be
    the all-cores CPU run that the N=1 arm is, and say how many threads it used."""
    env = dict(os.environ, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2",
                        "--warmup", "1", "--rows", "20000"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
