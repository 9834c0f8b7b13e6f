"""bench.py's reference arm runs on the host alone (no GPU) and prints the JSON line of the measurement contract;
the product arm must refuse to run without a GPU instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*argv, env=None):
    e = dict(os.environ, CUDA_VISIBLE_DEVICES="") if env is None else env
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *argv], capture_output=True, text=True,
                          cwd=ROOT, env=e, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "1080p", "--cpu-budget", "4")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "Mpixel/s" and j["higher_is_better"] is True
    assert j["steps"] == 1 and j["warmup"] == 0 and j["n_gpus"] == 1 and j["value"] > 0
    assert j["config"]["workload"].startswith("configs[1]")
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and cb["sample"]
    assert f"OpenMP {cb['cores']} threads" in cb["sample"]   # the team size is set and reported, not assumed
    assert j["e2e"] == {"value": j["value"], "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["vs_baseline"] is None   # BASELINE.md publishes no number for this metric


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = _run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--workload", "1080p", env=env)
    assert r.returncode == 0 and not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_reference_arm_ignores_torchruns_single_thread_default():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm must still use the host's cores (round 1's N > 1
    reference numbers were single-threaded by accident)."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1", RANK="0", LOCAL_RANK="0", WORLD_SIZE="2")
    r = _run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--workload", "tiny", "--cpu-budget", "2", env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    j = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][0])
    assert j["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0)) and j["n_gpus"] == 2 and j["scaling"] == "strong"


def test_product_arm_fails_loudly_without_a_gpu():
    r = _run("--steps", "1", "--warmup", "0", "--no-cpu-baseline")
    assert r.returncode != 0
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


import pytest  # noqa: E402


@pytest.mark.gpu
def test_product_arm_prints_the_contract_line_on_a_gpu():
    """One short run of the default arm on cuda:0: every key of the measurement contract is present and sane."""
    r = _run("--steps", "3", "--warmup", "3", "--workload", "tiny", "--no-cpu-baseline", env=dict(os.environ))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in j, k
    assert j["n_gpus"] == 1 and j["steps"] == 3 and j["warmup"] >= 3 and j["value"] > 0 and j["unit"] == "Mpixel/s"
    assert j["gpu_launches"] == 7 * 3 and "workload" in j["config"] and "model" not in j["config"]
    assert j["e2e"]["d2h_bytes_per_step"] == 4 * 320 * 180 and j["e2e_fp32"]["d2h_bytes_per_step"] == 20 * 320 * 180
    rf = j["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and 0 < rf["frac"] < 1 and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-6
    e = j["e2e"]
    # (no ordering between e2e and value is asserted: at this tiny size both are bound by host-side launch cost)
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert "sm_mhz" in j["clocks"] and "reasons" in j["clocks"]
