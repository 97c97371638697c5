"""bench.py's reference arm (CPU only) against the JSON-line contract: one line on stdout, the keys the driver
reads, `impl = reference`, a cpu_baseline describing the run; ranks other than 0 print nothing and exit 0."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                           "--warmup", "0", "--cpu-steps", "2"], capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_prints_one_json_line():
    r = _run({"RANK": "0", "WORLD_SIZE": "2"})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["steps"] == 1 and d["higher_is_better"] is True
    assert d["metric"] == "wavefunction grid-pt*timesteps/s" and d["unit"] == "grid-pt*timesteps/s"
    assert d["config"]["workload"] == "krotov_batch_8192x2048" and d["vs_baseline"] is None and d["dtype"] == "f64"
    cb = d["cpu_baseline"]
    from oracle import ref_loader
    want_kind = "reference" if ref_loader.available_root() else "port"      # oracle/_ref travels to the GPU box
    assert cb["kind"] == want_kind and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] > 0
    for key in ("runs_per_gpu", "nt", "np", "nch", "nlevs"):                # the same config object as the GPU arm
        assert key in d["config"], key
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_are_silent():
    r = _run({"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


import pytest


@pytest.mark.gpu
def test_gpu_arm_line_has_the_contract_keys():
    """A small instance of the bench workload (296 runs x 4 steps) through bench.py itself: one JSON line with the
    base contract's keys plus roofline / e2e / clocks / gpu_launches, all finite and positive."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--runs", "296", "--nt", "4", "--steps", "1",
                        "--warmup", "3", "--no-cpu", "--ut-iterations", "2", "--api-runs", "36", "--api-nt", "8"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "clocks", "gpu_launches", "cpu_baseline"):
        assert key in d, key
    assert d["n_gpus"] == 1 and d["value"] > 0 and d["gpu_launches"] > 0 and d["cpu_baseline"] is None
    roof = d["roofline"]
    assert roof["unit"] == "TFLOP/s" and 0.0 < roof["frac"] < 1.0 and abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-12
    e2e = d["e2e"]
    assert e2e["value"] > 0 and e2e["h2d_bytes_per_step"] == e2e["d2h_bytes_per_step"] == 296 * 2 * 2048 * 16
    assert d["clocks"]["sm_max_mhz"] and isinstance(d["clocks"]["reasons"], list)
    ut = d["extra"]["ut_jx_tsweep"]
    assert ut["ut_iterations_per_s"] > 0 and ut["runs_per_rank"] == [400] and ut["imbalance_max_over_mean"] == 1.0
    assert 0.0 < ut["hbm"]["frac"] < 1.0
    api = d["e2e_api"]
    assert api["runs"] == 36 and api["value"] > 0 and api["krotov_iterations"] == 2 and api["setup_s"] > 0
