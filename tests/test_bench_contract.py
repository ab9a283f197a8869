"""CPU: the parts of bench.py's contract that run without a GPU -- the reference arm's JSON line (rank 0 only under
torchrun), the workload description, and that the B200 arm refuses to run without a device instead of falling back."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          env=e, timeout=600)


def test_reference_arm_line():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--crop-w", "100")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "1080p detect+recognize frames/sec" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1 and d["data"] == "synthetic"
    assert d["value"] > 0 and d["ms_per_step"] == pytest.approx(1e3 / d["value"], rel=1e-6)
    cfg = d["config"]
    assert cfg["workload"].startswith("configs[2]") and "32x100" in cfg["workload"]
    assert cfg["frame"] == [1080, 1920] and cfg["det"] == [736, 1312] and cfg["crop"] == [32, 100]
    assert "model" not in cfg
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"]
    assert "50.0 boxes/frame" in cb["sample"]                   # the planted plane yields the workload's 50 boxes
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_print_nothing():
    r = run_bench("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                  env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_b200_arm_refuses_without_device():
    r = run_bench("--steps", "1", "--no-cpu-baseline")
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_hbm_stage_rooflines_from_a_committed_profile():
    """bench.py's `hbm_stages` table, computed offline from a committed --profile-out file of the B200 run."""
    sys.path.insert(0, ROOT)
    import bench
    d = json.load(open(os.path.join(ROOT, "profiles", "r01_ops_final_events.json")))
    rows = bench.hbm_stage_rooflines(d["ops"]["stages"], d["steps"], d["batch"], 800, 800 * 4800 * 3, 6531.9)
    assert [r["stage"] for r in rows] == ["preprocess", "head_tail", "boxes", "crop", "ctc"]
    pre = rows[0]
    assert pre["algorithmic_bytes_per_step"] == 16 * (1080 * 1920 * 3 + 3 * 736 * 1312 * 2)
    assert pre["achieved_gbs"] == pytest.approx(pre["algorithmic_bytes_per_step"] / (pre["ms_per_step"] * 1e-3) / 1e9)
    assert all(0 < r["frac_of_hbm_peak"] < 1 for r in rows)
    assert bench.hbm_stage_rooflines([], 3, 16, 0, 0, 6531.9) == []
