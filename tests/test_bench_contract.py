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


def test_oracle_is_only_imported_where_it_may_be():
    """The product package never imports oracle/, and bench.py imports it only inside cpu_models() (the cpu_baseline
    leg and the reference arm); the B200 arm draws weights, frames and the planted plane from the package's synthetic.py."""
    import ast
    import glob
    pkg = os.path.join(ROOT, "video_text_detection_system_b200")
    for path in glob.glob(os.path.join(pkg, "*.py")):
        tree = ast.parse(open(path).read())
        for n in ast.walk(tree):
            if isinstance(n, ast.ImportFrom):
                assert (n.module or "").split(".")[0] != "oracle", path
            if isinstance(n, ast.Import):
                assert all(a.name.split(".")[0] != "oracle" for a in n.names), path
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    where = []
    for fn in [n for n in tree.body if isinstance(n, ast.FunctionDef)]:
        for n in ast.walk(fn):
            if isinstance(n, ast.ImportFrom) and (n.module or "").split(".")[0] == "oracle":
                where.append(fn.name)
    assert where == ["cpu_models"]
    assert not any(isinstance(n, (ast.Import, ast.ImportFrom)) and "oracle" in ast.dump(n) for n in tree.body)


def test_synthetic_workload_is_seeded_and_plants_fifty_boxes():
    """bench.py's inputs: the random-init state dicts are deterministic and in the reference's key layout, and on the CPU
    path (oracle) the planted plane yields exactly the workload's 50 boxes per frame with these weights."""
    from oracle import port
    from video_text_detection_system_b200 import synthetic
    a, ar = synthetic.random_state_dicts(seed=0)
    b, br = synthetic.random_state_dicts(seed=0)
    assert all(torch.equal(a[k], b[k]) for k in a) and all(torch.equal(ar[k], br[k]) for k in ar)
    det, rec = port.build_dbnet("resnet18", seed=0), port.build_crnn(seed=0)
    assert list(det.state_dict()) == list(a) and list(rec.state_dict()) == list(ar)
    det.load_state_dict(a)
    rec.load_state_dict(ar)
    frame = synthetic.synthetic_frames(1, 1080, 1920, seed=5)[0]
    bias = synthetic.planted_logit_bias(1, 736, 1312, seed=7, boxes=50)
    regions = port.process_frame(det.eval(), rec.eval(), frame, 0.5, 736, 1312, 128, torch.from_numpy(bias)[None], per_crop=False)
    assert len(regions) == 50
