"""bench.py's reference arm runs on CPU and prints ONE JSON line with the contract keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sample", "8192"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["config"]["workload"].startswith("synthetic frozen SD-tree")
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                         text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


import pytest


@pytest.mark.gpu
def test_config_shaped_passes_run_on_a_small_shape(monkeypatch):
    """bench.py's extras.config_shaped_passes (the tree work of one training pass at a scene config's wavefront shape):
    runs end to end on a small shape and reports consistent counts"""
    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    import bench
    from practical_path_guiding_lab_b200 import SDTree
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import sdt_cases as cases
    ctx = cases.Ctx(make=lambda **kw: SDTree(device=0, **kw),
                    dev=lambda x: None if x is None else torch.from_numpy(np.ascontiguousarray(x).view(np.int32) if np.asarray(x).dtype == np.uint32
                                                                          else np.ascontiguousarray(x)).cuda(),
                    host=lambda x: x if isinstance(x, np.ndarray) else x.cpu().numpy())
    t, cur, prev = cases.train(ctx, iters=2, n=8000, max_leaf=400)
    monkeypatch.setattr(bench, "CONFIG_SHAPES", [("tiny 32x16", 32, 16, 5), ("small 64x64", 64, 64, 3)])
    r = bench.config_shaped_passes(t, torch.device("cuda", 0), 0, reps=2)
    assert len(r["passes"]) == 2
    for p in r["passes"]:
        assert p["record_slots"] == p["lanes_per_pass"] * p["max_depth"]
        assert p["lanes_per_pass"] <= p["path_vertices_per_pass"] <= p["record_slots"]       # every lane has at least its first vertex
        assert p["tree_ms_per_pass"] > 0 and p["launches_per_pass"] >= 2
    assert t.sizes()["error"] == 0
