#!/usr/bin/env python
"""Refine experiments on the GPU box: wall time of one training-iteration end (energy sweeps + sdt_refine) on the
config-2 forest for a list of tuning settings, measured like bench.py's per_iteration.refine_ms (every timed refine
follows a splat of 4 Mi records; two untimed iterations first).
    python tools/refine_bench.py base helper_ctas_per_sm=2 use_pdl=0,use_graph=0
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from practical_path_guiding_lab_b200 import SDTree, synthetic as syn  # noqa: E402

dev = torch.device("cuda", 0)
n = 1 << 22
rec = syn.Scene().records(4, n)
d_rec = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in rec.items()}
frozen = bench.frozen_tree()
for spec in sys.argv[1:] or ["base"]:
    tree = SDTree(device=0, kd_max_depth=20, quad_max_depth=20, store_nee=False)
    if spec != "base":
        for kv in spec.split(","):
            k, v = kv.split("=")
            tree.set_tuning(k, int(v))
    tree.upload(frozen)
    tree.set_max_leaf_size(1e9)
    reps, ms, l0 = 8, [], 0
    for it in range(2 + reps):
        tree.splat_records(d_rec['position'], d_rec['direction'], d_rec['radiance'], d_rec['wo_pdf'])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        l0 = tree.kernel_launches()
        e0.record()
        tree.refine()
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            ms.append(e0.elapsed_time(e1))
    print(json.dumps({"spec": spec, "refine_ms_mean": float(np.mean(ms)), "min": float(np.min(ms)), "max": float(np.max(ms)),
                      "launches": tree.kernel_launches() - l0, "n_quad": tree.sizes()["n_quad"]}), flush=True)
    del tree
