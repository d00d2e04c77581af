#!/usr/bin/env python
"""Kernel experiments on the GPU box: runs bench.py (device-timed part only) for a list of
`name[:lib][:key=val,...]` variants and prints one line per variant (ms of sample / pdf / splat / step).
    python tools/kbench.py base sgrid:variants/libsdtree_sgrid.so agg::splat_aggregate=1
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = []
for spec in sys.argv[1:]:
    parts = spec.split(":")
    name = parts[0]
    lib = parts[1] if len(parts) > 1 and parts[1] else None
    tunes = parts[2].split(",") if len(parts) > 2 and parts[2] else []
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "20", "--warmup", "5", "--no-e2e", "--no-cpu-baseline"] + (["--layout", "soa"] if name.startswith("soa") else [])
    if lib:
        cmd += ["--lib", os.path.join(ROOT, lib)]
    for t in tunes:
        cmd += ["--tune", t]
    r = subprocess.run(cmd, capture_output=True, text=True)
    line = [x for x in r.stdout.splitlines() if x.startswith("{")]
    if r.returncode != 0 or not line:
        print(f"{name}: FAILED rc={r.returncode}\n{r.stderr[-2000:]}", flush=True)
        continue
    d = json.loads(line[-1])
    pk = d["roofline"]["per_kernel"]
    ex = d.get("extras", {})
    row = {"name": name, "fused": pk["sample_pdf"]["ms"], "sample": pk["sample"]["ms"], "pdf": pk["pdf"]["ms"], "splat": pk["splat"]["ms"], "step": d["ms_per_step"],
           "req_frac": {k: round(v["frac_of_request_roof"], 3) for k, v in pk.items()},
           "guided": ex.get("sdt_guided", {}).get("ms"), "guided_em": ex.get("sdt_guided", {}).get("ms_with_emitter_pdf"), "pdf_em_sep": ex.get("sdt_guided", {}).get("ms_emitter_pdf_as_a_separate_call"), "coh": ex.get("coherent_wavefront"), "path": ex.get("sdt_splat_path_data", {}).get("ms"),
           "refine": d.get("refine_ms"), "mhz": d["clocks"]["sm_mhz"]}
    rows.append(row)
    print(json.dumps(row), flush=True)
    with open(os.path.join(ROOT, "gpurun_out", f"kb_{name}.json"), "w") as f:
        f.write(line[-1] + "\n")
