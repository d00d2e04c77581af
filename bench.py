#!/usr/bin/env python
"""bench.py -- guided samples/s of the SD-tree hot path on the synthetic frozen-tree
workload of BASELINE.json configs[1].

One STEP = one pass of the hot path over one wavefront of n = 2^24 path vertices:
    KDTree.sample  (spatial descent + quadtree sample + pdf of the sample, A1+A3+A4)
  + KDTree.pdf     (quadtree pdf of a given direction at the same vertex,  A1+A4)      } sdt_sample_pdf: one call, one
                                                                                         } spatial descent for both
  + sdt_splat_records (spatial descent + quadtree descent + accumulation, A6+A7)
(the two query operations are also timed as the separate calls sdt_sample / sdt_pdf: roofline.per_kernel)
on a frozen tree (~4k spatial leaves, quadtree depth <= 20) that the library itself
trained on the synthetic records (splat + device-side refine, 6 iterations).
`value` = vertices through the whole step per second with inputs resident in HBM;
`e2e` = the same three C-ABI calls with HOST buffers (pinned), H2D/D2H inside the timed
region.  `--impl reference` times the CPU restatement of the reference (oracle/) on a
bounded sample of the same workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--n LANES]
    torchrun --nproc-per-node N bench.py --gpus N ...        (one rank per GPU)
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "guided samples/s (sample + pdf + splat per path vertex, synthetic frozen SD-tree)"
UNIT = "samples/s"
N_DEFAULT = 1 << 24
CPU_SAMPLE = 1 << 21


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=N_DEFAULT, help="path vertices per step per GPU")
    ap.add_argument("--cpu-sample", type=int, default=CPU_SAMPLE)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--tune", action="append", default=[], help="key=value passed to sdt_set_tuning (repeatable)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config-shapes", action="store_true", help="skip extras.config_shaped_passes (the scene configs' wavefront shapes)")
    ap.add_argument("--layout", default="aos", choices=["soa", "aos"],
                    help="device-resident vectors interleaved (n,3) or as separate component planes (Dr.Jit's Vector3f layout); measured within 1 %% of each other")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin the rank to its GPU's NUMA-local CPUs (e2e experiments)")
    ap.add_argument("--lib", default=None, help="path of an alternative build of libsdtree.so (kernel experiments)")
    return ap.parse_args()


def workload_config(n, world, layout="aos"):
    from practical_path_guiding_lab_b200 import synthetic as syn
    return {"workload": "synthetic frozen SD-tree microbench (BASELINE configs[1])",
            "queries_per_step_per_gpu": n, "ops_per_query": ["sample", "pdf", "splat"],
            "tree_build": {"iterations": syn.BUILD_ITERS, "records_iter0": syn.BUILD_N0, "c": syn.BUILD_C,
                           "seed": syn.BUILD_SEED},
            "l2_policy": "inputs (>= 200 MB per array set) larger than the 126 MB L2; the tree (~13 MB of records) is meant to stay L2-resident",
            "vector_layout": {"soa": "component planes (Dr.Jit Vector3f: x, y, z separate arrays) for the device-resident step; the host-buffer e2e leg sends interleaved (n,3) arrays",
                              "aos": "interleaved (n,3)"}[layout],
            "parallelism": f"replicated tree, vertices sharded x{world}"}


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML, every ~2 ms)."""

    def __init__(self, index):
        self.index = index
        self.sm, self.reasons, self.mx = [], 0, None
        self.stop = False
        self.th = None
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self.stop:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.reasons |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv is not None:
            self.th = threading.Thread(target=self._run, daemon=True)
            self.th.start()
        return self

    def __exit__(self, *a):
        if self.nv is not None:
            try:        # one reading taken with the work still queued / just drained, whatever the thread managed
                self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                self.reasons |= int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
        self.stop = True
        if self.th is not None:
            self.th.join(timeout=10)

    def summary(self):
        if self.nv is None or not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.mx, "reasons": ["unavailable"]}
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.mx, "reasons": [k for k, bit in names.items() if self.reasons & bit],
                "samples": len(sm), "sm_mhz_min": sm[0]}


# ------------------------------------------------------------------------------ CPU port (oracle) arm
def frozen_tree():
    """the frozen benchmark tree (npz schema), built once by the numpy oracle and used as INPUT DATA by both arms
    (oracle/bench_tree.py; cached in the temp directory so the two arms of one run build it once)"""
    from oracle import bench_tree
    return bench_tree.frozen_tree_arrays()


class CpuPort:
    """the reference's per-vertex operations in C + OpenMP on all host cores (oracle/sdtree_port.c,
    checked against the numpy oracle in tests/test_oracle_port.py) on a tree in the npz schema"""

    def __init__(self, tree_arrays):
        from oracle.port import PortTree
        self.prev = PortTree(tree_arrays)
        z = dict(tree_arrays)
        z['kdtree_vertCount'] = np.zeros_like(np.asarray(tree_arrays['kdtree_vertCount'], np.float32))
        z['quadtree_irradiance'] = np.zeros_like(np.asarray(tree_arrays['quadtree_irradiance'], np.float32))
        self.cur = PortTree(z)
        # torchrun exports OMP_NUM_THREADS=1 to its workers: ask for every core this process may run on
        want = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        self.cores = self.prev.set_threads(want)
        self.cur.set_threads(want)

    def step(self, pos, dirs, rec, seed):
        self.prev.sample(pos, seed)
        self.prev.pdf(pos, dirs)
        self.cur.splat(rec['position'], rec['direction'], rec['radiance'], rec['wo_pdf'])

    def timed(self, pos, dirs, rec, min_seconds=10.0, min_reps=2):
        self.step(pos[:4096], dirs[:4096], {k: v[:4096] for k, v in rec.items()}, 3)
        t0 = time.perf_counter()
        reps = 0
        while reps < min_reps or time.perf_counter() - t0 < min_seconds:
            self.step(pos, dirs, rec, 3)
            reps += 1
        return (time.perf_counter() - t0) / reps, reps


def cpu_inputs(m):
    from practical_path_guiding_lab_b200 import synthetic as syn
    return syn.uniform_box(1, m), syn.uniform_sphere(2, m), syn.Scene().records(4, m)


def run_reference(args):
    """reference arm: the CPU restatement of the reference's algorithm on the host cores -- Mitsuba 3 /
    Dr.Jit are not installable in this image, so the tree is trained by the numpy
    oracle (oracle/sdtree_oracle.py) and the per-vertex operations run in its C + OpenMP port"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    m = int(min(args.n, args.cpu_sample))
    port = CpuPort(frozen_tree())
    pos, dirs, rec = cpu_inputs(m)
    for _ in range(args.warmup):
        port.step(pos, dirs, rec, 3)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        port.step(pos, dirs, rec, 3)
    dt = (time.perf_counter() - t0) / args.steps
    v = m / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args.n, args.gpus),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": port.cores, "kind": "port",
                             "sample": f"{m} of the {args.n} vertices per step (C + OpenMP port of the reference's per-vertex operations, {port.cores} threads of {os.cpu_count()} host cores; the same frozen tree as the B200 arm, built by the numpy oracle)"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def bind_to_gpu_numa_node(index):
    """pin this rank's threads to the CPUs NVML names as local to its GPU, BEFORE the pinned staging buffers are allocated
    (first touch decides the NUMA node of pinned memory): a rank on the far socket pays the inter-socket link on every
    H2D / D2H byte of the end-to-end leg.  -> (applied, cpus now allowed)"""
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index))
        return True, len(os.sched_getaffinity(0))
    except Exception:
        return False, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


# ------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from practical_path_guiding_lab_b200 import SDTree, synthetic as syn
    from practical_path_guiding_lab_b200.build import build
    build()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_bound, cpus_allowed = bind_to_gpu_numa_node(local) if not args.no_numa_bind else (False, len(os.sched_getaffinity(0)))
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.n
    tree = SDTree(device=local, kd_max_depth=20, quad_max_depth=20, store_nee=False, lib_path=args.lib)
    for kv in args.tune:
        key, val = kv.split("=")
        tree.set_tuning(key, int(val))
    to_dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    allreduce = None
    if world > 1:
        ids = [tree.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        tree.comm_init(ids[0], rank, world)
        allreduce = lambda records=None: tree.allreduce(torch.cuda.current_stream().cuda_stream, records_all_ranks=records)
    t_build0 = time.perf_counter()
    syn.build_tree(tree, to_dev=to_dev, rank=rank, world=world, allreduce=allreduce)
    torch.cuda.synchronize()
    sizes_trained = tree.sizes()
    t_build = time.perf_counter() - t_build0
    assert sizes_trained["error"] == 0, sizes_trained
    # the timed region runs on the SAME frozen tree as the reference arm: the oracle-built one, uploaded in the
    # reference's npz schema (the device-trained tree above differs from it by a few quadtree nodes: fp32 atomics)
    tree.upload(frozen_tree())
    sizes = tree.sizes()
    assert sizes["error"] == 0 and sizes["n_kd"] == sizes_trained["n_kd"], (sizes, sizes_trained)

    # inputs of this rank's shard (host, pinned for the e2e arm) and their device copies
    scene = syn.Scene()
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_pos = pin(syn.uniform_box(1 + 1000 * rank, n))
    h_dir = pin(syn.uniform_sphere(2 + 1000 * rank, n))
    rec = scene.records(4 + 1000 * rank, n)
    h_rec = {k: pin(v) for k, v in rec.items()}
    d_pos, d_dir = h_pos.to(dev), h_dir.to(dev)
    d_rec = {k: v.to(dev) for k, v in h_rec.items()}
    o_dir = torch.empty(n, 3, device=dev)
    if args.layout == "soa":
        # component planes: mi.Vector3f / Point3f are three separate arrays in Dr.Jit, so this is the layout the integrator's
        # buffers have; a lane's x, y, z are then three coalesced accesses instead of three 12-byte-strided ones
        planes = lambda t: tuple(t[:, k].contiguous() for k in range(t.shape[1]))
        d_pos, d_dir, o_dir = planes(d_pos), planes(d_dir), planes(o_dir)
        d_rec = dict(d_rec, position=planes(d_rec['position']), direction=planes(d_rec['direction']))
    sub = lambda v, ix: tuple(c[ix].contiguous() for c in v) if isinstance(v, tuple) else v[ix].contiguous()      # rows of a vector array in either layout
    o_pdf = torch.empty(n, device=dev)
    o_pdf2 = torch.empty(n, device=dev)
    lane0 = rank * n

    def step(ev=None):
        if ev:
            ev[0].record()
        tree.sample_pdf(d_pos, d_dir, seed=3, lane_offset=lane0, out=(o_dir, o_pdf, o_pdf2))
        if ev:
            ev[1].record()
        tree.splat_records(d_rec['position'], d_rec['direction'], d_rec['radiance'], d_rec['wo_pdf'])
        if ev:
            ev[2].record()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    barrier()
    launches0 = tree.kernel_launches()
    with ClockSampler(local) as clk:
        barrier()
        for k in range(args.steps):
            step(evs[k])
        barrier()
    launches = tree.kernel_launches() - launches0
    total_ms = evs[0][0].elapsed_time(evs[-1][2])
    f_ms = [float(np.mean([e[j].elapsed_time(e[j + 1]) for e in evs])) for j in range(2)]      # fused query call, splat
    t = torch.tensor([total_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = n * world / (ms_per_step * 1e-3)

    def timeit(fn, reps=10):
        for _ in range(3):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    # the two query operations as separate calls (what the step's fused call replaces)
    k_ms = {"sample_pdf": f_ms[0], "splat": f_ms[1],
            "sample": timeit(lambda: tree.sample(d_pos, seed=3, lane_offset=lane0, out=(o_dir, o_pdf))),
            "pdf": timeit(lambda: tree.pdf(d_pos, d_dir, out=o_pdf2))}

    # ---- algorithmic bytes from the measured depths of a 2^20-lane subset
    m = min(n, 1 << 20)
    tr = tree.download(0)
    leaf, _ = tree.locate(sub(d_pos, slice(0, m)))
    ds = tr['kdtree_depth'][leaf.cpu().numpy().view(np.uint32)].astype(np.float64).mean()
    _, _, dbg = tree.sample(sub(d_pos, slice(0, m)), seed=3, lane_offset=lane0, debug=True)
    dbg = dbg.cpu().numpy().view(np.uint32)
    dq_s = tr['quadtree_depth'][dbg[:, 2]].astype(np.float64).mean()
    redescend = float((dbg[:, 2] != dbg[:, 3]).mean())
    _, dbgp = tree.pdf(sub(d_pos, slice(0, m)), sub(d_dir, slice(0, m)), debug=True)
    dq_pl = tr['quadtree_depth'][dbgp.cpu().numpy().view(np.uint32)[:, 2]].astype(np.float64)
    dq_p = dq_pl.mean()
    # depth reached by the splat's directions (drawn from the lobes, deeper than the pdf's uniform ones): a pdf query with
    # the record's own direction lands on the node the splat updates
    cxy = rec['direction'][:m].astype(np.float64)
    ct = 2.0 * cxy[:, 1] - 1.0
    st_ = np.sqrt(np.maximum(0.0, 1.0 - ct * ct))
    rdir = torch.from_numpy(np.stack([st_ * np.cos(2 * np.pi * cxy[:, 0]), st_ * np.sin(2 * np.pi * cxy[:, 0]), ct], 1).astype(np.float32)).to(dev)
    _, dbgr = tree.pdf(sub(d_rec['position'], slice(0, m)), rdir, debug=True)
    dq_rl = tr['quadtree_depth'][dbgr.cpu().numpy().view(np.uint32)[:, 2]].astype(np.float64)
    dq_r = dq_rl.mean()
    JL, J2 = 5, 8                                                # depth <= 5 ends inside the root jump table, <= 8 inside a second-stage table
    below_p, below_r = np.maximum(dq_pl - J2, 0).mean(), np.maximum(dq_rl - J2, 0).mean()      # records walked below the tables
    mid_p, mid_r = float((dq_pl > JL).mean()), float((dq_rl > JL).mean())                      # one more gather: second-stage entry (or the level-5 record)
    deep_p = float((dq_pl > J2).mean())                                                        # pdf below the tables: the leaf's path product is a gather of its own
    bytes_q = {"sample": 28 + 4 * ds + 4 + 20 * dq_s,            # ONE quadtree descent (fused pdf); re-descents not counted
               "pdf": 28 + 4 * ds + 4 + 20 * dq_p + 4,
               "splat": 28 + 4 * ds + 4 + 4 * dq_r + 8}            # leaf-only update + sweep
    bytes_q["sample_pdf"] = 44 + 4 * ds + 4 + 20 * dq_s + 20 * dq_p + 4          # pos 12 + dir 12 in, 16 + 4 out; one spatial descent
    stream_b = {"sample": 28.0, "pdf": 28.0, "splat": 28.0, "sample_pdf": 44.0}  # query / record bytes streamed from and to HBM
    # divergent sector requests per query actually issued (one lane, one unrelated 32 B sector, one slot of the SM's L1 -> L2
    # request port): sampling = a record per level + the leaf's path product; pdf = a jump-table entry (+ a second-stage
    # entry beyond level 5, + records and the leaf's product beyond level 8); splat = the same entries + records + the leaf RED
    req_q = {"sample": dq_s + 1.0, "pdf": 1.0 + mid_p + below_p + deep_p, "splat": 1.0 + mid_r + below_r + 1.0}
    req_q["sample_pdf"] = req_q["sample"] + req_q["pdf"]
    names = ["sample_pdf", "splat", "sample", "pdf"]
    dom = "sample_pdf"
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    l2_gbs = tree.measure_l2(32 << 20, 50)                    # sequential sweep of an L2-resident set (ld.global.cg)
    gather_gbs = tree.measure_gather(16 << 20, 200, True)     # random 32 B sectors of an L2-resident set, through L1 (the descents' pattern)
    gather_cg_gbs = tree.measure_gather(16 << 20, 200, False)
    tree_b = {k: bytes_q[k] - stream_b[k] for k in names}     # tree bytes: served by shared memory / L1 / L2
    klane = {"sample_pdf": "SamplePdfLane", "splat": "SplatRecordsLane", "sample": "SampleLane", "pdf": "PdfLane"}
    traffic = None
    try:        # DRAM bytes per launch of this kernel from the committed ncu --set full capture (same 16 Mi-vertex launch)
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if n == N_DEFAULT:
            key = [k for k in tj if k.startswith(klane[dom])][0]
            traffic = tj[key]["dram_bytes_read"] + tj[key]["dram_bytes_write"]
    except Exception:
        pass
    req_peak = gather_gbs / 32.0 if gather_gbs else None       # G requests/s: the probe delivers one sector per lane request

    def per_kernel(k):
        sec = k_ms[k] * 1e-3
        tree_gbs, stream_gbs = tree_b[k] * n / sec / 1e9, stream_b[k] * n / sec / 1e9
        greq = req_q[k] * n / sec / 1e9
        return {"ms": k_ms[k], "bytes_per_query": bytes_q[k], "tree_bytes_per_query": tree_b[k], "stream_bytes_per_query": stream_b[k],
                "tree_gbs": tree_gbs, "frac_of_l2": tree_gbs / l2_gbs if l2_gbs else None,
                "frac_of_gather_roof": tree_gbs / gather_gbs if gather_gbs else None,
                "sector_requests_per_query": req_q[k], "g_requests_per_s": greq, "frac_of_request_roof": greq / req_peak if req_peak else None,
                "stream_gbs": stream_gbs, "frac_of_hbm": stream_gbs / hbm_peak, "queries_per_s": n / sec}
    pk = {k: per_kernel(k) for k in names}
    d = pk[dom]
    roof = {"bound": "l2", "kernel": f"k_wavefront<{klane[dom]}>", "achieved": d["tree_gbs"], "peak": l2_gbs, "unit": "GB/s",
            "frac": d["frac_of_l2"], "traffic": traffic,
            "peak_source": "measured in this run by sdt_measure_l2 (32 MiB resident set, 50 passes, ld.global.cg; an L2 figure is not in "
                           "MEASURED_PEAKS.json): the friendliest access pattern there is.  The descents read ONE unrelated 32 B sector per "
                           "lane and level; that pattern's own roof is gather_roof / request_roof below",
            "what": "achieved = algorithmic TREE bytes of the dominant kernel (4 B per spatial level + 4 B root id + 20 B per quadtree "
                    "level [+4 B path product / +8 B leaf update], SURVEY 8d) x queries / its launch time (CUDA events); the bytes "
                    "streamed from / to HBM per query are accounted separately (hbm)",
            "algorithmic_bytes_per_query": bytes_q[dom], "kernel_ms": k_ms[dom],
            "gather_roof": {"gbs_via_l1": gather_gbs, "gbs_l2_only": gather_cg_gbs,
                            "how": "sdt_measure_gather: 16 MiB resident set, every lane loads one random 32 B sector per 256-bit load, 8 independent loads in flight per lane",
                            "frac": d["frac_of_gather_roof"]},
            "request_roof": {"peak_g_requests_per_s": req_peak, "achieved_g_requests_per_s": d["g_requests_per_s"], "frac": d["frac_of_request_roof"],
                             "what": "a lane that reads its own 32 B sector occupies one slot of the SM's L1 -> L2 request port whatever it needs of the "
                                     "sector (ncu: l1tex__m_l1tex2xbar_req_cycles_active is the busiest unit of the sampling kernels); the gather probe "
                                     "saturates that port (1.0 sector per clock and SM), so requests / s against it is the hardware roof of a per-lane "
                                     "descent, and 20 algorithmic bytes of every 32 B sector cap frac_of_gather_roof at 0.625"},
            "hbm": {"achieved": d["stream_gbs"], "peak": hbm_peak, "frac": d["frac_of_hbm"],
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650"},
            "per_kernel": pk,
            "step_unfused_ms": k_ms["sample"] + k_ms["pdf"] + k_ms["splat"],
            "mean_depths": {"spatial": ds, "quad_sample": dq_s, "quad_pdf": dq_p, "quad_splat": dq_r, "sample_redescend_frac": redescend,
                            "records_below_the_jump_tables": {"pdf": below_p, "splat": below_r},
                            "lanes_beyond_the_root_table": {"pdf": mid_p, "splat": mid_r}}}

    # ---- refine + allreduce wall time (per training iteration, not part of a step).  Steady state: the refine replays a
    # CUDA graph captured on first use (one per buffer parity), so two untimed iterations come first; every timed refine
    # follows a splat, i.e. it includes the bottom-up sweeps of the statistics like a training iteration's does.
    tree.set_max_leaf_size(1e9)            # statistics are K steps of the same records: keep the spatial tree frozen
    ms_ar, ms_rf, reps_rf = 0.0, 0.0, 4
    for it in range(2 + reps_rf):
        tree.splat_records(d_rec['position'], d_rec['direction'], d_rec['radiance'], d_rec['wo_pdf'])
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        torch.cuda.synchronize()
        e0.record()
        if allreduce:
            allreduce(n * world)             # every rank splatted n records: no leaf can have counted more (sdt_hint_records)
        e1.record()
        tree.refine()
        e2.record()
        torch.cuda.synchronize()
        if it >= 2:
            ms_ar += e0.elapsed_time(e1) / reps_rf
            ms_rf += e1.elapsed_time(e2) / reps_rf
    per_iter = {"allreduce_ms": ms_ar if allreduce else 0.0, "refine_ms": ms_rf,
                "what": f"mean of {reps_rf} training-iteration ends (sweeps + refine of the {sizes['n_quad']}-node forest) after 2 untimed ones"}
    tree.upload(frozen_tree())             # back to the frozen benchmark tree for the extras / e2e legs

    # ---- extras (not part of the step): the integrator's own entry points on the same vertices
    extras = {}
    try:
        gg = torch.Generator(device=dev).manual_seed(99 + rank)
        mode = (torch.rand(n, device=dev, generator=gg) * 2.2).to(torch.uint8).clamp_(max=2)      # ~45 % sample, ~45 % pdf+mixture, ~10 % idle
        bp = torch.rand(n, device=dev, generator=gg)
        bv = torch.rand(n, 3, device=dev, generator=gg)
        g_dir = torch.zeros(n, 3, device=dev); g_sp = torch.zeros(n, device=dev); g_wp = torch.zeros(n, device=dev); g_w = torch.zeros(n, 3, device=dev)
        md = 8
        rays = n // md
        lfin = torch.rand(rays, 3, device=dev, generator=gg) * 4
        tr_ = torch.rand(n, 3, device=dev, generator=gg)
        tb_ = torch.rand(n, 3, device=dev, generator=gg)
        bs_ = torch.rand(n, 3, device=dev, generator=gg)
        act_ = (torch.rand(n, device=dev, generator=gg) < 0.6).to(torch.uint8)

        ms_g = timeit(lambda: tree.guided(d_pos, mode, wo=d_dir, seed=5, lane_offset=lane0, bsdf_pdf=bp, bsdf_value=bv,
                                          dir_out=g_dir, sdtree_pdf_out=g_sp, wo_pdf_out=g_wp, weight_out=g_w))
        em_act = (torch.rand(n, device=dev, generator=gg) < 0.8).to(torch.uint8)
        g_ep = torch.ones(n, device=dev)
        ms_ge = timeit(lambda: tree.guided(d_pos, mode, wo=d_dir, seed=5, lane_offset=lane0, bsdf_pdf=bp, bsdf_value=bv,
                                           dir_out=g_dir, sdtree_pdf_out=g_sp, wo_pdf_out=g_wp, weight_out=g_w,
                                           em_dir=o_dir, em_active=em_act, sdtree_pdf_em_out=g_ep))
        ms_pe = timeit(lambda: tree.pdf(d_pos, o_dir, active=em_act, out=o_pdf2))
        ms_p = timeit(lambda: tree.splat_path_data(md, lfin, tr_, tb_, bs_, d_rec['position'], d_rec['direction'], d_rec['wo_pdf'], active=act_))
        act15 = (torch.rand(n, device=dev, generator=gg) < 0.15).to(torch.uint8)
        sparse = {}
        for comp in (1, 0):
            tree.set_tuning("use_compaction", comp)
            key = "compacted" if comp else "masked"
            sparse[key] = {"splat_path_data_ms": timeit(lambda: tree.splat_path_data(md, lfin, tr_, tb_, bs_, d_rec['position'], d_rec['direction'], d_rec['wo_pdf'], active=act15)),
                           "pdf_ms": timeit(lambda: tree.pdf(d_pos, d_dir, active=act15, out=o_pdf2)),
                           "sample_ms": timeit(lambda: tree.sample(d_pos, active=act15, seed=3, out=(o_dir, o_pdf)))}
        tree.set_tuning("use_compaction", 1)
        # ---- coherent wavefront: the SAME vertices / records ordered by spatial leaf, the way the lanes of a render
        # wavefront are (neighbouring pixels see neighbouring surface points): the lanes of a warp then share a quadtree,
        # its top records are fetched once per warp instead of once per lane, and records of one warp hit the same nodes
        coherent = {}
        try:
            leaf_q, _ = tree.locate(d_pos)
            oq = torch.argsort(leaf_q.view(torch.int32).to(torch.int64), stable=True)
            c_pos, c_dir = sub(d_pos, oq), sub(d_dir, oq)
            leaf_r, _ = tree.locate(d_rec['position'])
            orr = torch.argsort(leaf_r.view(torch.int32).to(torch.int64), stable=True)
            c_rec = {k: sub(v, orr) for k, v in d_rec.items()}
            del leaf_q, leaf_r, oq, orr
            coherent["sample_ms"] = timeit(lambda: tree.sample(c_pos, seed=3, lane_offset=lane0, out=(o_dir, o_pdf)))
            coherent["pdf_ms"] = timeit(lambda: tree.pdf(c_pos, c_dir, out=o_pdf2))
            for agg in (0, 1):
                tree.set_tuning("splat_aggregate", agg)
                coherent[f"splat_ms_aggregate{agg}"] = timeit(lambda: tree.splat_records(c_rec['position'], c_rec['direction'], c_rec['radiance'], c_rec['wo_pdf']))
                coherent[f"incoherent_splat_ms_aggregate{agg}"] = timeit(lambda: tree.splat_records(d_rec['position'], d_rec['direction'], d_rec['radiance'], d_rec['wo_pdf']))
            tree.set_tuning("splat_aggregate", 0)
            coherent["what"] = ("the step's vertices / records sorted by spatial leaf (stand-in for a pixel-coherent render wavefront); "
                                "splat with and without the warp aggregation of same-node adds (splat_aggregate), on both orders")
            del c_pos, c_dir, c_rec
        except Exception as e:
            coherent = {"error": repr(e)}
        extras = {"coherent_wavefront": coherent, "sparse_wavefront_15pct_active": dict(sparse, what="same calls with 15 % of the lanes active (late bounces / numRays*max_depth record slots): lanes of a tile sorted into dense warps vs plain masking"),
                  "sdt_guided": {"ms": ms_g, "lanes_per_s": n / (ms_g * 1e-3), "what": "one bounce: ~45 % lanes sampled, ~45 % pdf + fused mixture, ~10 % idle",
                                 "ms_with_emitter_pdf": ms_ge, "ms_emitter_pdf_as_a_separate_call": ms_pe,
                                 "what_with_emitter_pdf": "the same call also returning the tree's pdf of the emitter direction (NEE MIS) on 80 % of the lanes: every tree query of a path vertex in one launch.  The emitter directions of this measurement are directions SAMPLED from the tree (deep leaves: the expensive case for a pdf); ms_emitter_pdf_as_a_separate_call = sdt_pdf of the same directions and mask"},
                  "sdt_splat_path_data": {"ms": ms_p, "slots_per_s": n / (ms_p * 1e-3), "what": f"processPathData + filter + splat fused, {n} slots (max_depth {md}), 60 % active"}}
        tree.reset_stats()
        if not args.no_config_shapes:
            try:
                extras["config_shaped_passes"] = config_shaped_passes(tree, dev, rank)
            except Exception as e:
                extras["config_shaped_passes"] = {"error": repr(e)}
            tree.reset_stats()
    except Exception as e:            # extras never break the contract line
        extras = {"error": repr(e)}

    # ---- end to end: the step's C-ABI calls on HOST buffers (pinned)
    e2e = None
    if not args.no_e2e:
        tree2 = tree
        hp, hd = h_pos.numpy(), h_dir.numpy()
        hr = {k: v.numpy() for k, v in h_rec.items()}
        ho_dir = torch.empty(n, 3).pin_memory().numpy()
        ho_pdf = torch.empty(n).pin_memory().numpy()
        ho_pdf2 = torch.empty(n).pin_memory().numpy()

        def step_host():
            tree2.sample_pdf(hp, hd, seed=3, lane_offset=lane0, out=(ho_dir, ho_pdf, ho_pdf2))
            tree2.splat_records(hr['position'], hr['direction'], hr['radiance'], hr['wo_pdf'])
            tree2.synchronize()                         # every output of the step is in host memory here
            torch.cuda.synchronize()
        ke = max(1, min(args.steps, 10))

        def timed_host():
            for _ in range(2):
                step_host()
            barrier()
            t0 = time.perf_counter()
            for _ in range(ke):
                step_host()
            barrier()
            dt = torch.tensor([(time.perf_counter() - t0) / ke], device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            return float(dt.item())
        tree2.host_wait = True                          # each call returns with its outputs on the host
        dt_wait = timed_host()
        tree2.host_wait = False                         # SDT_NO_WAIT: the calls overlap, one synchronize per step
        dt_e2e = timed_host()
        tree2.host_wait = True
        # what the link gives a plain pinned copy (context for the number above): this rank alone is not measurable under
        # torchrun without serialising the ranks, so ALL ranks copy at the same time after a barrier -- the host fabric's
        # ceiling at this GPU count (slowest rank and sum over ranks)
        big = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
        dbig = torch.empty_like(big, device=dev)
        dbig.copy_(big, non_blocking=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(4):
            dbig.copy_(big, non_blocking=True)
        torch.cuda.synchronize()
        h2d_gbs = 4 * big.numel() / (time.perf_counter() - t0) / 1e9
        # ... and the same H2D stream with the step's share of D2H traffic running against it on a second stream (the step
        # moves 20 B back for every 52 B it sends): what the fabric gives the step's mix of directions
        back = torch.empty(int(big.numel() * 20 / 52), dtype=torch.uint8).pin_memory()
        dback = torch.empty_like(back, device=dev)
        s2 = torch.cuda.Stream(device=dev)
        barrier()
        t0 = time.perf_counter()
        for _ in range(4):
            dbig.copy_(big, non_blocking=True)
            with torch.cuda.stream(s2):
                back.copy_(dback, non_blocking=True)
        torch.cuda.synchronize()
        h2d_bidir_gbs = 4 * big.numel() / (time.perf_counter() - t0) / 1e9
        del big, dbig, back, dback
        h2d_min, h2d_sum = h2d_gbs, h2d_gbs
        bidir_min, bidir_sum = h2d_bidir_gbs, h2d_bidir_gbs
        if world > 1:
            tt = torch.tensor([h2d_bidir_gbs], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MIN)
            bidir_min = float(tt.item())
            tt = torch.tensor([h2d_bidir_gbs], device=dev)
            dist.all_reduce(tt)
            bidir_sum = float(tt.item())
            tt = torch.tensor([h2d_gbs], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MIN)
            h2d_min = float(tt.item())
            tt = torch.tensor([h2d_gbs], device=dev)
            dist.all_reduce(tt)
            h2d_sum = float(tt.item())
        e2e = {"value": n * world / dt_e2e, "unit": UNIT, "ms_per_step": dt_e2e * 1e3,
               "h2d_bytes_per_step": n * (24 + 28), "d2h_bytes_per_step": n * (16 + 4),
               "ms_per_step_waiting_calls": dt_wait * 1e3, "pinned_h2d_copy_gbs": h2d_min, "pinned_h2d_copy_gbs_all_ranks": h2d_sum,
               "pinned_h2d_copy_gbs_with_d2h": bidir_min, "pinned_h2d_copy_gbs_with_d2h_all_ranks": bidir_sum,
               "h2d_gbs_in_step_all_ranks": n * world * (24 + 28) / dt_e2e / 1e9,
               "numa_bound": numa_bound, "cpus_allowed_per_rank": cpus_allowed,
               "h2d_gbs_in_step": n * (24 + 28) / dt_e2e / 1e9,
               "how": "sdt_sample_pdf (positions cross the bus once for both queries) + sdt_splat_records with SDT_HOST_PTRS | SDT_NO_WAIT on pinned host arrays and one "
                      "sdt_synchronize per step (all outputs on the host); staging copies inside the calls.  "
                      "ms_per_step_waiting_calls: the same without SDT_NO_WAIT, every call returning with its outputs on the "
                      "host; pinned_h2d_copy_gbs: a plain pinned H2D copy, all ranks copying at the same time (slowest rank; "
                      "_all_ranks: their sum; _with_d2h: the same with the step's 20:52 share of D2H bytes copied back on a second stream at the same time) -- the step is bound by the host link (h2d_gbs_in_step per rank / _all_ranks, with the "
                      "D2H traffic beside it); numa_bound: the rank was pinned to its GPU's NUMA-local CPUs before the pinned buffers were allocated"}

    # ---- CPU port of the reference, timed beside it (rank 0, N=1)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        mc = int(min(n, args.cpu_sample))
        port = CpuPort(tr)
        dt, reps = port.timed(h_pos.numpy()[:mc], h_dir.numpy()[:mc], {k: v.numpy()[:mc] for k, v in h_rec.items()})
        cpu = {"value": mc / dt, "unit": UNIT, "cores": port.cores, "kind": "port",
               "sample": f"first {mc} of the {n} vertices per step, {reps} repetitions (C + OpenMP port of the reference's per-vertex operations on the reference's SoA layout, {port.cores} threads; host has {os.cpu_count()} cores)"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": workload_config(n, world, args.layout), "clocks": clk.summary(), "e2e": e2e,
                "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu,
                "tree": {k: sizes[k] for k in ("n_kd", "kd_leaves", "n_quad", "n_interior", "n_levels")},
                "tree_device_trained": dict({k: sizes_trained[k] for k in ("n_kd", "kd_leaves", "n_quad", "n_interior", "n_levels")},
                                            build_s=t_build, what="the same schedule trained by the library (splat + allreduce + device refine); the timed region uses the oracle-built tree above, like the reference arm"),
                "allreduce_ms": per_iter["allreduce_ms"], "refine_ms": per_iter["refine_ms"],
                "e2e_per_gpu": (e2e["value"] / world if e2e else None),
                "tree_build_s": t_build, "per_iteration": per_iter, "extras": extras}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# BASELINE.json configs 1 / 3 / 4 / 5 need a renderer (Mitsuba's cuda variant; absent from this image).  What the LIBRARY
# does in one training pass of each of them depends on the scene only through the SHAPE of the wavefront -- lanes per pass
# = W x H, record slots = lanes x max_depth -- so the tree work of such a pass is measured here on synthetic
# vertices of that shape against the frozen benchmark tree: per bounce ONE sdt_guided call with the emitter-direction pdf
# (every tree query of a path vertex, what PathGuidingCore.bounce issues) on the lanes still alive, then ONE
# sdt_splat_path_data over all record slots.  Path lengths are geometric (75 % of the lanes survive a bounce, the deep
# bounces run a few percent of the lanes through the tile compaction), vertices uniform in the box.
CONFIG_SHAPES = [("cornell-box 256x256 (config 1)", 256, 256, 30), ("torus 512x512 (config 3)", 512, 512, 30),
                 ("veach-ajar 1024x1024 (config 4)", 1024, 1024, 13), ("veach-bidir 3840x2160 (config 5)", 3840, 2160, 7)]


def config_shaped_passes(tree, dev, rank, survive=0.75, reps=3):
    import torch
    out = []
    for name, w, h, md in CONFIG_SHAPES:
        lanes, slots = w * h, w * h * md
        g = torch.Generator(device=dev).manual_seed(1234 + rank)
        length = torch.clamp((torch.log(torch.rand(lanes, device=dev, generator=g)) / np.log(survive)).floor() + 1, max=md).to(torch.int32)
        pos = torch.rand(slots, 3, device=dev, generator=g)
        cdir = torch.rand(slots, 2, device=dev, generator=g)
        wo = torch.nn.functional.normalize(torch.randn(lanes, 3, device=dev, generator=g), dim=1)
        em = torch.nn.functional.normalize(torch.randn(lanes, 3, device=dev, generator=g), dim=1)
        lfin = torch.rand(lanes, 3, device=dev, generator=g) * 4
        tr_, tb_, bs_ = (torch.rand(slots, 3, device=dev, generator=g) for _ in range(3))
        wop = torch.rand(slots, device=dev, generator=g) + 0.05
        depth_of_slot = torch.arange(md, device=dev, dtype=torch.int32).repeat(lanes)
        act = (depth_of_slot < length.repeat_interleave(md)).to(torch.uint8)
        del depth_of_slot
        bp, bv = torch.rand(lanes, device=dev, generator=g), torch.rand(lanes, 3, device=dev, generator=g)
        choose = (torch.rand(lanes, device=dev, generator=g) < 0.5)
        modes = [torch.where(length > b, torch.where(choose, 1, 2), 0).to(torch.uint8) for b in range(md)]     # mode 1 sampled, 2 pdf + mixture, 0 dead
        alive = [int((m != 0).sum().item()) for m in modes]
        em_act = [(m != 0).to(torch.uint8) for m in modes]
        o_dir, o_sp, o_wp, o_w, o_em = (torch.zeros(lanes, 3, device=dev), torch.zeros(lanes, device=dev), torch.zeros(lanes, device=dev),
                                        torch.zeros(lanes, 3, device=dev), torch.ones(lanes, device=dev))
        pos_b = torch.rand(md, lanes, 3, device=dev, generator=g)          # per-bounce vertex positions, contiguous per bounce

        def one_pass():
            for b in range(md):
                if alive[b] == 0:
                    break
                tree.guided(pos_b[b], modes[b], wo=wo, seed=7 + b, bsdf_pdf=bp, bsdf_value=bv, dir_out=o_dir, sdtree_pdf_out=o_sp,
                            wo_pdf_out=o_wp, weight_out=o_w, em_dir=em, em_active=em_act[b], sdtree_pdf_em_out=o_em)
            tree.splat_path_data(md, lfin, tr_, tb_, bs_, pos, cdir, wop, active=act)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        one_pass()
        torch.cuda.synchronize()
        l0 = tree.kernel_launches()
        ev[0].record()
        for _ in range(reps):
            for b in range(md):
                if alive[b] == 0:
                    break
                tree.guided(pos_b[b], modes[b], wo=wo, seed=7 + b, bsdf_pdf=bp, bsdf_value=bv, dir_out=o_dir, sdtree_pdf_out=o_sp,
                            wo_pdf_out=o_wp, weight_out=o_w, em_dir=em, em_active=em_act[b], sdtree_pdf_em_out=o_em)
        ev[1].record()
        for _ in range(reps):
            tree.splat_path_data(md, lfin, tr_, tb_, bs_, pos, cdir, wop, active=act)
        ev[2].record()
        torch.cuda.synchronize()
        ms_b, ms_s = ev[0].elapsed_time(ev[1]) / reps, ev[1].elapsed_time(ev[2]) / reps
        launches_per_pass = (tree.kernel_launches() - l0) / reps
        # the same bounce calls with their argument structs built once (SDTree.prepare_guided: a wavefront loop that keeps
        # its buffers): what is left of the host side is the ctypes call and the launch
        calls = [tree.prepare_guided(pos_b[b], modes[b], wo=wo, seed=7 + b, bsdf_pdf=bp, bsdf_value=bv, dir_out=o_dir, sdtree_pdf_out=o_sp,
                                     wo_pdf_out=o_wp, weight_out=o_w, em_dir=em, em_active=em_act[b], sdtree_pdf_em_out=o_em)
                 for b in range(md) if alive[b] > 0]
        for c in calls:
            c()
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(reps):
            for c in calls:
                c()
        ev[1].record()
        torch.cuda.synchronize()
        ms_bp = ev[0].elapsed_time(ev[1]) / reps
        verts = int(sum(alive))
        out.append({"config": name, "lanes_per_pass": lanes, "max_depth": md, "record_slots": slots, "path_vertices_per_pass": verts,
                    "bounce_calls_ms": ms_b, "bounce_calls_prepared_ms": ms_bp, "splat_path_data_ms": ms_s, "tree_ms_per_pass": ms_b + ms_s,
                    "tree_ms_per_pass_prepared": ms_bp + ms_s, "guided_samples_per_s": verts / ((ms_b + ms_s) * 1e-3),
                    "guided_samples_per_s_prepared": verts / ((ms_bp + ms_s) * 1e-3), "launches_per_pass": launches_per_pass})
        del pos, cdir, tr_, tb_, bs_, wop, act, modes, em_act, pos_b, calls
        torch.cuda.empty_cache()
    return {"passes": out,
            "what": "tree work of ONE training pass (1 spp) at the wavefront shape of each scene config: max_depth sdt_guided calls with the "
                    "emitter-direction pdf on the surviving lanes + one sdt_splat_path_data over lanes x max_depth record slots; synthetic "
                    "vertices, frozen benchmark tree, device-resident buffers, host enqueue included (no synchronisation inside a pass; "
                    "_prepared: the bounce calls through SDTree.prepare_guided, argument structs built once); "
                    "the renderer's own work (Mitsuba ray tracing / BSDFs) is not part of it"}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
