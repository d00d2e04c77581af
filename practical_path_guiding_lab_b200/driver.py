"""Parameterised equivalent of the reference's training/render driver (/root/reference/main.py:
97-416): the same iteration doubling (4, 8, 16, ... spp), 1 spp per pass
while training and `batch_spp` per pass in the final iteration, the same stop rule (variance
estimate after `stable_variance_spp_threshold`, hard stop at 1000 cumulative spp), the same
two-iteration image blending, refine after every training iteration, and the same outputs
(image, .npz tree, .obj boxes, CSV records) -- with scene, resolution, budget, seed as arguments.

`renderer` is anything with the integrator-side interface the reference script uses:
render(spp, seed) -> (H,W,3) image, setIteration, resetVarianceCounter, computeVariance,
computeMSE, refineAndPrepareSDTreeForNextIteration, saveSDTreeToFile, saveSDTreeOBJ
(`cornell.CornellBox` here; the Mitsuba plugin when Mitsuba is installed).
"""
import csv
import math
import os
import time


def possible_cumm_spps(budget_spp):
    """main.py:105-118"""
    cumm, k, out = 0, 0, []
    while cumm < budget_spp:
        cumm += 2 ** (k + 2)
        out.append(cumm)
        k += 1
    return out


def shard_plan(passes, world, mode="auto"):
    """-> (tiles_per_pass, groups): how `world` ranks share the `passes` render passes of one iteration.
    Ranks form `groups` groups of `tiles_per_pass` ranks; pass p goes to group p % groups, and inside the group rank
    r renders tile r % tiles_per_pass of the film.  auto: whole passes per rank while there are at least as many passes
    as ranks (full-width wavefronts: the per-bounce host cost of a wavefront does not shrink with its width), film tiles
    only in the early iterations that have fewer passes than ranks (4 and 8, main.py:170) so that no GPU idles."""
    if mode == "passes" or world == 1:
        return 1, world
    if mode == "tiles":
        return world, 1
    t = 1
    while passes * t < world and world % (2 * t) == 0:
        t *= 2
    return t, world // t


class SingleRank:
    rank, world = 0, 1
    mode = "passes"

    def sum_image(self, x):
        return x

    def sync_iteration(self, renderer, passes=None):
        pass


class TorchDistRanks:
    """One process per GPU.  The passes of an iteration (seed = seed0 + cumm_spp exactly as in the
    sequential loop, main.py:208-218) are shared by image tiles and sample batches (shard_plan): mode "passes" deals whole
    passes round-robin, "tiles" gives every rank its band of the film in every pass, "auto" (default) uses whole passes
    and splits the film only while an iteration has fewer passes than there are ranks.  Lanes keep globally unique RNG
    keys.  At the end of the iteration the statistics of `current` are combined with ONE sdt_allreduce (NCCL), the
    variance counters and the image with torch.distributed, and every rank runs the same deterministic refine ->
    bit-identical trees without a broadcast."""

    def __init__(self, mode="auto", tiles=None):
        import torch.distributed as dist
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.mode = "tiles" if tiles else ("passes" if tiles is not None else mode)

    def sum_image(self, x):
        self.dist.all_reduce(x)
        return x

    def sync_iteration(self, renderer, passes=None):
        renderer.allreduce_statistics(self.dist, passes)


def train_and_render(renderer, budget_spp, seed=0, batch_spp=4, stable_variance_spp_threshold=256,
                     ground_truth=None, out_dir=None, scene_name="scene", log=None, on_iteration=None, ranks=None,
                     record_in_iteration=False):
    """-> dict(image, records=[per-iteration dicts], iterations=[(iteration, spp, refined)])"""
    ranks = ranks or SingleRank()
    log = log or (lambda *a: None)
    cumm_spp = cumm_spp_prev = 0
    image_spp = 0
    remaining = budget_spp
    is_final = False
    is_train = True
    is_clear = True
    it = 0
    variance_prev = 0.0
    cumm_time = 0.0
    prev_iter_image = None
    image = None
    records, schedule, in_iter = [], [], []
    if out_dir:
        for sub in ("image", "tree-data", "obj", "performance"):
            os.makedirs(os.path.join(out_dir, sub), exist_ok=True)
    while remaining > 0:
        t0 = time.perf_counter()
        if is_clear:
            renderer.resetVarianceCounter()
            image_spp = 0
        curr = None
        if not is_final:
            iter_spp = 2 ** (it + 2)                                   # main.py:170
            if iter_spp == remaining:
                is_final = True
        else:
            iter_spp = remaining
        renderer.setIteration(it, is_final)
        spp_per_pass = batch_spp if is_final else 1                   # main.py:192-199
        passes = math.ceil(iter_spp / spp_per_pass)
        tiles_per_pass, groups = shard_plan(passes, ranks.world, ranks.mode)
        if ranks.world > 1 and hasattr(renderer, "set_tile"):
            renderer.set_tile(ranks.rank % tiles_per_pass, tiles_per_pass)
        elif tiles_per_pass > 1:
            raise RuntimeError("this renderer cannot render film tiles: use shard mode 'passes'")
        done = 0
        for p_i in range(passes):
            s = min(spp_per_pass, iter_spp - done)
            if p_i % groups == ranks.rank // tiles_per_pass:
                one = renderer.render(s, seed + cumm_spp)              # main.py:218
                w = one * float(s / iter_spp)
                curr = w if curr is None else curr + w
            image_spp += s
            done += s
            cumm_spp += s
            if record_in_iteration and ranks.world == 1:               # main.py:245-265 (isRecordPerformanceInIteration)
                in_iter.append(dict(time=(time.perf_counter() - t0) + cumm_time, spp=image_spp, cumm_spp=cumm_spp, iteration=it,
                                    variance=renderer.computeVariance(image_spp),
                                    variance_groundTruth=renderer.computeVariance(image_spp, ground_truth) if ground_truth is not None else 0,
                                    mse_groundTruth=renderer.computeMSE(image_spp, ground_truth) if ground_truth is not None else 0))
        if ranks.world > 1:
            if curr is None:
                curr = renderer.zero_image()
            curr = ranks.sum_image(curr)
            ranks.sync_iteration(renderer, passes)                     # one exchange per iteration
        if is_final and not is_train and prev_iter_image is not None:   # main.py:287-291
            image = (curr * iter_spp + prev_iter_image * (image_spp - iter_spp)) / image_spp
        else:
            image = curr
        variance = renderer.computeVariance(image_spp)
        var_gt = renderer.computeVariance(image_spp, ground_truth) if ground_truth is not None else None
        mse_gt = renderer.computeMSE(image_spp, ground_truth) if ground_truth is not None else None
        elapsed = (time.perf_counter() - t0) + cumm_time
        variance_current = (variance * image_spp) / (budget_spp - cumm_spp_prev)    # main.py:325-326
        was_final, trained_this = is_final, is_train
        # next-iteration conditions, main.py:335-377
        next_spp = 2 ** (it + 3)
        remaining = budget_spp - cumm_spp
        stop = (cumm_spp > stable_variance_spp_threshold) and (variance_current > variance_prev)
        if cumm_spp >= 1000:
            stop = True
        if next_spp < remaining:
            if stop:
                is_final, is_train, is_clear = True, False, False
        elif next_spp == remaining:
            is_final = True
            if stop:
                is_train, is_clear = False, False
        else:
            is_final, is_train, is_clear = True, False, False
        refined = False
        if is_train:
            renderer.refineAndPrepareSDTreeForNextIteration()          # main.py:382-383
            refined = True
        prev_iter_image = image
        cumm_time += time.perf_counter() - t0
        rec = dict(iteration=it, iter_spp=iter_spp, spp=image_spp, cumm_spp=cumm_spp, time=elapsed, variance=variance,
                   variance_groundTruth=var_gt, mse_groundTruth=mse_gt, variance_estimated_final=variance_current,
                   isFinalIter=was_final, refined=refined)
        records.append(rec)
        schedule.append((it, iter_spp, refined))
        log(f"iteration {it}: spp {iter_spp}, cumm {cumm_spp}, final {was_final}, refined {refined}, variance {variance:.5g}"
            + (f", mse {mse_gt:.5g}" if mse_gt is not None else ""))
        if out_dir and ranks.rank == 0:
            renderer.saveSDTreeToFile(os.path.join(out_dir, "tree-data", f"{scene_name}_iter-{it}.npz"))
            renderer.saveSDTreeOBJ(os.path.join(out_dir, "obj", f"{scene_name}_iter-{it}.obj"))
            save_image(os.path.join(out_dir, "image", f"{scene_name}_iter-{it}_spp-{image_spp}_cumm_spp-{cumm_spp}"), image)
        if on_iteration:
            on_iteration(rec, image)
        variance_prev = variance_current
        it += 1
        cumm_spp_prev = cumm_spp
    if out_dir and ranks.rank == 0:
        # PerformanceData.saveToFile (src/common.py:86-97): six columns, the unused one of variance / mse stays 0;
        # file names of main.py:425-429
        files = [(records, "variance", "variance_endIter.csv"), (records, "variance_groundTruth", "variance_groundTruth_endIter.csv"),
                 (records, "mse_groundTruth", "mse_groundTruth_endIter.csv"), (records, "variance_estimated_final", "variance_estimated_final.csv")]
        if in_iter:                                                    # main.py:420-423
            files += [(in_iter, "variance", "variance_inIter.csv"), (in_iter, "variance_groundTruth", "variance_groundTruth_inIter.csv"),
                      (in_iter, "mse_groundTruth", "mse_groundTruth_inIter.csv")]
        for rows, key, name in files:
            with open(os.path.join(out_dir, "performance", name), "w", newline="") as f:
                w = csv.writer(f)
                w.writerow(["time", "spp", "cumm_spp", "iteration", "variance", "mse"])
                for r in rows:
                    v = 0 if r[key] is None else r[key]
                    w.writerow([r["time"], r["spp"], r["cumm_spp"], r["iteration"]] + ([0, v] if key.startswith("mse") else [v, 0]))
    return dict(image=image, records=records, iterations=schedule)


def render_fixed_trees(renderer, tree_files, iter_spp=1024, batch_spp=4, seed=0, ground_truth=None, log=None):
    """The fixed-tree study of the reference's repeat_high_spp_renderer.py:69-160: for iteration k the
    tree saved after iteration k-1 is loaded (`loadSDTreeFromFile`; iteration 0 renders unguided), the
    integrator is put in final mode (`setIteration(k, True)`: nothing is recorded) and `iter_spp`
    samples are rendered in `batch_spp` passes; variance / MSE are recorded per iteration.
    tree_files[k-1] = npz of the tree to use in iteration k."""
    log = log or (lambda *a: None)
    records = []
    cumm_spp = 0
    for it in range(len(tree_files) + 1):
        renderer.resetVarianceCounter()
        renderer.setIteration(it, True)
        if it > 0:
            renderer.loadSDTreeFromFile(tree_files[it - 1])
        t0 = time.perf_counter()
        image, done = None, 0
        for _ in range(math.ceil(iter_spp / batch_spp)):
            s = min(batch_spp, iter_spp - done)
            one = renderer.render(s, seed + cumm_spp) * float(s / iter_spp)
            image = one if image is None else image + one
            done += s
            cumm_spp += s
        rec = dict(iteration=it, spp=iter_spp, time=time.perf_counter() - t0, variance=renderer.computeVariance(iter_spp),
                   variance_groundTruth=renderer.computeVariance(iter_spp, ground_truth) if ground_truth is not None else None,
                   mse_groundTruth=renderer.computeMSE(iter_spp, ground_truth) if ground_truth is not None else None)
        records.append(rec)
        log(f"fixed tree, iteration {it}: {iter_spp} spp, variance {rec['variance']:.5g}")
    return records


def save_image(stem, image):
    import numpy as np
    a = image.detach().cpu().numpy() if hasattr(image, "detach") else np.asarray(image)
    np.save(stem + ".npy", a.astype(np.float32))
    try:
        os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
        import cv2
        cv2.imwrite(stem + ".png", (np.clip(a[..., ::-1], 0, 1) ** (1 / 2.2) * 255).astype(np.uint8))
    except Exception:
        pass


def main(argv=None):
    """python -m practical_path_guiding_lab_b200.driver --scene cornell-box --res 256 --budget 64"""
    import argparse
    import json
    import numpy as np
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="cornell-box", choices=["cornell-box"])
    ap.add_argument("--res", type=int, default=256)
    ap.add_argument("--budget", type=int, default=64)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--max-depth", type=int, default=30)
    ap.add_argument("--out", default=None)
    ap.add_argument("--ground-truth", default=None, help=".npy (H,W,3) linear RGB")
    ap.add_argument("--no-guiding", action="store_true", help="BSDF-only baseline: never refine (tree stays a single leaf)")
    ap.add_argument("--shard", default="auto", choices=["auto", "tiles", "passes"],
                    help="multi-GPU work split: whole passes, film tiles within every pass, or (auto) tiles only while an iteration has fewer passes than ranks")
    ap.add_argument("--record-in-iteration", action="store_true", help="variance / MSE after every pass (main.py isRecordPerformanceInIteration; one GPU)")
    a = ap.parse_args(argv)
    from .cornell import CornellBox
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    ranks = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        ranks = TorchDistRanks(mode=a.shard)
    r = CornellBox(a.res, a.res, max_depth=a.max_depth, device=f"cuda:{local}")
    r.setup()
    if ranks is not None:
        r.comm_init(ranks.dist)
    gt = None
    gt_native = None
    if a.ground_truth:
        g = np.load(a.ground_truth).astype(np.float32)
        gt_native = torch.from_numpy(np.ascontiguousarray(g)).cuda()
        if g.shape[0] != a.res:                      # box-resample the fixture to the render resolution
            g = torch.nn.functional.interpolate(torch.from_numpy(g).permute(2, 0, 1)[None], size=(a.res, a.res), mode="area")[0].permute(1, 2, 0).numpy()
        gt = torch.from_numpy(np.ascontiguousarray(g)).cuda()
    if a.no_guiding:
        r.refineAndPrepareSDTreeForNextIteration = lambda: None
    t0 = time.perf_counter()
    rank0 = ranks is None or ranks.rank == 0
    res = train_and_render(r, a.budget, seed=a.seed, ground_truth=gt, out_dir=a.out, scene_name=a.scene,
                           log=print if rank0 else None, ranks=ranks, record_in_iteration=a.record_in_iteration)
    torch.cuda.synchronize()
    sizes = r.core.tree.sizes()
    if ranks is not None:                            # every rank must hold the same tree
        sig = torch.tensor([sizes["n_kd"], sizes["n_quad"], sizes["n_roots"]], device="cuda", dtype=torch.int64)
        lo, hi = sig.clone(), sig.clone()
        ranks.dist.all_reduce(lo, op=ranks.dist.ReduceOp.MIN)
        ranks.dist.all_reduce(hi, op=ranks.dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), "trees differ between ranks"
    extra = {}
    if gt_native is not None and gt_native.shape[0] < a.res and a.res % gt_native.shape[0] == 0:
        # a ground truth of lower resolution than the render was up-sampled above: every pixel of a block is compared with
        # the block's mean, so mse_groundTruth has a floor (edges, texture inside a block) that no sample count removes.
        # The same image box-filtered DOWN to the ground truth's own resolution has no such floor:
        k = a.res // gt_native.shape[0]
        img = res["image"].view(gt_native.shape[0], k, gt_native.shape[1], k, 3).mean(dim=(1, 3))
        lum = torch.tensor([0.212671, 0.715160, 0.072169], device=img.device)
        extra["mse_at_ground_truth_resolution"] = float((((img - gt_native) ** 2) * lum).sum(-1).clamp_max(10000).mean())
        extra["ground_truth_resolution"] = int(gt_native.shape[0])
    if rank0:
        print(json.dumps(dict(scene=a.scene, res=a.res, budget=a.budget, n_gpus=world, shard=(a.shard if world > 1 else None),
                              seconds=time.perf_counter() - t0, iterations=res["iterations"], final=res["records"][-1], tree=sizes, **extra)))
    if ranks is not None:
        ranks.dist.destroy_process_group()


if __name__ == "__main__":
    main()
