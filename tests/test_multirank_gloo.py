"""N>1 path on CPU: world_size-2 `gloo` run of the sharded training iteration -- every rank
splats its slice of the records into its replica, the LEAF statistics are all-reduced
(here through sdt_stat_buffers + torch.distributed, on the GPU through sdt_allreduce/NCCL),
the deterministic refine then yields the same tree on every rank, equal to the tree a
single rank builds from all the records.  Uses the host emulation of the kernels."""
import ctypes
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _worker(rank, world, port, lib, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import sdt_cases as cases
    from practical_path_guiding_lab_b200 import SDTree
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    t = SDTree(lib_path=lib, store_nee=False, kd_capacity=1 << 12, quad_capacity=1 << 17)
    for it in range(3):
        rec = cases.dyadic_records(16000, 300 + it, ((0.3, 0.7, 0.02), (0.8, 0.2, 0.005)))
        n = rec.position.shape[0]
        sl = slice(rank * n // world, (rank + 1) * n // world)
        t.splat_records(rec.position[sl], rec.direction[sl], rec.radiance[sl], rec.woPdf[sl])
        qp, nq, kp, nk = t.stat_buffers()
        q = np.ctypeslib.as_array(ctypes.cast(qp, ctypes.POINTER(ctypes.c_float)), shape=(nq,))
        k = np.ctypeslib.as_array(ctypes.cast(kp, ctypes.POINTER(ctypes.c_float)), shape=(nk,))
        buf = torch.from_numpy(np.concatenate([q, k]))
        dist.all_reduce(buf)                                  # one exchange per training iteration
        q[:] = buf[:nq].numpy()
        k[:] = buf[nq:].numpy()
        t.set_max_leaf_size(500)
        t.refine()
    d = t.download(0)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **d)
    dist.destroy_process_group()


def test_two_rank_training_matches_single_rank(tmp_path):
    import torch.multiprocessing as mp
    from hostemu.build_hostemu import build as build_hostemu
    import sdt_cases as cases
    from practical_path_guiding_lab_b200 import SDTree
    lib = build_hostemu()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, lib, str(tmp_path)), nprocs=2, join=True)
    t = SDTree(lib_path=lib, store_nee=False, kd_capacity=1 << 12, quad_capacity=1 << 17)
    for it in range(3):
        rec = cases.dyadic_records(16000, 300 + it, ((0.3, 0.7, 0.02), (0.8, 0.2, 0.005)))
        t.splat_records(rec.position, rec.direction, rec.radiance, rec.woPdf)
        t.set_max_leaf_size(500)
        t.refine()
    one = t.download(0)
    assert one['kdtree_depth'].shape[0] > 31
    for r in range(2):
        d = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        for k in one:
            a, b = np.asarray(one[k]), np.asarray(d[k])
            assert a.shape == b.shape and np.array_equal(a, b), (r, k)


def _tile_worker(rank, world, port, lib, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from practical_path_guiding_lab_b200 import driver
    from practical_path_guiding_lab_b200.cornell import CornellBox
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ranks = driver.TorchDistRanks(mode="tiles")
    r = CornellBox(16, 16, max_depth=5, device="cpu", lib_path=lib, kd_capacity=1 << 12, quad_capacity=1 << 16)
    r.setup(sdTreeMaxDepth=8, quadTreeMaxDepth=6)
    res = driver.train_and_render(r, 28, seed=3, ranks=ranks)
    assert r.core.numRays == 128                      # half of the film per rank
    np.savez(os.path.join(out_dir, f"tile_rank{rank}.npz"), image=res["image"].numpy(), **r.core.tree.download(0))
    dist.destroy_process_group()


def test_tile_sharded_driver_two_ranks(tmp_path):
    """driver.TorchDistRanks(tiles=True): every rank renders its band of the film in every pass, one statistics exchange
    per iteration -> the same tree on both ranks, and an image whose two halves both carry light"""
    import torch.multiprocessing as mp
    from hostemu.build_hostemu import build as build_hostemu
    lib = build_hostemu()
    port = 31500 + os.getpid() % 2000
    mp.spawn(_tile_worker, args=(2, port, lib, str(tmp_path)), nprocs=2, join=True)
    a = np.load(os.path.join(str(tmp_path), "tile_rank0.npz"))
    b = np.load(os.path.join(str(tmp_path), "tile_rank1.npz"))
    for k in a.files:
        assert np.array_equal(a[k], b[k], equal_nan=True), k      # trees AND the all-reduced image
    img = a["image"]
    assert img.shape == (16, 16, 3) and np.isfinite(img).all()
    assert img[:8].sum() > 0 and img[8:].sum() > 0
    assert a["kdtree_depth"].shape[0] >= 1


def test_shard_plan_covers_every_pass_and_tile_once():
    """driver.shard_plan: whole passes while an iteration has at least as many passes as ranks, film tiles below that"""
    sys.path.insert(0, ROOT)
    from practical_path_guiding_lab_b200 import driver
    assert driver.shard_plan(4, 8) == (2, 4) and driver.shard_plan(8, 8) == (1, 8) and driver.shard_plan(256, 8) == (1, 8)
    assert driver.shard_plan(4, 2) == (1, 2) and driver.shard_plan(1, 8) == (8, 1) and driver.shard_plan(4, 1) == (1, 1)
    assert driver.shard_plan(4, 8, "tiles") == (8, 1) and driver.shard_plan(4, 8, "passes") == (1, 8)
    for world in (1, 2, 4, 8):
        for passes in (1, 2, 4, 8, 16, 128):
            for mode in ("auto", "tiles", "passes"):
                t, g = driver.shard_plan(passes, world, mode)
                assert t * g == world
                seen = {}
                for rank in range(world):
                    for p in range(passes):
                        if p % g == rank // t:
                            seen[(p, rank % t)] = seen.get((p, rank % t), 0) + 1
                if mode != "passes" or passes >= 1:
                    assert all(v == 1 for v in seen.values())
                    assert len(seen) == passes * t, (world, passes, mode)
