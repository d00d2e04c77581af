"""Synthetic frozen-SD-tree workload of BASELINE.json configs[1].

Records: positions ~ 50 % uniform in the unit box + 8 Gaussian blobs (sigma 0.05, clipped);
directions per spatial octant from a 3-lobe von-Mises-Fisher mixture, kappa in
{10, 1e3, 1e5} (the sharp lobes drive the quadtrees to their depth cap);
radiance ~ LogNormal(0,1); woPdf ~ U(0.05, 2); no NEE radiance.
Everything is seeded numpy so that the GPU arm and the CPU reference arm of bench.py see
the same inputs.
"""
import numpy as np

F = np.float32
BUILD_SEED = 20240611


def _vmf(rng, mu, kappa, n):
    """n unit vectors ~ vMF(mu, kappa) on S^2 (inverse-CDF in the polar angle)"""
    u = rng.random(n)
    w = 1.0 + np.log(u + (1.0 - u) * np.exp(-2.0 * kappa)) / kappa
    w = np.clip(w, -1.0, 1.0)
    phi = 2.0 * np.pi * rng.random(n)
    s = np.sqrt(np.maximum(0.0, 1.0 - w * w))
    a = np.where(np.abs(mu[:, 2:3]) < 0.9, np.array([[0.0, 0.0, 1.0]]), np.array([[1.0, 0.0, 0.0]]))
    t1 = np.cross(mu, a)
    t1 /= np.linalg.norm(t1, axis=1, keepdims=True)
    t2 = np.cross(mu, t1)
    return (s * np.cos(phi))[:, None] * t1 + (s * np.sin(phi))[:, None] * t2 + w[:, None] * mu


def dir_to_canonical(d):
    """(n,3) unit vectors -> (n,2) in [0,1]^2, the map of src/common.py:132-158 (data
    generation only; the kernels' own map is in csrc/sdt_core.h)"""
    phi = np.arctan2(d[:, 1], d[:, 0])
    phi = np.where(phi < 0, phi + 2.0 * np.pi, phi)
    x = np.clip(phi / (2.0 * np.pi), 0.0, 1.0)
    y = np.clip((np.clip(d[:, 2], -1.0, 1.0) + 1.0) / 2.0, 0.0, 1.0)
    return np.stack([x, y], 1).astype(F)


class Scene:
    """the fixed part of the distribution: blob centres and per-octant lobes"""

    def __init__(self, seed=BUILD_SEED):
        rng = np.random.default_rng(seed)
        self.blobs = rng.random((8, 3))
        lobes = rng.standard_normal((8, 3, 3))
        self.lobes = lobes / np.linalg.norm(lobes, axis=2, keepdims=True)
        self.kappa = np.array([10.0, 1e3, 1e5])
        self.weights = np.array([0.5, 0.3, 0.2])

    def positions(self, rng, n):
        pos = rng.random((n, 3))
        blob = rng.random(n) < 0.5
        k = rng.integers(0, 8, n)
        g = self.blobs[k] + 0.05 * rng.standard_normal((n, 3))
        pos = np.where(blob[:, None], np.clip(g, 0.0, 1.0), pos)
        return pos.astype(F)

    def records(self, seed, n):
        """-> dict(position (n,3), direction (n,2) canonical, radiance (n,), wo_pdf (n,))"""
        rng = np.random.default_rng(seed)
        pos = self.positions(rng, n)
        octant = (pos[:, 0] >= 0.5).astype(int) + 2 * (pos[:, 1] >= 0.5).astype(int) + 4 * (pos[:, 2] >= 0.5).astype(int)
        lobe = rng.choice(3, n, p=self.weights)
        d = _vmf(rng, self.lobes[octant, lobe], self.kappa[lobe], n)
        return dict(position=pos, direction=dir_to_canonical(d),
                    radiance=rng.lognormal(0.0, 1.0, n).astype(F),
                    wo_pdf=(0.05 + 1.95 * rng.random(n)).astype(F))


def uniform_sphere(seed, n):
    rng = np.random.default_rng(seed)
    d = rng.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return d.astype(F)


def uniform_box(seed, n):
    return np.random.default_rng(seed).random((n, 3)).astype(F)


# tree-build schedule of the microbench: K iterations, record count doubling like the
# reference's spp (main.py:170), spatial threshold c*sqrt(2^k) (src/kdtree.py:327-330) with c
# scaled so that the frozen tree has ~4096 spatial leaves (BASELINE.json configs[1])
BUILD_ITERS = 6
BUILD_N0 = 1 << 15
BUILD_C = 68.0


def build_schedule(iters=BUILD_ITERS, n0=BUILD_N0, c=BUILD_C):
    return [dict(iteration=k, n=n0 << k, max_leaf_size=float(np.float32(c * np.sqrt(2.0 ** k))), seed=BUILD_SEED + 1 + k)
            for k in range(iters)]


def build_tree(tree, scene=None, schedule=None, to_dev=None, rank=0, world=1, allreduce=None):
    """trains `tree` (an SDTree) on the synthetic records: splat + device-side refine per
    iteration.  With world > 1 every rank splats its slice of the iteration's records and
    `allreduce(records)` combines the statistics before the (deterministic) refine; records = the iteration's records
    over all ranks (SDTree.allreduce's records_all_ranks: bounds the spatial split rounds)."""
    scene = scene or Scene()
    to_dev = to_dev or (lambda x: x)
    for it in (schedule or build_schedule()):
        rec = scene.records(it['seed'], it['n'])
        sl = slice(rank * it['n'] // world, (rank + 1) * it['n'] // world)
        tree.splat_records(to_dev(rec['position'][sl]), to_dev(rec['direction'][sl]), to_dev(rec['radiance'][sl]), to_dev(rec['wo_pdf'][sl]))
        if allreduce is not None:
            allreduce(it['n'])
        tree.set_max_leaf_size(it['max_leaf_size'])
        tree.refine()
    return tree
