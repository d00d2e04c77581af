"""Parity cases shared by the CPU suite (host emulation of the kernels' index logic,
tests/test_hostemu_parity.py) and the GPU suite (libsdtree.so on the B200,
tests/test_gpu_parity.py).  Every case drives the library through its C ABI (via the
ctypes wrapper) and holds the result against the oracle on the same seeded inputs.

`ctx.make(**cfg)` builds an SDTree on the library under test; `ctx.dev(x)` moves a numpy
array to where the test wants the call's buffers (numpy = host-pointer path, torch CUDA =
device-pointer path); `ctx.host(x)` brings a result back as numpy.
"""
import numpy as np

from oracle import sdtree_oracle as so
from oracle import drjit_math as dm

F = np.float32
U = np.uint32


def beq(a, b):
    """bit-exact fp32 equality; NaNs match NaNs (their payload/sign differs between x86 and sm_100)"""
    a = np.ascontiguousarray(a, F)
    b = np.ascontiguousarray(b, F)
    if a.shape != b.shape:
        return False
    return bool(np.all((a.view(U) == b.view(U)) | (np.isnan(a) & np.isnan(b))))


class Ctx:
    def __init__(self, make, dev=None, host=None):
        self.make = make
        self.dev = dev or (lambda x: x)
        self.host = host or (lambda x: np.asarray(x))

    def u32(self, x):
        return self.host(x).view(U) if self.host(x).dtype != U else self.host(x)


# fl(NEE_GREEN_UNIT * 0.715160f) == 1.0f exactly, so mi.luminance((0, k * NEE_GREEN_UNIT, 0)) == k for every power of
# two k: NEE energy that is NON-ZERO and still dyadic, i.e. topology after NEE deposits stays bit-reproducible
NEE_GREEN_UNIT = np.array([1068694302], np.uint32).view(np.float32)[0]
assert np.float32(NEE_GREEN_UNIT * np.float32(0.715160)) == np.float32(1.0)


def dyadic_nee(rng, n):
    nee = np.zeros((n, 3), F)
    nee[:, 1] = rng.choice(np.array([0, 0, 0.25, 1, 2, 8], F), n) * NEE_GREEN_UNIT
    return nee


# ------------------------------------------------------------------------------ data
def dyadic_records(n, seed, lobes=((0.3, 0.7, 0.02),), box=1.0, nee=False):
    """records whose radiance/woPdf are multiples of 1/16 (<= 8): fp32 sums of up to
    ~1e5 of them are exact in any order, so splat results are bit-reproducible"""
    rng = np.random.default_rng(seed)
    pos = (rng.random((n, 3)) ** 1.5 * box).astype(F)
    d = rng.random((n, 2)).astype(F)
    k = rng.integers(0, len(lobes) + 1, n)
    for j, (cx, cy, s) in enumerate(lobes):
        m = k == j
        d[m] = np.clip(np.stack([cx + s * rng.standard_normal(m.sum()), cy + s * rng.standard_normal(m.sum())], 1), 0, 1).astype(F)
    radiance = (rng.integers(0, 17, n) / 8.0).astype(F)
    wo = rng.choice(np.array([0.25, 0.5, 1.0, 2.0], F), n).astype(F)
    rec = so.SurfaceInteractionRecord(pos, d, radiance, wo)
    if nee:
        rec.radiance_nee = dyadic_nee(rng, n)
        rec.direction_nee = rng.random((n, 2)).astype(F)
        m = rng.random(n) < 0.5                                  # an NEE lobe of its own: it shapes the quadtrees
        rec.direction_nee[m] = np.clip(np.stack([0.15 + 0.01 * rng.standard_normal(m.sum()),
                                                 0.85 + 0.01 * rng.standard_normal(m.sum())], 1), 0, 1).astype(F)
    return rec


def oracle_pair(bbox_min=(0, 0, 0), bbox_max=(1, 1, 1), kd_max_depth=20, quad_max_depth=20, store_nee=False):
    cur = so.KDTree(maxDepth=kd_max_depth)
    cur.setup(bbox_min, bbox_max)
    cur.quadTree.maxDepth = quad_max_depth
    cur.quadTree.isStoreNEERadiance = store_nee
    prev = so.KDTree(maxDepth=kd_max_depth)
    prev.copyFrom(cur)
    return cur, prev


def oracle_refine(cur, prev, max_leaf_size, kd=True, quad=True):
    """refineAndPrepareSDTreeForNextIteration with an explicit KDTree.maxLeafSize"""
    cur.maxLeafSize = max_leaf_size
    if kd:
        cur.refine()
    if quad:
        cur.setQuadTreeRefinementThreshold()
        cur.refineAllQuadTree()
    cur.cleanUnusedQuadTree()
    prev.copyFrom(cur)
    cur.resetTreeVertCount()
    cur.resetAllQuadTreeIrradiance()


def splat(tree, ctx, rec, active=None):
    tree.splat_records(ctx.dev(rec.position), ctx.dev(rec.direction), ctx.dev(rec.radiance), ctx.dev(rec.woPdf),
                       ctx.dev(rec.radiance_nee), ctx.dev(rec.direction_nee),
                       None if active is None else ctx.dev(active.astype(np.uint8)))


def assert_tree_equal(got, want_tree, exact_energy=True, rtol=1e-4):
    """got: dict in the npz schema from SDTree.download; want_tree: oracle KDTree"""
    want = want_tree.to_arrays()
    for k in so.KDTree.NPZ_KEYS:
        g, w = np.asarray(got[k]), np.asarray(want[k])
        if k == 'kdtree_maxLeafSize':
            assert np.float32(g) == np.float32(w), k
            continue
        assert g.shape == w.shape, (k, g.shape, w.shape)
        if g.dtype.kind == 'f' and not exact_energy and k in ('quadtree_irradiance', 'quadtree_refinementThreshold'):
            np.testing.assert_allclose(g, w, rtol=rtol, atol=1e-30, err_msg=k)
        elif g.dtype.kind == 'f':
            assert (beq(g, w) if g.dtype == F else np.array_equal(g, w)), k
        else:
            assert np.array_equal(g, w), k


def train(ctx, iters=4, n=20000, max_leaf=600, kd_max_depth=20, quad_max_depth=20, store_nee=False, box=1.0, caps=None, tuning=()):
    """identical splat+refine loop on the library and on the oracle; returns both"""
    caps = caps or dict(kd_capacity=1 << 14, quad_capacity=1 << 18)
    t = ctx.make(bbox_min=(0, 0, 0), bbox_max=(box, box, box), kd_max_depth=kd_max_depth, quad_max_depth=quad_max_depth,
                 store_nee=store_nee, **caps)
    for k, v in tuning:
        t.set_tuning(k, v)
    cur, prev = oracle_pair((0, 0, 0), (box, box, box), kd_max_depth, quad_max_depth, store_nee)
    lobes = [((0.3, 0.7, 0.02),), ((0.3, 0.7, 0.004), (0.8, 0.2, 0.05)), ((0.8, 0.2, 0.01),), ((0.55, 0.5, 0.001), (0.1, 0.1, 0.1))]
    for it in range(iters):
        rec = dyadic_records(n, 100 + it, lobes[it % len(lobes)], box=box, nee=store_nee)
        splat(t, ctx, rec)
        cur.addDataPropagate(rec)
        t.set_max_leaf_size(max_leaf)
        t.refine()
        oracle_refine(cur, prev, max_leaf)
    return t, cur, prev


# ------------------------------------------------------------------------------ cases
def case_initial_tree(ctx):
    t = ctx.make(bbox_min=(0, 0, 0), bbox_max=(100, 100, 100), kd_max_depth=10, quad_max_depth=20, store_nee=False,
                 kd_capacity=64, quad_capacity=256)
    cur, prev = oracle_pair((0, 0, 0), (100, 100, 100), 10, 20, False)
    cur.maxLeafSize = 1
    assert_tree_equal(t.download(0), cur)
    s = t.sizes()
    assert (s['n_kd'], s['n_quad'], s['n_roots'], s['n_levels'], s['error']) == (1, 1, 1, 1, 0)
    pos = np.array([[1, 2, 3], [101, 0, 0], [np.nan, 0, 0]], F)
    d, p = t.sample(ctx.dev(pos), seed=1)
    np.testing.assert_array_equal(ctx.host(p), np.full(3, dm.INV_FOUR_PI, F))


def case_golden_upload_download(ctx):
    """hand-derived golden trees: upload in the reference schema, queries, download"""
    k = so.KDTree()
    k.setup([0, 0, 0], [100, 100, 100])
    k.split(k.getAllLeafNodeIndex())
    k.split(k.getAllLeafNodeIndex())
    # give every tree a different shape: non-canonical node order on purpose
    q = k.quadTree.quadTreeNode
    for r in (3, 1):
        q.split(q.getAllLeafNodeIndex(np.array([r], U)))
    q.split(q.getAllLeafNodeIndex(np.array([1], U))[1:3])
    rng = np.random.default_rng(5)
    q.irradiance[:] = rng.integers(1, 9, q.getWidth()).astype(F)
    t = ctx.make(bbox_min=(0, 0, 0), bbox_max=(100, 100, 100), kd_max_depth=10, quad_max_depth=20, store_nee=False,
                 kd_capacity=64, quad_capacity=256)
    t.upload(k.to_arrays())
    k.cleanUnusedQuadTree()                       # canonical layout = what download returns
    assert_tree_equal(t.download(0), k)
    pos = np.array([[75, 25, 25], [50, 50, 1], [50, 49, 1], [49, 50, 1], [0, 0, 0], [100, 100, 100],
                    [101, 1, 1], [np.nan, 1, 1], [-1e-3, 5, 5]], F)
    leaf, root = t.locate(ctx.dev(pos))
    assert ctx.host(leaf).view(U).tolist() == [5, 6, 5, 4, 3, 6, 0, 0, 0]
    assert ctx.host(root).view(U).tolist() == [1, 3, 1, 2, 0, 3, 0, 0, 0]
    act = np.zeros(9, np.uint8)
    leaf, root = t.locate(ctx.dev(pos), ctx.dev(act))
    assert ctx.host(leaf).view(U).tolist() == [0] * 9


def check_queries(ctx, t, prev, n=4096, seed=7, box=1.0, explicit=True):
    rng = np.random.default_rng(seed)
    pos = (rng.random((n, 3)) * box * 1.02 - 0.01 * box).astype(F)      # a few lanes outside the box
    pos[:8] = np.array([[0.5, 0.5, 0.5], [0.25, 0.5, 0.75], [0, 0, 0], [1, 1, 1], [0.5, 0.25, 0.125],
                        [np.nan, 0.5, 0.5], [0.75, 0.75, 0.75], [0.5, 0.5, 0.0]], F) * box
    active = rng.random(n) < 0.9
    # locate
    leaf, root = t.locate(ctx.dev(pos), ctx.dev(active.astype(np.uint8)))
    o_leaf = prev.getLeafNodeIndex(pos, active)
    o_root = so.gather(prev.kdTreeNode.quadTreeRootIndex, o_leaf, active)
    assert np.array_equal(ctx.host(leaf).view(U), o_leaf)
    assert np.array_equal(ctx.host(root).view(U), o_root)
    # sample
    depth = prev.quadTree.maxDepth
    if explicit:
        u = rng.random((n, 3 * (depth + 2))).astype(F)
        u[:16, 2::3] = np.array([0.0, 0.25, 0.5, 0.75] * 4, F)[:, None]     # selection uniforms on bin edges
        sampler = so.ExplicitSampler(u=u)
        d, p, dbg = t.sample(ctx.dev(pos), ctx.dev(active.astype(np.uint8)), u=ctx.dev(u), debug=True)
    else:
        sampler = so.ExplicitSampler(seed=1234, n=n, lane_offset=17)
        d, p, dbg = t.sample(ctx.dev(pos), ctx.dev(active.astype(np.uint8)), seed=1234, lane_offset=17, debug=True)
    od, op, odbg = prev.sample(pos, sampler, active, return_debug=True)
    dbg = ctx.host(dbg).view(U)
    a = active
    assert np.array_equal(dbg[a, 0], odbg['leaf'][a])
    assert np.array_equal(dbg[a, 1], odbg['root'][a])
    assert np.array_equal(dbg[a, 2], odbg['sample_node'][a]), "sampled quadtree node"
    assert np.array_equal(dbg[a, 3], odbg['pdf_node'][a]), "node reached by the pdf of the sample"
    # the oracle's math is restated op for op (no FMA): directions and pdfs are bit-identical
    assert beq(ctx.host(d), od)
    assert beq(ctx.host(p), op)
    np.testing.assert_allclose(ctx.host(d), od, rtol=1e-5, atol=1e-7)       # the north_star tolerance
    np.testing.assert_allclose(ctx.host(p), op, rtol=1e-5)
    # pdf of arbitrary directions (+ axis-aligned and degenerate ones: tie rules, NaN, zero)
    dirs = rng.standard_normal((n, 3)).astype(F)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    dirs[:10] = np.array([[1, 0, 0], [0, 1, 0], [-1, 0, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1], [0, 0, 0],
                          [np.nan, 0, 1], [np.inf, 0, 0], [0.70710678, 0.70710678, 0]], F)
    pp, pdbg = t.pdf(ctx.dev(pos), ctx.dev(dirs), ctx.dev(active.astype(np.uint8)), debug=True)
    opp, opdbg = prev.pdf(pos, dirs, active, return_debug=True)
    pdbg = ctx.host(pdbg).view(U)
    assert np.array_equal(pdbg[a, 2], opdbg['pdf_node'][a])
    assert beq(ctx.host(pp), opp)
    # without the debug output the kernel reads the leaf's path product straight from its second jump table
    pp2 = t.pdf(ctx.dev(pos), ctx.dev(dirs), ctx.dev(active.astype(np.uint8)))
    assert beq(ctx.host(pp2), opp), "pdf through the path-product jump table"
    # sample + pdf of the given directions fused into one call (one spatial descent): the same bits as the two calls
    if explicit:
        fd, fp, fq = t.sample_pdf(ctx.dev(pos), ctx.dev(dirs), ctx.dev(active.astype(np.uint8)), u=ctx.dev(u))
    else:
        fd, fp, fq = t.sample_pdf(ctx.dev(pos), ctx.dev(dirs), ctx.dev(active.astype(np.uint8)), seed=1234, lane_offset=17)
    assert beq(ctx.host(fd), od) and beq(ctx.host(fp), op) and beq(ctx.host(fq), opp), "fused sample + pdf"
    return dict(pos=pos, active=active, dirs=dirs)


def case_train_refine_topology(ctx):
    """splat + refine x4: post-refine topology (all 23 arrays) bit-exact; then queries"""
    t, cur, prev = train(ctx, iters=4)
    s = t.sizes()
    assert s['error'] == 0
    assert s['n_kd'] == prev.kdTreeNode.getWidth() > 15
    assert s['n_quad'] == prev.quadTree.quadTreeNode.getWidth() > 200
    assert_tree_equal(t.download(0), prev)
    assert_tree_equal(t.download(1), cur)
    assert prev.validateTreeNodeBBox() and prev.quadTree.validateQuadTreeNodeBBox()
    check_queries(ctx, t, prev, explicit=True)
    check_queries(ctx, t, prev, explicit=False)


def case_threshold_reciprocal_variant(ctx):
    """semantics switch "quad_thr_reciprocal" (E * fp32(0.01) instead of E / 100): library and
    oracle agree bit for bit in that mode too, and the mode really changes some thresholds"""
    base = train(ctx, iters=3)[2].quadTree.quadTreeNode.refinementThreshold.copy()
    so.QUAD_THR_RECIPROCAL = True
    try:
        t, cur, prev = train(ctx, iters=3, tuning=(("quad_thr_reciprocal", 1),))
    finally:
        so.QUAD_THR_RECIPROCAL = False
    assert_tree_equal(t.download(0), prev)
    assert_tree_equal(t.download(1), cur)
    thr = prev.quadTree.quadTreeNode.refinementThreshold
    assert thr.shape != base.shape or not np.array_equal(thr, base)


def case_train_refine_nee_shallow(ctx):
    """NEE storage on, shallow depth caps (maxDepth limits bite), non-unit box"""
    t, cur, prev = train(ctx, iters=3, n=15000, max_leaf=300, kd_max_depth=5, quad_max_depth=4, store_nee=True, box=100.0)
    assert_tree_equal(t.download(0), prev)
    assert int(prev.kdTreeNode.depth.max()) == 5 and int(prev.quadTree.quadTreeNode.depth.max()) == 4
    check_queries(ctx, t, prev, box=100.0)


def case_jump_table_equals_descent(ctx):
    """the 32x32 jump table over the top 5 quadtree levels and the 8x8 second-stage tables over the next 3 answer pdf /
    splat descents with the same bits as the level-by-level descent (points on the 1/32 and 1/256 grid lines included:
    they take the slow paths)"""
    t, cur, prev = train(ctx, iters=6, n=30000, caps=dict(kd_capacity=1 << 14, quad_capacity=1 << 20))
    assert t.sizes()['jump_trees'] > 0 and t.sizes()['jump2_tables'] > 0
    assert int(prev.quadTree.quadTreeNode.depth.max()) > 9
    rng = np.random.default_rng(17)
    n = 20000
    pos = rng.random((n, 3)).astype(F)
    dirs = rng.standard_normal((n, 3)).astype(F)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    d2 = rng.random((n, 2)).astype(F)
    d2[:64] = (rng.integers(0, 33, (64, 2)) / 32.0).astype(F)            # exactly on the 1/32 grid lines
    d2[64:128, 0] = (rng.integers(0, 33, 64) / 32.0).astype(F)
    d2[128:192] = (rng.integers(0, 257, (64, 2)) / 256.0).astype(F)      # ... on the 1/256 lines of the second stage
    d2[192:256, 1] = (rng.integers(0, 257, 64) / 256.0).astype(F)
    m = slice(256, 6000)                                                  # many points in the lobes: the deep subtrees
    d2[m] = np.clip(np.stack([0.3 + 0.004 * rng.standard_normal(5744), 0.7 + 0.004 * rng.standard_normal(5744)], 1), 0, 1).astype(F)
    dirs[:6000] = dm.canonical_to_dir(d2[:6000])
    out = {}
    for use, use2 in ((1, 1), (1, 0), (0, 1)):
        t.set_tuning("use_jump", use)
        t.set_tuning("use_jump2", use2)
        p, dbg = t.pdf(ctx.dev(pos), ctx.dev(dirs), debug=True)
        p2 = t.pdf(ctx.dev(pos), ctx.dev(dirs))                           # through the path-product tables
        t.reset_stats()
        t.splat_records(ctx.dev(pos), ctx.dev(d2), ctx.dev(np.ones(n, F)), ctx.dev(np.ones(n, F)))
        out[(use, use2)] = (ctx.host(p).copy(), ctx.host(dbg).copy(), t.download(1)['quadtree_irradiance'].copy(), ctx.host(p2).copy())
    t.set_tuning("use_jump", 1)
    t.set_tuning("use_jump2", 1)
    t.reset_stats()
    ref = out[(0, 1)]
    for k in ((1, 1), (1, 0)):
        assert beq(out[k][0], ref[0]) and np.array_equal(out[k][1], ref[1]) and np.array_equal(out[k][2], ref[2]) and beq(out[k][3], ref[0]), k
    op, odbg = prev.pdf(pos, dirs, True, return_debug=True)
    assert beq(ref[0], op) and np.array_equal(ref[1].view(U)[:, 2], odbg['pdf_node'])
    depth = prev.quadTree.quadTreeNode.depth[odbg['pdf_node']]
    assert (depth > 8).sum() > 500 and ((depth > 5) & (depth <= 8)).sum() > 500      # both stages and the records below them were used


def case_spatial_descent_variants(ctx):
    """the three spatial-descent code paths of the kernels give the same leaves: grid over the first
    11 levels + staged tree (default), staged tree walked from the root, partly staged tree (global loads)"""
    t, cur, prev = train(ctx, iters=3, max_leaf=12, quad_max_depth=7, caps=dict(kd_capacity=1 << 14, quad_capacity=1 << 20))   # deep spatial tree: leaves beyond level 11
    assert t.sizes()["error"] == 0
    assert int(prev.kdTreeNode.depth.max()) > 11
    for grid, smem in ((1, 24576), (0, 24576), (1, 16), (0, 16)):     # kernel modes 2, 1, 3, 0 of sdt_kd_descend
        t.set_tuning("use_kd_grid", grid)
        t.set_tuning("kd_smem_nodes", smem)
        check_queries(ctx, t, prev, n=6000, seed=23)
        rec = dyadic_records(4000, 31, ((0.3, 0.7, 0.02),))
        t.reset_stats()
        splat(t, ctx, rec)
        cur.resetTreeVertCount(); cur.resetAllQuadTreeIrradiance()
        cur.addDataPropagate(rec)
        assert_tree_equal(t.download(1), cur)
    t.set_tuning("kd_smem_nodes", 24576)
    t.set_tuning("use_kd_grid", 1)
    # points exactly on split planes of the first levels (the right child wins, src/kdtree.py:462-468)
    g = np.array([0.0, 0.125, 0.25, 0.375, 0.5, 0.625, 0.75, 0.875, 1.0], F)
    pos = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    leaf, root = t.locate(ctx.dev(pos))
    assert np.array_equal(ctx.host(leaf).view(U), prev.getLeafNodeIndex(pos))


def case_grid_cell_boundaries(ctx):
    """the kernels find the 16x16x8 grid cell of a vertex from a scaled guess corrected against the exact cell
    boundaries (sdt_axis_cell) instead of walking 11 midpoint halvings: points exactly ON every boundary of an
    awkward, non-dyadic box and one ulp to either side must land in the reference's leaf (right child wins on the
    plane, src/kdtree.py:462-468), with the tree deep enough that every one of the 11 levels exists somewhere"""
    lo = np.array([-1.37, 0.1, 3.0e-3], F)
    hi = np.array([2.91, 0.1000061, 7.7e3], F)
    t = ctx.make(bbox_min=tuple(float(v) for v in lo), bbox_max=tuple(float(v) for v in hi), kd_max_depth=20, quad_max_depth=2,
                 store_nee=False, kd_capacity=1 << 17, quad_capacity=1 << 20)
    cur, prev = oracle_pair(tuple(float(v) for v in lo), tuple(float(v) for v in hi), 20, 2, False)
    rng = np.random.default_rng(77)
    n = 30000
    for it in range(3):
        pos = (lo + rng.random((n, 3)).astype(F) * (hi - lo)).astype(F)
        pos = np.minimum(np.maximum(pos, lo), hi)
        rec = so.SurfaceInteractionRecord(pos, rng.random((n, 2)).astype(F), np.ones(n, F), np.ones(n, F))
        splat(t, ctx, rec)
        cur.addDataPropagate(rec)
        t.set_max_leaf_size(3)
        t.refine()
        oracle_refine(cur, prev, 3)
    assert t.sizes()['error'] == 0 and int(prev.kdTreeNode.depth.max()) >= 12, int(prev.kdTreeNode.depth.max())
    assert_tree_equal(t.download(0), prev)

    def bounds(a, b, nlev):               # the reference's own midpoints, level by level, in fp32
        v = [F(a), F(b)]
        for _ in range(nlev):
            w = [v[0]]
            for x, y in zip(v[:-1], v[1:]):
                w += [F((F(x) + F(y)) / F(2.0)), y]
            v = w
        return np.array(v, F)
    pts = []
    for ax, nlev in enumerate((5, 5, 4)):           # one level finer than the grid: planes below it too
        b = bounds(lo[ax], hi[ax], nlev)
        cand = np.concatenate([b, np.nextafter(b, F(np.inf)), np.nextafter(b, F(-np.inf))]).astype(F)
        cand = cand[(cand >= lo[ax]) & (cand <= hi[ax])]
        other = (lo + rng.random((len(cand), 3)).astype(F) * (hi - lo)).astype(F)
        other = np.minimum(np.maximum(other, lo), hi)
        other[:, ax] = cand
        pts.append(other)
        both = other.copy()                          # ... and on boundaries of two axes at once
        ax2 = (ax + 1) % 3
        b2 = bounds(lo[ax2], hi[ax2], 4)
        both[:, ax2] = b2[rng.integers(0, len(b2), len(both))]
        pts.append(both)
    pos = np.concatenate(pts).astype(F)
    want = prev.getLeafNodeIndex(pos)
    for grid, smem in ((1, 24576), (1, 16), (0, 24576)):
        t.set_tuning("use_kd_grid", grid)
        t.set_tuning("kd_smem_nodes", smem)
        leaf, root = t.locate(ctx.dev(pos))
        assert np.array_equal(ctx.host(leaf).view(U), want), (grid, smem)
    t.set_tuning("kd_smem_nodes", 24576)
    t.set_tuning("use_kd_grid", 1)


def case_deep_quadtree_beyond_fp32_grid(ctx):
    """QuadTree.maxDepth 28: every record has the same direction, so the tree is a chain down to level 28,
    past the level (23) where cell corners stop being exact in fp32 -- the sampler falls back to float
    cell tracking and must reproduce the reference's rounding of (min+max)/2"""
    t = ctx.make(kd_capacity=64, quad_capacity=1 << 12, kd_max_depth=4, quad_max_depth=28, store_nee=False)
    cur, prev = oracle_pair(kd_max_depth=4, quad_max_depth=28)
    rng = np.random.default_rng(3)
    n = 4000
    for it in range(8):
        rec = so.SurfaceInteractionRecord(rng.random((n, 3)).astype(F), np.tile(np.array([[0.3, 0.7]], F), (n, 1)),
                                          np.ones(n, F), np.ones(n, F))
        rec.direction[: n // 8] = rng.random((n // 8, 2)).astype(F)       # a little energy elsewhere
        splat(t, ctx, rec)
        cur.addDataPropagate(rec)
        t.set_max_leaf_size(1e9)
        t.refine()
        oracle_refine(cur, prev, 1e9)
    assert int(prev.quadTree.quadTreeNode.depth.max()) == 28 and t.sizes()['error'] == 0
    assert_tree_equal(t.download(0), prev)
    check_queries(ctx, t, prev, n=3000, seed=5, explicit=True)
    check_queries(ctx, t, prev, n=3000, seed=6, explicit=False)
    # pdf exactly at / next to the chain's direction
    d = dm.canonical_to_dir(np.array([[0.3, 0.7], [np.nextafter(F(0.3), F(1)), 0.7], [0.3, np.nextafter(F(0.7), F(0))]], F))
    pos = np.full((3, 3), 0.5, F)
    p, dbg = t.pdf(ctx.dev(pos), ctx.dev(d), debug=True)
    op, odbg = prev.pdf(pos, d, True, return_debug=True)
    assert beq(ctx.host(p), op) and np.array_equal(ctx.host(dbg).view(U)[:, 2], odbg['pdf_node'])


def case_fused_equals_two_descents(ctx):
    t, cur, prev = train(ctx, iters=3)
    rng = np.random.default_rng(3)
    n = 8192
    pos = rng.random((n, 3)).astype(F)
    t.set_tuning("fuse_sample_pdf", 1)
    d1, p1, g1 = t.sample(ctx.dev(pos), seed=9, debug=True)
    t.set_tuning("fuse_sample_pdf", 0)
    d0, p0, g0 = t.sample(ctx.dev(pos), seed=9, debug=True)
    assert np.array_equal(ctx.host(d1).view(U), ctx.host(d0).view(U))
    assert np.array_equal(ctx.host(p1).view(U), ctx.host(p0).view(U))
    assert np.array_equal(ctx.host(g1), ctx.host(g0))
    t.set_tuning("fuse_sample_pdf", 1)


def case_host_pipeline_chunks(ctx):
    """large SDT_HOST_PTRS calls run as a chunked 3-stream pipeline: same results, chunk edges included"""
    t, cur, prev = train(ctx, iters=3)
    t.set_tuning("host_chunk", 1000)                  # 9 chunks of 1000 + a ragged tail
    check_queries(ctx, t, prev, n=9531, explicit=True)
    check_queries(ctx, t, prev, n=9531, explicit=False)
    rec = dyadic_records(9531, 77, ((0.3, 0.7, 0.02),))
    splat(t, ctx, rec)
    cur.addDataPropagate(rec)
    assert_tree_equal(t.download(1), cur)


def case_host_calls_no_wait(ctx):
    """SDT_NO_WAIT: back-to-back host-pointer calls (pipelined and small ones mixed, different buffer layouts) return
    before their outputs landed and overlap on the device; after one synchronize the results are those of the
    waiting calls.  (Host buffers only: with device buffers the flag does nothing.)"""
    t, cur, prev = train(ctx, iters=3)
    rng = np.random.default_rng(5)
    n = 20000
    pos = rng.random((n, 3)).astype(F)
    dirs = rng.standard_normal((n, 3)).astype(F)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    rec = dyadic_records(n, 78, ((0.3, 0.7, 0.02),))
    want_d, want_p = prev.sample(pos, so.ExplicitSampler(seed=7, n=n, lane_offset=0), np.ones(n, bool))
    want_q = prev.pdf(pos, dirs, np.ones(n, bool))
    t.set_tuning("host_chunk", 1500)
    for rep in range(2):
        t.host_wait = False
        outs = []
        for k in range(3):                              # 3 rounds in flight: 13 chunks + ragged tail each, then a small call
            d, p = t.sample(pos, seed=7)
            q = t.pdf(pos, dirs)
            qs = t.pdf(pos[:500], dirs[:500])           # small call: whole-arena staging on the call's stream
            splat(t, Ctx(make=None), rec)
            outs.append((d, p, q, qs))
        t.synchronize()
        t.host_wait = True
        for d, p, q, qs in outs:
            assert beq(d, want_d) and beq(p, want_p) and beq(q, want_q) and beq(qs, want_q[:500])
    for _ in range(6):                                  # 120 k dyadic records: the fp32 sums stay exact
        cur.addDataPropagate(rec)
    assert_tree_equal(t.download(1), cur)


def case_splat_float_tolerance(ctx):
    """general fp32 radiance: energies within 1e-4 relative of the exactly-rounded sums;
    conservation root = sum leaves = sum radiance/woPdf (src/quadtree.py:1205-1218)"""
    t, cur, prev = train(ctx, iters=3)
    rng = np.random.default_rng(11)
    n = 50000
    rec = so.SurfaceInteractionRecord(rng.random((n, 3)).astype(F) * 1.05 - 0.02, rng.random((n, 2)).astype(F),
                                      rng.lognormal(0, 1, n).astype(F), (rng.random(n) * 1.95 + 0.05).astype(F))
    rec.direction[:5] = np.array([[0.5, 0.5], [0, 0], [1, 1], [1.5, 0.5], [np.nan, 0.2]], F)
    active = rng.random(n) < 0.95
    splat(t, ctx, rec, active)
    sub = so.SurfaceInteractionRecord(rec.position[active], rec.direction[active], rec.radiance[active], rec.woPdf[active])
    cur.addDataPropagate(sub, exact=True)
    got = t.download(1)
    np.testing.assert_array_equal(got['kdtree_vertCount'], cur.kdTreeNode.vertCount)
    np.testing.assert_allclose(got['quadtree_irradiance'], cur.quadTree.quadTreeNode.irradiance, rtol=1e-4, atol=1e-6)
    q = cur.quadTree.quadTreeNode
    roots = got['quadtree_irradiance'][:q.rootNodeIndex.shape[0]].astype(np.float64).sum()
    leaves = got['quadtree_irradiance'][got['quadtree_isLeaf']].astype(np.float64).sum()
    inside = dm_inside(sub.direction)
    truth = (sub.radiance[inside].astype(np.float64) / sub.woPdf[inside]).sum()
    assert abs(roots - truth) <= 1e-4 * truth and abs(leaves - truth) <= 1e-4 * truth
    inbox = so.bbox_contains(F(0), F(1), sub.position)
    assert got['kdtree_vertCount'][got['kdtree_isLeaf']].sum() == inbox.sum() == got['kdtree_vertCount'][0]


def dm_inside(d):
    with np.errstate(invalid='ignore'):
        return np.all((d >= 0) & (d <= 1), axis=1)


def case_path_data(ctx):
    """processPathData + filter + splat in one call vs the oracle's three steps"""
    t, cur, prev = train(ctx, iters=2, store_nee=True)
    rng = np.random.default_rng(21)
    rays, md = 3000, 5
    n = rays * md
    Lf = rng.random((rays, 3)).astype(F) * 4
    tr = (rng.random((n, 3)) * 2).astype(F)
    tb = rng.random((n, 3)).astype(F)
    bsdf = rng.random((n, 3)).astype(F)
    tb[rng.random(n) < 0.1] = 0           # -> inf / NaN scrubbing
    bsdf[rng.random(n) < 0.1] = 0
    tr[rng.random(n) < 0.05] = np.nan
    pos = rng.random((n, 3)).astype(F)
    d = rng.random((n, 2)).astype(F)
    wo = (rng.random(n) * 2).astype(F)
    wo[rng.random(n) < 0.1] = 0
    wo[rng.random(n) < 0.05] = np.nan
    nee = (rng.random((n, 3)) * (rng.random((n, 1)) < 0.5)).astype(F)
    nee[rng.random(n) < 0.05, 1] = np.nan
    dnee = rng.random((n, 2)).astype(F)
    active = rng.random(n) < 0.7
    rad = t.splat_path_data(md, ctx.dev(Lf), ctx.dev(tr), ctx.dev(tb), ctx.dev(bsdf), ctx.dev(pos), ctx.dev(d), ctx.dev(wo),
                            ctx.dev(nee), ctx.dev(dnee), ctx.dev(active.astype(np.uint8)), want_radiance=True)
    _, orad = so.process_path_data(Lf, tr, tb, bsdf, md)
    keep, orad2, onee = so.filter_records(active, orad, nee, wo)
    assert beq(ctx.host(rad), orad2)
    sub = so.SurfaceInteractionRecord(pos[keep], d[keep], orad2[keep], wo[keep], onee[keep], dnee[keep])
    cur.addDataPropagate(sub, exact=True)
    got = t.download(1)
    np.testing.assert_array_equal(got['kdtree_vertCount'], cur.kdTreeNode.vertCount)
    e = cur.quadTree.quadTreeNode.irradiance
    with np.errstate(invalid='ignore'):
        fin = np.isfinite(e)
    assert np.array_equal(np.isfinite(got['quadtree_irradiance']), fin)
    np.testing.assert_allclose(got['quadtree_irradiance'][fin], e[fin], rtol=1e-4, atol=1e-6)


def case_mis(ctx):
    rng = np.random.default_rng(4)
    n = 5000
    a = [rng.random(n).astype(F) * 3 for _ in range(5)]
    a[0][:5] = [0, 1, np.inf, 0, np.nan]
    a[4][:5] = [0, 0, np.inf, 1, 1]
    delta = rng.random(n) < 0.2
    t = ctx.make(kd_capacity=16, quad_capacity=64)
    for it in (1, 2):
        s, m = t.mis_nee(ctx.dev(a[0]), ctx.dev(a[1]), ctx.dev(a[2]), ctx.dev(a[3]), ctx.dev(a[4]),
                         ctx.dev(delta.astype(np.uint8)), 0.5, it)
        os_, om = so.nee_mis(a[0], a[1], a[2], a[3], a[4], delta, 0.5, it)
        assert beq(ctx.host(s), os_)
        assert beq(ctx.host(m), om)
    val = rng.random((n, 3)).astype(F)
    do = rng.random(n) < 0.6
    wo, w = t.mis_mixture(ctx.dev(a[0]), ctx.dev(a[1]), ctx.dev(val), ctx.dev(do.astype(np.uint8)), 0.5)
    owo, ow = so.mixture(a[0], a[1], val, do, 0.5)
    assert beq(ctx.host(wo), owo)
    assert beq(ctx.host(w), ow)


def case_guided_bounce(ctx):
    """sdt_guided = sample on mode-1 lanes, pdf + fused mixture on mode-2 lanes, in one pass"""
    t, cur, prev = train(ctx, iters=3)
    rng = np.random.default_rng(8)
    n = 6000
    pos = rng.random((n, 3)).astype(F)
    mode = rng.integers(0, 3, n).astype(np.uint8)
    wo = rng.standard_normal((n, 3)).astype(F)
    wo /= np.linalg.norm(wo, axis=1, keepdims=True)
    bp = (rng.random(n) * 2).astype(F)
    bv = rng.random((n, 3)).astype(F)
    d, sp, wp, wt = t.guided(ctx.dev(pos), ctx.dev(mode), wo=ctx.dev(wo), seed=77, bsdf_pdf=ctx.dev(bp), bsdf_value=ctx.dev(bv))
    d, sp, wp, wt = ctx.host(d), ctx.host(sp), ctx.host(wp), ctx.host(wt)
    m1, m2 = mode == 1, mode == 2
    od, op = prev.sample(pos, so.ExplicitSampler(seed=77, n=n), m1)
    assert beq(d[m1], od[m1]) and beq(sp[m1], op[m1])
    op2 = prev.pdf(pos, wo, m2)
    assert beq(sp[m2], op2[m2])
    owo, ow = so.mixture(bp, op2, bv, m2, 0.5)
    assert beq(wp[m2], owo[m2]) and beq(wt[m2], ow[m2])
    assert not d[mode == 0].any() and not sp[mode == 0].any()
    # the prepared form of the same call (argument struct built once, re-issued on the same buffers): same results every
    # time, and a new seed draws new directions
    call = t.prepare_guided(ctx.dev(pos), ctx.dev(mode), wo=ctx.dev(wo), seed=77, bsdf_pdf=ctx.dev(bp), bsdf_value=ctx.dev(bv))
    for _ in range(2):
        d2, sp2, wp2, wt2 = (ctx.host(x).copy() for x in call())
        assert beq(d2, d) and beq(sp2, sp) and beq(wp2, wp) and beq(wt2, wt)
    d3 = ctx.host(call(seed=78)[0])
    od3, _ = prev.sample(pos, so.ExplicitSampler(seed=78, n=n), m1)
    assert beq(d3[m1], od3[m1])


def case_refine_flags_and_frozen_stats(ctx):
    """refine from frozen stat buffers (sdt_upload_stats), KD-only and quad-only"""
    t, cur, prev = train(ctx, iters=2)
    rng = np.random.default_rng(31)
    k, q = cur.kdTreeNode, cur.quadTree.quadTreeNode
    # arbitrary (non-additive) frozen statistics, interior values included
    q.irradiance[:] = (rng.random(q.getWidth()) ** 4 * 50).astype(F)
    k.vertCount[:] = rng.integers(0, 4000, k.getWidth()).astype(F)
    t.upload_stats(q.irradiance, k.vertCount)
    t.set_max_leaf_size(700)
    t.refine(kd=True, quad=False)
    oracle_refine(cur, prev, 700, kd=True, quad=False)
    assert_tree_equal(t.download(0), prev)
    q = cur.quadTree.quadTreeNode
    q.irradiance[:] = (rng.random(q.getWidth()) ** 6 * 80).astype(F)
    R = q.rootNodeIndex.shape[0]
    q.irradiance[:R] = (2000 * (0.5 + rng.random(R))).astype(F)      # thr 10..30: leaves split 0..2 levels
    t.upload_stats(q.irradiance, None)
    t.refine(kd=False, quad=True)
    oracle_refine(cur, prev, 700, kd=False, quad=True)
    assert_tree_equal(t.download(0), prev)


def case_record_bound_hint(ctx):
    """sdt_hint_records: after statistics arrive from elsewhere (an all-reduce, sdt_upload_stats) the refine launches every
    split round unless the caller bounds the records behind them; the right bound gives the same tree with fewer launches,
    a bound that is too small raises device error flag 8"""
    from practical_path_guiding_lab_b200.sdtree import SDTreeError
    rec = dyadic_records(20000, 7, ((0.3, 0.7, 0.02),))
    trees, launches = [], []
    for bound in (None, 20000, 2 * 20000 + 5):
        t = ctx.make(bbox_min=(0, 0, 0), bbox_max=(1, 1, 1), kd_max_depth=20, quad_max_depth=20, store_nee=False,
                     kd_capacity=1 << 14, quad_capacity=1 << 18)
        splat(t, ctx, rec)
        tr = t.download(1)
        t.reset_stats()
        # "another rank's" identical statistics arrive: twice the counts / energies of one splat
        t.upload_stats(2 * tr['quadtree_irradiance'], 2 * tr['kdtree_vertCount'])
        if bound is not None:
            t.hint_records(bound)
        t.set_max_leaf_size(300)
        l0 = t.kernel_launches()
        t.refine()
        launches.append(t.kernel_launches() - l0)
        trees.append((t, t.download(0), t.sizes()['error']))
    (t0, full, e0), (t1, short, e1), (t2, good, e2) = trees
    assert e0 == 0 and e2 == 0 and e1 == 8, (e0, e1, e2)
    for k in full:
        np.testing.assert_array_equal(full[k], good[k], err_msg=k)
    assert launches[2] < launches[0], launches                       # the split rounds nobody can reach were not launched
    assert short['kdtree_depth'].shape[0] < full['kdtree_depth'].shape[0]      # the wrong bound cut the splits short ...
    try:
        t1.check_error()                                               # ... and that does not go unnoticed
        raise AssertionError("a record bound that is too small was not surfaced")
    except SDTreeError as e:
        assert e.code == -1 and "hint_records" in str(e)


def case_capacity_error(ctx):
    t = ctx.make(kd_capacity=8, quad_capacity=64, kd_max_depth=20, quad_max_depth=20, store_nee=False)
    rec = dyadic_records(20000, 1)
    splat(t, ctx, rec)
    t.set_max_leaf_size(100)
    t.refine()
    s = t.sizes()
    assert s['error'] != 0 and s['n_kd'] <= 8 and s['n_quad'] <= 64
    # the tree is still a valid tree: queries run
    leaf, root = t.locate(ctx.dev(rec.position[:100]))
    assert ctx.host(root).view(U).max() < s['n_roots']
    # ... but it is NOT the reference's tree any more: the Python face raises on request (the integrator asks every iteration)
    from practical_path_guiding_lab_b200.sdtree import SDTreeError
    try:
        t.check_error()
        raise AssertionError("arena exhaustion not surfaced")
    except SDTreeError as e:
        assert e.code == -3 and "arena" in str(e)
    splat(t, ctx, rec)
    try:
        t.refine(check=True)
        raise AssertionError("arena exhaustion not surfaced by refine(check=True)")
    except SDTreeError as e:
        assert e.code == -3


def case_zero_total_energy(ctx):
    """negative radiance can cancel a tree's total energy to exactly 0: the threshold E_root/100 is then 0 and every
    leaf that holds any positive energy splits down to the depth cap, in the reference too (src/quadtree.py:519,
    615-637).  Moderate cap: bit for bit like the oracle; deep cap: arena exhaustion reported (error bit 2) and the
    handle stays usable"""
    n = 4000
    rng = np.random.default_rng(3)
    pos = rng.random((n, 3)).astype(F)
    first = so.SurfaceInteractionRecord(pos, rng.random((n, 2)).astype(F), np.ones(n, F), np.ones(n, F))
    spot = np.tile(np.array([[0.40625, 0.40625]], F), (n, 1))           # all the negative energy in one cell
    second = so.SurfaceInteractionRecord(np.concatenate([pos, pos]), np.concatenate([first.direction, spot]),
                                         np.concatenate([np.ones(n, F), -np.ones(n, F)]), np.ones(2 * n, F))
    for qd, cap in ((6, 1 << 15), (20, 1 << 12)):
        t = ctx.make(kd_max_depth=2, quad_max_depth=qd, store_nee=False, kd_capacity=64, quad_capacity=cap)
        cur, prev = oracle_pair((0, 0, 0), (1, 1, 1), 2, qd, False)
        for rec in (first, second):
            splat(t, ctx, rec)
            t.set_max_leaf_size(1e9)                                    # one spatial leaf: the tree's total is exactly 0
            t.refine()
            if qd == 6:
                cur.addDataPropagate(rec)
                oracle_refine(cur, prev, 1e9)
        s = t.sizes()
        if qd == 6:
            assert prev.quadTree.quadTreeNode.refinementThreshold[0] == 0 and s['error'] == 0
            assert int(prev.quadTree.quadTreeNode.depth.max()) == 6 and s['n_quad'] > 3000       # nearly the full 4^6 tree
            assert_tree_equal(t.download(0), prev)
            check_queries(ctx, t, prev, n=512)
        else:
            assert s['error'] == 2 and s['n_quad'] <= cap
            dd, p = t.sample(ctx.dev(pos), seed=1)
            assert ctx.host(p).shape == (n,)


def case_edge_inputs_and_errors(ctx):
    """empty / single / ragged wavefronts, SoA (Dr.Jit-style) component planes, error reporting"""
    from practical_path_guiding_lab_b200 import SDTreeError
    t, cur, prev = train(ctx, iters=2)
    z3 = np.zeros((0, 3), F)
    d, p = t.sample(ctx.dev(z3), seed=1)
    assert ctx.host(d).shape == (0, 3) and ctx.host(p).shape == (0,)
    assert ctx.host(t.pdf(ctx.dev(z3), ctx.dev(z3))).shape == (0,)
    t.splat_records(ctx.dev(z3), ctx.dev(np.zeros((0, 2), F)), ctx.dev(np.zeros(0, F)), ctx.dev(np.zeros(0, F)))
    for n in (1, 31, 33, 1025):
        check_queries(ctx, t, prev, n=max(n, 16), seed=n)
    # SoA planes (stride 1) == interleaved (stride 3)
    rng = np.random.default_rng(12)
    n = 777
    pos = rng.random((n, 3)).astype(F)
    planes = tuple(ctx.dev(np.ascontiguousarray(pos[:, k])) for k in range(3))
    d1, p1 = t.sample(ctx.dev(pos), seed=5)
    d2, p2 = t.sample(planes, seed=5)
    assert beq(ctx.host(d1), ctx.host(d2)) and beq(ctx.host(p1), ctx.host(p2))
    dirs = rng.standard_normal((n, 3)).astype(F)
    dplanes = tuple(ctx.dev(np.ascontiguousarray(dirs[:, k])) for k in range(3))
    assert beq(ctx.host(t.pdf(ctx.dev(pos), ctx.dev(dirs))), ctx.host(t.pdf(planes, dplanes)))
    # errors: negative status + message, the handle stays usable
    bad = dict(prev.to_arrays())
    bad['kdtree_child_right_index'] = bad['kdtree_child_right_index'].copy()
    bad['kdtree_child_right_index'][0] += 1
    t2 = ctx.make(kd_capacity=1 << 14, quad_capacity=1 << 18)
    try:
        t2.upload(bad)
        raise AssertionError("invalid spatial layout accepted")
    except SDTreeError as e:
        assert e.code == -4 and "adjacent" in str(e)
    # boxes that are not the midpoint split on axis depth % 3 (the descent recomputes planes, it does not read boxes)
    bad = dict(prev.to_arrays())
    bad['kdtree_bbox_max'] = bad['kdtree_bbox_max'].copy()
    left = int(bad['kdtree_child_left_index'][0])
    bad['kdtree_bbox_max'][left, 0] = np.nextafter(bad['kdtree_bbox_max'][left, 0], F(2))
    try:
        t2.upload(bad)
        raise AssertionError("non-midpoint spatial boxes accepted")
    except SDTreeError as e:
        assert e.code == -4 and "midpoint" in str(e)
    small = ctx.make(kd_capacity=4, quad_capacity=16)
    try:
        small.upload(prev.to_arrays())
        raise AssertionError("oversized tree accepted")
    except SDTreeError as e:
        assert e.code == -3
    try:
        t.set_tuning("no_such_key", 1)
        raise AssertionError("unknown tuning key accepted")
    except SDTreeError as e:
        assert e.code == -1
    t2.upload(prev.to_arrays())
    assert_tree_equal(t2.download(0), prev)


def case_counter_generator_statistics(ctx):
    """the perf-mode generator (hash for the lane key and the leaf position, an LCG stream for the per-level child
    selection) must sample the tree's distribution: leaf frequencies of 2^18 draws against prod(E_child / sum of
    siblings) (chi-square over all levels jointly), the position inside the leaves uniform, and the mean of
    1 / (4 pi pdf) equal to the fraction of the sphere that has energy (the estimator the integrator relies on)"""
    t, cur, prev = train(ctx, iters=4)
    q = prev.quadTree.quadTreeNode
    n = 1 << 18
    for point, seed in (((0.3, 0.3, 0.3), 11), ((0.8, 0.6, 0.1), 4242)):
        pos = np.tile(np.array([point], F), (n, 1))
        d, p, dbg = t.sample(ctx.dev(pos), seed=seed, debug=True)
        dbg = ctx.host(dbg).view(U)
        node = dbg[:, 2]
        root = int(q.rootNodeIndex[dbg[0, 1]])
        prob = {root: 1.0}
        stack = [root]
        leaves = []
        while stack:
            v = stack.pop()
            if q.isLeaf[v]:
                leaves.append(v)
                continue
            ch = [int(c[v]) for c in (q.child_1_index, q.child_2_index, q.child_3_index, q.child_4_index)]
            e = np.array([q.irradiance[c] for c in ch], np.float64)
            for c, w in zip(ch, e / e.sum()):
                prob[c] = prob[v] * w
                stack.append(c)
        assert len(leaves) > 20
        obs = np.bincount(node, minlength=q.getWidth())[leaves].astype(np.float64)
        exp = np.array([prob[v] for v in leaves]) * n
        assert obs.sum() == n
        big = exp >= 8
        chi2 = float((((obs - exp) ** 2) / np.maximum(exp, 1e-300))[big].sum())
        df = int(big.sum()) - 1
        assert abs(chi2 - df) < 5.0 * np.sqrt(2.0 * df) + 5.0, (chi2, df)
        pdf = ctx.host(p).astype(np.float64)
        est = (1.0 / (4.0 * np.pi * pdf)).mean()                        # = integral of 1/(4 pi) over the sampled support
        sd = (1.0 / (4.0 * np.pi * pdf)).std() / np.sqrt(n)
        support = sum(float(np.prod(q.bbox_max[v] - q.bbox_min[v])) for v in leaves if prob[v] > 0)   # zero-energy leaves are never sampled
        assert abs(est - support) < 5.0 * sd + 1e-3, (est, support, sd)
        # position inside the most frequent leaf: both canonical coordinates uniform over the cell
        v = leaves[int(np.argmax(obs))]
        sel = node == v
        c2 = dm.dir_to_canonical(ctx.host(d)[sel])
        lo, hi = q.bbox_min[v], q.bbox_max[v]
        uu = (c2 - lo) / (hi - lo)
        m = int(sel.sum())
        for k in range(2):
            h = np.bincount(np.clip((uu[:, k] * 16).astype(int), 0, 15), minlength=16).astype(np.float64)
            c = float(((h - m / 16.0) ** 2 / (m / 16.0)).sum())
            assert c < 15 + 5.0 * np.sqrt(30.0) + 5.0, (k, c)


def case_golden_fixture(ctx):
    """the committed golden file (tests/golden/sdtree_golden.npz, written by make_sdtree_golden.py with
    the oracle): upload its tree, replay its queries and records, compare with its stored answers"""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sdtree_golden.npz"))
    tree = {k[5:]: g[k] for k in g.files if k.startswith("tree_")}
    t = ctx.make(kd_capacity=1 << 12, quad_capacity=1 << 17)
    t.upload(tree)
    a = g['active']
    act = ctx.dev(a.astype(np.uint8))
    leaf, root = t.locate(ctx.dev(g['pos']), act)
    assert np.array_equal(ctx.host(leaf).view(U), g['leaf']) and np.array_equal(ctx.host(root).view(U)[a], g['root'][a])
    d, p, dbg = t.sample(ctx.dev(g['pos']), act, u=ctx.dev(g['u']), debug=True)
    dbg = ctx.host(dbg).view(U)
    assert np.array_equal(dbg[a, 2], g['sample_node'][a]) and np.array_equal(dbg[a, 3], g['pdf_node'][a])
    assert beq(ctx.host(d), g['sample_dir']) and beq(ctx.host(p), g['sample_pdf'])
    np.testing.assert_allclose(ctx.host(p), g['sample_pdf'], rtol=1e-5)              # the north_star tolerance
    d, p = t.sample(ctx.dev(g['pos']), act, seed=77, lane_offset=3)
    assert beq(ctx.host(d), g['counter_dir']) and beq(ctx.host(p), g['counter_pdf'])
    pp, pdbg = t.pdf(ctx.dev(g['pos']), ctx.dev(g['dirs']), act, debug=True)
    assert beq(ctx.host(pp), g['pdf']) and np.array_equal(ctx.host(pdbg).view(U)[a, 2], g['pdf_query_node'][a])
    t.splat_records(ctx.dev(g['rec_position']), ctx.dev(g['rec_direction']), ctx.dev(g['rec_radiance']), ctx.dev(g['rec_wo_pdf']))
    cur = t.download(1)
    assert np.array_equal(cur['kdtree_vertCount'], g['splat_vert_count']) and beq(cur['quadtree_irradiance'], g['splat_irradiance'])


def case_npz_roundtrip(ctx, tmp_path):
    t, cur, prev = train(ctx, iters=2)
    f = str(tmp_path / "tree.npz")
    t.save_npz(f)
    d = np.load(f)
    assert set(d.files) == set(so.KDTree.NPZ_KEYS)
    o = so.KDTree()
    o.loadFromFile(f)                      # the oracle's reader follows src/kdtree.py:156-170
    t2 = ctx.make(kd_capacity=1 << 14, quad_capacity=1 << 18)
    t2.load_npz(f)
    assert_tree_equal(t2.download(0), o)
    check_queries(ctx, t2, o, n=1024)


ALL_CASES = [case_golden_fixture, case_grid_cell_boundaries, case_counter_generator_statistics, case_deep_quadtree_beyond_fp32_grid, case_jump_table_equals_descent, case_spatial_descent_variants, case_initial_tree, case_golden_upload_download, case_train_refine_topology, case_threshold_reciprocal_variant, case_zero_total_energy, case_host_pipeline_chunks, case_host_calls_no_wait,
             case_train_refine_nee_shallow, case_fused_equals_two_descents, case_splat_float_tolerance,
             case_path_data, case_mis, case_guided_bounce, case_refine_flags_and_frozen_stats,
             case_capacity_error, case_record_bound_hint, case_edge_inputs_and_errors]
