"""Pins the oracle (oracle/sdtree_oracle.py) on the only machine-checkable facts the
reference offers for this path: topology vectors hand-derived from the
reference source, the reference's own validators (child bbox inside parent) and its
conservation checks (root energy = sum of leaves = sum radiance/woPdf; sum of leaf
counts = number of records).  The reference has no stored expected outputs; the oracle is
pinned to the reference's own source run on numpy stand-ins for Dr.Jit / Mitsuba in
tests/test_reference_on_shim.py (the primitives' semantics stay assumed, see the oracle header)."""
import numpy as np

from oracle import sdtree_oracle as so
from oracle import drjit_math as dm

F = np.float32
U = np.uint32


def _records(n, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    return so.SurfaceInteractionRecord(
        position=(rng.random((n, 3)) * scale).astype(F),
        direction=rng.random((n, 2)).astype(F),
        radiance=rng.random(n).astype(F),
        woPdf=(rng.random(n) * 0.9 + 0.1).astype(F))


def test_kd_two_uniform_splits_golden():
    """src/kdtree.py:706-712 self-test; expected arrays hand-derived from :229-323."""
    t = so.KDTree()
    t.setup([0, 0, 0], [100, 100, 100])
    t.split(t.getAllLeafNodeIndex())
    t.split(t.getAllLeafNodeIndex())
    k = t.kdTreeNode
    assert k.getWidth() == 7
    assert k.child_left_index.tolist() == [1, 3, 5, 0, 0, 0, 0]
    assert k.child_right_index.tolist() == [2, 4, 6, 0, 0, 0, 0]
    assert k.depth.tolist() == [0, 1, 1, 2, 2, 2, 2]
    assert k.isLeaf.tolist() == [False, False, False, True, True, True, True]
    assert k.quadTreeRootIndex.tolist() == [0, 0, 1, 0, 2, 1, 3]
    assert k.bbox_min[5].tolist() == [50, 0, 0] and k.bbox_max[5].tolist() == [100, 50, 100]
    assert t.quadTree.quadTreeNode.rootNodeIndex.tolist() == [0, 1, 2, 3]
    assert t.validateTreeNodeBBox()
    leaf = t.getLeafNodeIndex(np.array([[75, 25, 25]], F))
    assert leaf.tolist() == [5]
    assert k.quadTreeRootIndex[leaf].tolist() == [1]


def test_kd_descent_tie_and_outside():
    """src/kdtree.py:446-468: right child wins on the plane; outside / NaN -> node 0."""
    t = so.KDTree()
    t.setup([0, 0, 0], [100, 100, 100])
    t.split(t.getAllLeafNodeIndex())
    t.split(t.getAllLeafNodeIndex())
    p = np.array([[50, 50, 1], [50, 49, 1], [49, 50, 1], [0, 0, 0], [100, 100, 100],
                  [101, 1, 1], [np.nan, 1, 1], [-1e-3, 5, 5]], F)
    assert t.getLeafNodeIndex(p).tolist() == [6, 5, 4, 3, 6, 0, 0, 0]
    # inactive lanes stay on node 0
    assert t.getLeafNodeIndex(p, np.zeros(8, bool)).tolist() == [0] * 8


def test_quadtree_three_uniform_splits_golden():
    """src/quadtree.py:1142-1152; split order hand-derived from :288-345 and :96-119."""
    q = so.QuadTree()
    root = np.zeros(1, U)
    for _ in range(3):
        q.quadTreeNode.split(q.getAllLeafNodeIndex(root))
    n = q.quadTreeNode
    assert n.getWidth() == 85
    assert q.getAllLeafNodeIndex(root)[:0].tolist() == []
    c1 = n.child_1_index
    assert c1[0] == 1 and c1[1] == 5 and c1[2] == 9 and c1[3] == 13 and c1[4] == 17
    # round-3 order [5,9,13,17, 6,10,14,18, ...]
    assert [int(c1[i]) for i in (5, 9, 13, 17, 6, 10)] == [21, 25, 29, 33, 37, 41]
    assert q.validateQuadTreeNodeBBox()
    # quadrants (src/quadtree.py:153-175): c1 upper-right, c2 upper-left, c3 lower-left, c4 lower-right
    assert n.bbox_min[1].tolist() == [0.5, 0.5] and n.bbox_max[1].tolist() == [1, 1]
    assert n.bbox_min[2].tolist() == [0, 0.5] and n.bbox_max[2].tolist() == [0.5, 1]
    assert n.bbox_min[3].tolist() == [0, 0] and n.bbox_max[3].tolist() == [0.5, 0.5]
    assert n.bbox_min[4].tolist() == [0.5, 0] and n.bbox_max[4].tolist() == [1, 0.5]
    # copyTree -> canonical node-major layout (src/quadtree.py:801-817)
    c = q.copyTree(np.zeros(1, U))
    assert c.getWidth() == 85
    assert [int(c.child_1_index[i]) for i in (0, 1, 2, 3, 4, 5, 6, 7)] == [1, 5, 9, 13, 17, 21, 25, 29]
    assert c.child_4_index[5] == 24
    assert q.validateQuadTreeNodeBBox(c)
    assert c.rootNodeIndex.tolist() == [0]


def test_quadtree_energy_conservation_and_refine():
    """src/quadtree.py:1205-1237: root = sum(leaves) = sum(radiance/woPdf); refine at 1%."""
    q = so.QuadTree()
    root = np.zeros(1, U)
    for _ in range(3):
        q.quadTreeNode.split(q.getAllLeafNodeIndex(root))
    rec = _records(50000, 1)
    q.quadTreeNode.irradiance64 = q.quadTreeNode.irradiance.astype(np.float64)
    q.addDataPropagate(np.zeros(50000, U), rec)
    e = q.quadTreeNode.irradiance64
    true = float(np.sum(rec.radiance.astype(np.float64) / rec.woPdf.astype(np.float64)))
    leaves = q.getAllLeafNodeIndex(root)
    assert abs(e[0] - true) < 1e-6 * true
    assert abs(e[leaves].sum() - true) < 1e-6 * true
    q.quadTreeNode.irradiance = e.astype(F)
    q.quadTreeNode.irradiance64 = None
    q.setRefinementThreshold(root, q.quadTreeNode.irradiance[:1])
    assert np.all(q.quadTreeNode.refinementThreshold == q.quadTreeNode.irradiance[0] / F(100))
    q.refine(root)
    n = q.quadTreeNode
    # uniform data over 64 leaves at 1/64 > 1% -> every depth-3 leaf splits exactly once
    leaves = q.getAllLeafNodeIndex(root)
    assert leaves.shape[0] == 256 and np.all(n.depth[leaves] == 4)
    assert q.validateQuadTreeNodeBBox()
    assert np.all(n.irradiance[leaves] <= n.refinementThreshold[leaves])
    # merge: second refine with a huge threshold collapses everything to the root
    n.refinementThreshold[:] = F(1e30)
    q.refine(root)
    q.clearTreeUnusedNode()
    assert q.quadTreeNode.getWidth() == 1 and bool(q.quadTreeNode.isLeaf[0])


def test_kd_splat_refine_conservation():
    """src/kdtree.py:738-793: counts, energy and structure after one refine."""
    t = so.KDTree(maxDepth=20)
    t.setup([0, 0, 0], [100, 100, 100])
    t.quadTree.maxDepth = 20
    n = 60000
    rec = _records(n, 2, scale=100.0)
    t.addDataPropagate(rec)
    assert t.kdTreeNode.vertCount.tolist() == [float(n)]
    t.setRefinementThreshold(0)
    assert t.maxLeafSize == 12000.0
    t.refine()
    k = t.kdTreeNode
    leaves = t.getAllLeafNodeIndex()
    # 60000 -> 30000 -> 15000 -> 7500: three rounds, 8 leaves, counts halved exactly
    assert k.getWidth() == 15 and leaves.shape[0] == 8
    assert np.all(k.vertCount[leaves] == 7500.0)
    assert float(k.vertCount[leaves].sum()) == float(n)
    assert t.validateTreeNodeBBox()
    # every leaf owns one quadtree root, roots are a permutation of 0..7
    assert sorted(k.quadTreeRootIndex[leaves].tolist()) == list(range(8))
    t.setQuadTreeRefinementThreshold()
    t.refineAllQuadTree()
    t.cleanUnusedQuadTree()
    q = t.quadTree.quadTreeNode
    assert q.rootNodeIndex.tolist() == list(range(8))
    assert t.quadTree.validateQuadTreeNodeBBox()
    # right-child copies keep the parent's energies unhalved (src/kdtree.py:316-323)
    assert np.all(q.irradiance[:8] == q.irradiance[0])
    t.resetTreeVertCount()
    t.resetAllQuadTreeIrradiance()
    assert not k.vertCount.any() and not t.quadTree.quadTreeNode.irradiance.any()


def test_out_of_box_records_splat_into_tree_zero():
    """src/kdtree.py:193,224: records outside the root box add no count but their
    energy goes to the tree owned by node 0."""
    t = so.KDTree()
    t.setup([0, 0, 0], [1, 1, 1])
    rec = so.SurfaceInteractionRecord(
        position=np.array([[2, 2, 2], [0.5, 0.5, 0.5]], F), direction=np.array([[0.1, 0.1], [0.2, 0.2]], F),
        radiance=np.array([3, 5], F), woPdf=np.array([1, 1], F))
    t.addDataPropagate(rec)
    assert t.kdTreeNode.vertCount.tolist() == [1.0]
    assert t.quadTree.quadTreeNode.irradiance.tolist() == [8.0]


def test_sample_and_pdf_semantics():
    """src/quadtree.py:931-1101: uniform tree -> pdf == 1/(4 pi); inactive lanes ->
    direction (0,0,-1), pdf 1; RNG consumption = 3 uniforms per visited node."""
    q = so.QuadTree()
    root = np.zeros(1, U)
    for _ in range(2):
        q.quadTreeNode.split(q.getAllLeafNodeIndex(root))
    n = q.quadTreeNode
    n.irradiance[0] = 16
    n.irradiance[1:5] = 4
    n.irradiance[5:21] = 1
    rng = np.random.default_rng(0)
    lanes = 1000
    u = rng.random((lanes, 9)).astype(F)
    smp = so.ExplicitSampler(u=u)
    active = np.ones(lanes, bool)
    active[::7] = False
    d, node, pos = q.sampleQuadTree(np.zeros(lanes, U), smp, active, return_node=True)
    assert np.all(smp.cursor[active] == 9) and np.all(smp.cursor[~active] == 0)
    assert np.all(d[~active] == np.array([0, 0, -1], F))
    pdf = q.pdfQuadTree(np.zeros(lanes, U), d, active)
    assert np.all(pdf[~active] == 1)
    assert np.allclose(pdf[active], 1 / (4 * np.pi), rtol=1e-6)
    # leaf position uses the LEAF-level (u_x,u_y): u[:,6], u[:,7]
    bmin = n.bbox_min[node[active]]
    assert np.all(pos[active] == bmin + u[active][:, 6:8] * F(0.25))
    # all-zero energies pick child 4 (src/quadtree.py:983-991)
    n.irradiance[:] = 0
    d, node, pos = q.sampleQuadTree(np.zeros(4, U), so.ExplicitSampler(u=u[:4]), True, return_node=True)
    assert np.all(node == n.child_4_index[n.child_4_index[0]])
    # 0/0 -> pdf 0 (src/quadtree.py:1090-1092)
    assert np.all(q.pdfQuadTree(np.zeros(4, U), d) == 0)


def test_pdf_integrates_to_one_and_sampling_matches_it():
    """Standard anchors the reference does not test (SURVEY 8c iv): on a trained tree the pdf is a density
    on the sphere -- sum over the leaves of pdf(leaf centre) x solid angle(leaf) = 1 -- and the sampler draws
    leaves with probability pdf x solid angle (chi-square on 2^17 draws with explicit uniforms)."""
    rng = np.random.default_rng(5)
    n = 30000
    d2 = np.clip(np.stack([0.3 + 0.05 * rng.standard_normal(n), 0.6 + 0.1 * rng.standard_normal(n)], 1), 0, 1).astype(F)
    d2[: n // 3] = rng.random((n // 3, 2)).astype(F)
    rec = so.SurfaceInteractionRecord(rng.random((n, 3)).astype(F), d2, (rng.integers(1, 17, n) / 8.0).astype(F),
                                      rng.choice(np.array([0.25, 0.5, 1.0, 2.0], F), n).astype(F))
    cur = so.KDTree(maxDepth=20)
    cur.setup([0, 0, 0], [1, 1, 1])
    prev = so.KDTree(maxDepth=20)
    prev.copyFrom(cur)
    for _ in range(2):
        cur.addDataPropagate(rec)
        cur.maxLeafSize = 1e9                       # one spatial leaf: a single quadtree
        cur.refine()
        cur.setQuadTreeRefinementThreshold()
        cur.refineAllQuadTree()
        cur.cleanUnusedQuadTree()
        prev.copyFrom(cur)
        cur.resetTreeVertCount()
        cur.resetAllQuadTreeIrradiance()
    q = prev.quadTree.quadTreeNode
    leaves = np.nonzero(q.isLeaf)[0]
    assert len(leaves) > 50
    centre = ((q.bbox_min[leaves] + q.bbox_max[leaves]) / F(2)).astype(F)
    area = np.prod((q.bbox_max[leaves] - q.bbox_min[leaves]).astype(np.float64), axis=1)
    pos = np.full((len(leaves), 3), 0.5, F)
    pdf = prev.pdf(pos, dm.canonical_to_dir(centre), np.ones(len(leaves), bool)).astype(np.float64)
    prob = pdf * 4 * np.pi * area                   # canonical area x 4 pi = solid angle (equal-area map, src/common.py:100-129)
    assert abs(prob.sum() - 1.0) < 1e-5
    m = 1 << 17
    u = rng.random((m, 3 * 22)).astype(F)
    _, _, dbg = prev.sample(np.full((m, 3), 0.5, F), so.ExplicitSampler(u=u), np.ones(m, bool), return_debug=True)
    obs = np.bincount(dbg['sample_node'], minlength=q.getWidth())[leaves].astype(np.float64)
    exp = prob * m
    big = exp >= 8
    chi2 = float(((obs - exp) ** 2 / np.maximum(exp, 1e-300))[big].sum())
    df = int(big.sum()) - 1
    assert abs(chi2 - df) < 5 * np.sqrt(2 * df) + 5, (chi2, df)
    assert obs[exp == 0].sum() == 0


def test_pdf_tie_rules():
    """src/quadtree.py:1063-1075 (energy: first match) vs :1095-1098 (descent: last match)."""
    q = so.QuadTree()
    root = np.zeros(1, U)
    q.quadTreeNode.split(q.getAllLeafNodeIndex(root))
    n = q.quadTreeNode
    n.irradiance[:] = np.array([10, 1, 2, 3, 4], F)
    pos = np.array([[0.5, 0.5], [0.5, 0.75], [0.25, 0.5], [0.5, 0.25], [0.75, 0.5]], F)
    d = dm.canonical_to_dir(pos)
    # make sure the round trip is exact for these dyadic test points
    assert np.array_equal(dm.dir_to_canonical(d)[:, 1], pos[:, 1])
    pdf, node, ppos = q.pdfQuadTree(np.zeros(5, U), d, True, return_node=True)
    exp_desc, exp_e = [], []
    for x, y in ppos:
        exp_desc.append((4 if x >= 0.5 else 3) if y <= 0.5 else (2 if x <= 0.5 else 1))
        exp_e.append((1 if x >= 0.5 else 2) if y >= 0.5 else (3 if x <= 0.5 else 4))
    assert node.tolist() == exp_desc
    exp = np.array([F(4) * F(e) / F(10) * dm.INV_FOUR_PI for e in exp_e], F)
    assert np.allclose(pdf, exp, rtol=1e-6)


def test_integrator_pieces():
    """src/path_guiding_integrator.py:16-24, 434-478."""
    assert so.mis_weight(np.array([0, 1, 2, np.nan], F), np.array([1, 1, 0, 1], F)).tolist() == [0, 0.5, 1, 0]
    L = np.array([[4, 4, 4]], F)
    thr_rad = np.array([[1, 1, 1], [0, 0, 0]], F)
    thr_bsdf = np.array([[1, 1, 1], [0, 0, 0]], F)
    bsdf = np.array([[0.5, 0.5, 0.5], [0, 0, 0]], F)
    prod, rad = so.process_path_data(L, thr_rad, thr_bsdf, bsdf, 2)
    assert prod[0].tolist() == [3, 3, 3]
    assert abs(rad[0] - 6.0) < 1e-5
    keep, r, rn = so.filter_records(np.array([True, True]), np.array([1, np.nan], F),
                                    np.zeros((2, 3), F), np.array([1, 1], F))
    assert keep.tolist() == [True, False]
    keep, _, _ = so.filter_records(np.array([True, True, False]), np.array([1, 1, 1], F),
                                   np.zeros((3, 3), F), np.array([0, np.nan, 1], F))
    assert keep.tolist() == [False, False, False]


def test_npz_roundtrip(tmp_path):
    t = so.KDTree(maxDepth=20)
    t.setup([0, 0, 0], [1, 1, 1])
    t.addDataPropagate(_records(40000, 5))
    prev = so.KDTree()
    so.refine_and_prepare(t, prev, 0)
    f = str(tmp_path / 'tree.npz')
    prev.saveToFile(f)
    d = np.load(f)
    assert set(d.files) == set(so.KDTree.NPZ_KEYS)
    t2 = so.KDTree()
    t2.loadFromFile(f)
    for k, v in prev.to_arrays().items():
        if k == 'kdtree_maxLeafSize':
            continue
        assert np.array_equal(np.asarray(v), np.asarray(t2.to_arrays()[k])), k


def test_oracle_reproduces_committed_golden_file():
    """tests/golden/sdtree_golden.npz (made by tests/golden/make_sdtree_golden.py): the oracle, reloaded
    from the stored tree arrays, gives the stored answers -- any later change of the oracle is caught"""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sdtree_golden.npz"))
    t = so.KDTree()
    t.loadFromArrays({k[5:]: g[k] for k in g.files if k.startswith("tree_")})
    a = g['active']
    assert np.array_equal(t.getLeafNodeIndex(g['pos'], a), g['leaf'])
    d, p, dbg = t.sample(g['pos'], so.ExplicitSampler(u=g['u']), a, return_debug=True)
    assert np.array_equal(dbg['sample_node'][a], g['sample_node'][a])
    assert np.array_equal(d.view(U), g['sample_dir'].view(U)) and np.array_equal(p.view(U), g['sample_pdf'].view(U))
    pp = t.pdf(g['pos'], g['dirs'], a)
    assert np.array_equal(np.isnan(pp), np.isnan(g['pdf'])) and np.array_equal(pp[~np.isnan(pp)], g['pdf'][~np.isnan(pp)])
