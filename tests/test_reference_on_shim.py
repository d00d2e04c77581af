"""Pins oracle/sdtree_oracle.py to the reference's OWN source: /root/reference/src/{common,quadtree,
kdtree,path_guiding_integrator}.py are imported unmodified on the numpy Dr.Jit / Mitsuba stand-ins of
oracle/refshim/ and driven through (a) the hand-derived vectors of SURVEY 8a, (b) the reference's
self-test invariants (src/quadtree.py:1205-1218, src/kdtree.py:769-772), (c) the same splat / refine /
query sequences as the oracle -- all 23 arrays of both trees and every query output must be identical
bit for bit -- and (d) every parity case and fuzz seed of the suites with the reference itself as the
expected-value leg (the library under test there is the host emulation of the kernels).

Skipped where /root/reference does not exist (the GPU box); tests/golden/reference_on_shim.npz carries
outputs of this leg there (tests/golden/make_reference_golden.py)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import refshim  # noqa: E402

pytestmark = pytest.mark.skipif(not refshim.available(), reason="reference tree not present")

if refshim.available():
    from oracle.refshim import as_oracle as ro
    from oracle import sdtree_oracle as so
    import sdt_cases as cases
    import fuzz_cases

F, U = np.float32, np.uint32


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        return False
    if a.dtype.kind == 'f':
        return cases.beq(a, b)
    return bool(np.array_equal(a, b))


def assert_same_tree(o, r, what=''):
    ao, ar = o.to_arrays(), r.to_arrays()
    for k in so.KDTree.NPZ_KEYS:
        assert same(ao[k], ar[k]), f'{what}: {k}'


def pair(mod, lo, hi, kd, qd, nee):
    cur = mod.KDTree(maxDepth=kd)
    cur.setup(lo, hi)
    cur.quadTree.maxDepth = qd
    cur.quadTree.isStoreNEERadiance = nee
    prev = mod.KDTree(maxDepth=kd)
    prev.copyFrom(cur)
    return cur, prev


# ------------------------------------------------------------------------------------------- (a)
def test_reference_reproduces_hand_derived_vectors():
    """SURVEY 8a golden orderings, from the reference's own split / copyTree"""
    ref = refshim.load_reference()
    mi = ref.mi
    t = ref.kdtree.KDTree()
    t.setup(bbox_min=[0, 0, 0], bbox_max=[100, 100, 100])
    t.split(t.getAllLeafNodeIndex())
    t.split(t.getAllLeafNodeIndex())
    k = t.kdTreeNode
    assert k.child_left_index.numpy().tolist() == [1, 3, 5, 0, 0, 0, 0]
    assert k.child_right_index.numpy().tolist() == [2, 4, 6, 0, 0, 0, 0]
    assert k.depth.numpy().tolist() == [0, 1, 1, 2, 2, 2, 2]
    assert k.quadTreeRootIndex.numpy().tolist() == [0, 0, 1, 0, 2, 1, 3]
    assert k.bbox.min.numpy()[5].tolist() == [50, 0, 0] and k.bbox.max.numpy()[5].tolist() == [100, 50, 100]
    assert t.getLeafNodeIndex(mi.Vector3f(75, 25, 25)).numpy().tolist() == [5]
    assert t.validateTreeNodeBBox()
    q = ref.quadtree.QuadTree()
    for _ in range(3):
        q.quadTreeNode.split(q.getAllLeafNodeIndex(rootIndex=mi.UInt32(0)))
    assert q.quadTreeNode.getWidth() == 85
    assert q.quadTreeNode.child_1_index.numpy()[[5, 9, 13, 17, 6]].tolist() == [21, 25, 29, 33, 37]
    c = q.copyTree(rootIndex=mi.UInt32(0))
    assert c.child_1_index.numpy()[[5, 6, 7]].tolist() == [21, 25, 29]
    assert q.validateQuadTreeNodeBBox(c) and not ref.dr.printf_calls


# ------------------------------------------------------------------------------------------- (b)
def test_reference_self_test_invariants():
    """the sequence of the reference's __main__ blocks at a size a CPU test affords, with the three
    numbers it prints side by side asserted: root energy == sum over leaves == sum radiance / woPdf,
    and sum of leaf counts == records in the box"""
    ref = refshim.load_reference()
    dr, mi = ref.dr, ref.mi
    rng = np.random.default_rng(5)
    n = 50000
    rec = so.SurfaceInteractionRecord((rng.random((n, 3)) * 100).astype(F), rng.random((n, 2)).astype(F),
                                      (rng.integers(0, 17, n) / 8.0).astype(F), rng.choice(np.array([0.25, 0.5, 1, 2], F), n))
    # src/quadtree.py:1136-1218
    q = ref.quadtree.QuadTree()
    root = mi.UInt32(0)
    for _ in range(3):
        q.quadTreeNode.split(q.getAllLeafNodeIndex(rootIndex=root))
    q.addDataPropagate(dr.full(mi.UInt32, 0, n), ro.to_record(rec))
    e = q.quadTreeNode.irradiance.numpy()
    leaves = q.getAllLeafNodeIndex(root).numpy()
    true_sum = float(np.sum(rec.radiance.astype(np.float64) / rec.woPdf))
    assert e[0] == e[leaves].sum() == true_sum
    q.setRefinementThreshold(rootIndex=root, total_flux_prev_quadtree=mi.Float(true_sum))
    q.refine(root)
    leaves = q.getAllLeafNodeIndex(root).numpy()
    e = q.quadTreeNode.irradiance.numpy()
    assert e[leaves].sum() == true_sum and q.validateQuadTreeNodeBBox(q.quadTreeNode)
    thr = F(true_sum) / F(100)
    d = q.quadTreeNode.depth.numpy()
    assert np.all((e[leaves] <= thr) | (d[leaves] == q.maxDepth))
    # src/kdtree.py:699-785
    t = ref.kdtree.KDTree()
    t.setup(bbox_min=[0, 0, 0], bbox_max=[100, 100, 100])
    t.split(t.getAllLeafNodeIndex())
    t.split(t.getAllLeafNodeIndex())
    t.addDataPropagate(ro.to_record(rec))
    t.maxLeafSize = 700.0
    t.refine()
    t.setQuadTreeRefinementThreshold()
    t.refineAllQuadTree()
    t.cleanUnusedQuadTree()
    leaf = t.getAllLeafNodeIndex().numpy()
    vc = t.kdTreeNode.vertCount.numpy()
    assert vc[leaf].sum() == n and vc[0] == n and t.validateTreeNodeBBox()
    assert np.all((vc[leaf] <= 700) | (t.kdTreeNode.depth.numpy()[leaf] == t.maxDepth))
    assert t.getLeafNodeIndex(mi.Vector3f(75, 25, 25)).numpy()[0] in leaf


def test_reference_main_blocks_run_unmodified():
    """the reference's three print-only self-tests (src/quadtree.py:1105-1436, src/kdtree.py:667-835,
    src/common.py:270-307: 1 M unseeded records each) executed as they are; the numbers they print side by
    side are asserted here"""
    g, out = refshim.run_reference_main('quadtree')
    q = g['myTree'].quadTreeNode
    assert q.getWidth() > 85 and len(q.rootNodeIndex.numpy()) == 4
    assert g['myTree'].validateQuadTreeNodeBBox(q) and 'result: False' not in out
    g, out = refshim.run_reference_main('kdtree')
    t = g['myTree']
    assert out.count('bounding box result:\x1b[0m True') == 2 and t.validateTreeNodeBBox()
    assert 'vertCount sum after refine:\x1b[0m 1e+06' in out           # src/kdtree.py:769-772: N = 1e6 records kept
    leaf = t.getLeafNodeIndex(g['position']).numpy()[0]
    assert bool(t.kdTreeNode.isLeaf.numpy()[leaf])
    assert abs(float(np.linalg.norm(g['direction'].numpy())) - 1) < 1e-5 and g['pdf'].numpy()[0] > 0
    g, out = refshim.run_reference_main('common')
    assert 'UInt32([0, 1, 2, 0, 3, 6, 9, 12])' in out and 'UInt32([5, 6])' in out


# ------------------------------------------------------------------------------------------- (c)
def general_records(rng, n, lo, hi, nee, dyadic):
    ext = np.asarray(hi, F) - np.asarray(lo, F)
    pos = (np.asarray(lo, F) + rng.random((n, 3)) ** 1.5 * ext).astype(F)
    d = rng.random((n, 2)).astype(F)
    m = rng.random(n) < 0.6
    d[m] = np.clip(np.stack([0.3 + 0.01 * rng.standard_normal(m.sum()), 0.7 + 0.003 * rng.standard_normal(m.sum())], 1), 0, 1).astype(F)
    if dyadic:
        rad = (rng.integers(0, 17, n) / 8.0).astype(F)
        wo = rng.choice(np.array([0.25, 0.5, 1.0, 2.0], F), n).astype(F)
    else:
        rad = rng.lognormal(0, 1, n).astype(F)
        wo = rng.uniform(0.05, 2, n).astype(F)
    rec = so.SurfaceInteractionRecord(pos, d, rad, wo)
    if nee:
        rec.radiance_nee = (rng.random((n, 3)) * (rng.random((n, 1)) < 0.5)).astype(F)     # non-zero NEE energy
        rec.direction_nee = rng.random((n, 2)).astype(F)
    # the inputs the reference's masks exist for
    pos[:3] = [np.asarray(lo, F), np.asarray(hi, F), np.asarray(lo, F) + ext * F(0.5)]
    pos[3] = np.asarray(hi, F) + ext
    pos[4, 1] = np.nan
    d[5] = [0.5, 0.5]
    d[6] = [1.0, 0.0]
    d[7] = [1.5, 0.2]
    d[8, 0] = np.nan
    wo[9], wo[10], wo[11] = 0.0, -1.0, np.nan
    return rec


@pytest.mark.parametrize("cfg", [
    dict(lo=(0, 0, 0), hi=(1, 1, 1), kd=20, qd=20, nee=False, dyadic=True, leaf=300),
    dict(lo=(0, 0, 0), hi=(1, 1, 1), kd=20, qd=20, nee=True, dyadic=False, leaf=300),
    dict(lo=(-3.5, 0.25, -1), hi=(2.25, 7, 0.5), kd=6, qd=5, nee=True, dyadic=False, leaf=50),
    dict(lo=(0, 0, 0), hi=(100, 100, 100), kd=3, qd=20, nee=False, dyadic=False, leaf=1),
], ids=lambda c: f"kd{c['kd']}q{c['qd']}{'nee' if c['nee'] else ''}{'dy' if c['dyadic'] else 'fp'}")
@pytest.mark.parametrize("recip", [False, True], ids=["ieee_div", "recip_div"])
def test_oracle_equals_reference_train_and_query(cfg, recip):
    """identical records into the oracle and into the reference: statistics after every splat, all 23
    arrays of both trees after every refine, then leaf / root / node ids, directions and pdfs"""
    rng = np.random.default_rng(11)
    lo, hi = cfg['lo'], cfg['hi']
    ocur, oprev = pair(so, lo, hi, cfg['kd'], cfg['qd'], cfg['nee'])
    rcur, rprev = pair(ro, lo, hi, cfg['kd'], cfg['qd'], cfg['nee'])
    so.QUAD_THR_RECIPROCAL = ro.QUAD_THR_RECIPROCAL = recip
    try:
        for it in range(4):
            rec = general_records(rng, 6000, lo, hi, cfg['nee'], cfg['dyadic'])
            ocur.addDataPropagate(rec)
            rcur.addDataPropagate(rec)
            assert_same_tree(ocur, rcur, f'after splat {it}')
            if it == 3:
                break                    # the last iteration stays unrefined, as in main.py:371-377
            if it == 1:                  # the integrator's own entry point, threshold 12000*sqrt(2^it)
                so.refine_and_prepare(ocur, oprev, -8)
                ro.refine_and_prepare(rcur, rprev, -8)
            else:
                cases.oracle_refine(ocur, oprev, cfg['leaf'])
                cases.oracle_refine(rcur, rprev, cfg['leaf'])
            assert_same_tree(ocur, rcur, f'current after refine {it}')
            assert_same_tree(oprev, rprev, f'prev after refine {it}')
            assert rprev.validateTreeNodeBBox() and rprev.quadTree.validateQuadTreeNodeBBox()
    finally:
        so.QUAD_THR_RECIPROCAL = ro.QUAD_THR_RECIPROCAL = False
    assert oprev.kdTreeNode.getWidth() > 3 or cfg['kd'] < 2
    # queries on the frozen tree
    n = 3000
    ext = np.asarray(hi, F) - np.asarray(lo, F)
    pos = (np.asarray(lo, F) - 0.01 * ext + rng.random((n, 3)) * ext * 1.02).astype(F)
    pos[:4] = [np.asarray(lo, F), np.asarray(hi, F), np.asarray(lo, F) + ext * F(0.5), np.asarray(lo, F) + ext * F(0.25)]
    pos[4, 2] = np.nan
    active = rng.random(n) < 0.9
    assert same(oprev.getLeafNodeIndex(pos, active), rprev.getLeafNodeIndex(pos, active))
    u = rng.random((n, 3 * (cfg['qd'] + 2))).astype(F)
    u[:64] = rng.integers(0, 5, (64, u.shape[1])) / F(4)              # bin-edge uniforms, 0 and 1 included
    od, op, odbg = oprev.sample(pos, so.ExplicitSampler(u=u), active, return_debug=True)
    rd, rp, rdbg = rprev.sample(pos, ro.ExplicitSampler(u=u), active, return_debug=True)
    for k in ('leaf', 'root', 'sample_node', 'sample_pos', 'pdf_node', 'pdf_pos'):
        assert same(odbg[k], rdbg[k]), k
    assert same(od, rd) and same(op, rp)
    od2, op2 = oprev.sample(pos, so.ExplicitSampler(seed=9, n=n, lane_offset=5), active)
    rd2, rp2 = rprev.sample(pos, ro.ExplicitSampler(seed=9, n=n, lane_offset=5), active)
    assert same(od2, rd2) and same(op2, rp2)
    dirs = rng.standard_normal((n, 3)).astype(F)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    from oracle import drjit_math as dm
    grid = (rng.integers(0, 9, (200, 2)) / 8.0).astype(F)
    dirs[:200] = dm.canonical_to_dir(grid)                            # on quadrant borders (tie rules)
    dirs[200] = [0, 0, 1]
    dirs[201] = [0, 0, -1]
    dirs[202] = [np.nan, 0, 1]
    dirs[203] = [0, 0, 0]
    dirs[204] = [np.inf, 0, 0]
    opp, opdbg = oprev.pdf(pos, dirs, active, return_debug=True)
    rpp, rpdbg = rprev.pdf(pos, dirs, active, return_debug=True)
    for k in ('leaf', 'root', 'pdf_node', 'pdf_pos'):
        assert same(opdbg[k], rpdbg[k]), k
    assert same(opp, rpp)


def test_oracle_equals_reference_integrator_pieces():
    """mis_weight, processPathData, the scatterDataIntoSDTree filter: the reference's methods themselves"""
    rng = np.random.default_rng(3)
    n, md = 4096, 4
    a = rng.random(n).astype(F) * (rng.random(n) < 0.9)
    b = rng.random(n).astype(F) * (rng.random(n) < 0.9)
    a[:3], b[:3] = [0, np.nan, np.inf], [0, 1, np.inf]
    assert same(so.mis_weight(a, b), ro.mis_weight(a, b))
    slots = n
    Lf = rng.random((slots // md, 3)).astype(F) * 4
    tr = (rng.random((slots, 3)) * 2).astype(F)
    tb = (rng.random((slots, 3)) * (rng.random((slots, 1)) < 0.9)).astype(F)
    bs = (rng.random((slots, 3)) * (rng.random((slots, 3)) < 0.95)).astype(F)
    tr[:2], tb[2] = np.nan, np.inf
    oo, orad = so.process_path_data(Lf, tr, tb, bs, md)
    r_o, rrad = ro.process_path_data(Lf, tr, tb, bs, md)
    assert same(oo, r_o) and same(orad, rrad)
    nee = (rng.random((slots, 3)) * (rng.random((slots, 1)) < 0.3)).astype(F)
    wo = rng.random(slots).astype(F)
    wo[:64] = rng.choice(np.array([0, -1, np.nan, np.inf], F), 64)
    orad[64:80] = np.nan
    nee[80:90, 1] = np.nan
    active = rng.random(slots) < 0.7
    ok, orad2, onee = so.filter_records(active, orad, nee, wo)
    rk, rrad2, rnee = ro.filter_records(active, orad, nee, wo)
    assert same(ok, rk) and same(orad2, rrad2) and same(onee, rnee) and 0 < ok.sum() < slots


# ------------------------------------------------------------------------------------------- (d)
@pytest.fixture(scope="module")
def ref_ctx():
    from hostemu.build_hostemu import build as build_hostemu
    from practical_path_guiding_lab_b200 import SDTree
    lib = build_hostemu()
    return cases.Ctx(make=lambda **kw: SDTree(lib_path=lib, **kw))


@pytest.fixture()
def reference_as_expected(monkeypatch):
    """the parity cases compute their expected values with the reference's own source"""
    monkeypatch.setattr(cases, 'so', ro)
    monkeypatch.setattr(fuzz_cases, 'so', ro)


@pytest.mark.parametrize("case", (cases.ALL_CASES + fuzz_cases.SUITE_CASES) if refshim.available() else [],
                         ids=lambda c: c.__name__)
def test_parity_case_against_reference(ref_ctx, reference_as_expected, case):
    case(ref_ctx)


@pytest.mark.parametrize("seed", range(1000, 1040))
def test_fuzz_seed_against_reference(ref_ctx, reference_as_expected, seed):
    fuzz_cases.fuzz_one(ref_ctx, seed)
