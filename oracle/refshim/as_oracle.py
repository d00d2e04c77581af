"""ORACLE SUPPORT (test infrastructure only -- never imported by the product path).

The REFERENCE'S OWN SOURCE (run on the numpy Dr.Jit / Mitsuba stand-ins of this directory) behind
the same numpy-in / numpy-out surface as oracle/sdtree_oracle.py, so that every parity case written
against the oracle can be replayed with the reference itself as the expected-value leg:

    from oracle.refshim import as_oracle as ro         # needs /root/reference
    cur = ro.KDTree(maxDepth=20); cur.setup(lo, hi); cur.addDataPropagate(rec); cur.refine() ...

Each method converts its numpy arguments to shim arrays, calls the reference method of the same
name (file:line given), and converts the result back.  What is NOT the reference here is named in
place: the perf-mode counter generator (a property of this repo, served to the reference as an
explicit table), the `exact=` float64 accumulation switch and the `return_debug` outputs (obtained
by calling the reference's own sub-steps in the order KDTree.sample / KDTree.pdf call them).
"""
import numpy as np

from oracle import sdtree_oracle as so
from oracle.refshim import load_reference

F, U = np.float32, np.uint32

_ref = load_reference()
if _ref is None:                                          # pragma: no cover
    raise ImportError('reference tree not present (set SDT_REFERENCE_ROOT)')
dr, mi = _ref.dr, _ref.mi
R_KDTree, R_QuadTree = _ref.kdtree.KDTree, _ref.quadtree.QuadTree
R_Record = _ref.common.SurfaceInteractionRecord

# helpers of the test cases that are not reference code
gather, bbox_contains, counter_uniform = so.gather, so.bbox_contains, so.counter_uniform
SurfaceInteractionRecord = so.SurfaceInteractionRecord
nee_mis, mixture = so.nee_mis, so.mixture        # inline code of sample(); pinned by the stub-scene run instead

QUAD_THR_RECIPROCAL = False                      # mirrored into the shim's scalar-division mode per call


def _sync_div_mode():
    dr.SCALAR_DIV_RECIPROCAL = bool(QUAD_THR_RECIPROCAL)


def to_record(rec):
    """numpy record container -> the reference's SurfaceInteractionRecord (src/common.py:14-40)"""
    n = rec.position.shape[0]
    r = dr.zeros(R_Record, n)
    r.position = mi.Vector3f(rec.position.reshape(n, 3))
    r.direction = mi.Vector2f(rec.direction.reshape(n, 2))
    r.radiance = mi.Float(rec.radiance)
    r.woPdf = mi.Float(rec.woPdf)
    r.radiance_nee = mi.Color3f(rec.radiance_nee.reshape(n, 3))
    r.direction_nee = mi.Vector2f(rec.direction_nee.reshape(n, 2))
    r.active = mi.Bool(np.ones(n, bool))
    return r


def _w(x, n):
    """numpy value of a shim array, width-1 results broadcast to n lanes"""
    v = x.numpy()
    return np.broadcast_to(v, (n,) + v.shape[1:]).copy() if v.shape[0] == 1 and n != 1 else v


def _mask(active, n):
    return mi.Bool(np.broadcast_to(np.asarray(active, bool), (n,)).copy())


class ExplicitSampler:
    """explicit table, or the repo's counter generator tabulated for the reference to consume"""

    def __init__(self, u=None, seed=None, n=None, lane_offset=0, levels=64):
        if u is None:
            lane = (np.arange(n, dtype=np.uint64) + lane_offset).astype(U)
            idx = np.arange(3 * levels, dtype=U)
            u = so.counter_uniform(seed, lane[:, None], idx[None, :])
        self.shim = mi.ExplicitSampler(np.asarray(u, F))

    @property
    def cursor(self):
        return self.shim.cursor.numpy().astype(np.int64)


class _WriteBack(np.ndarray):
    """numpy value of a flat shim array whose item assignments are stored back into the shim array"""
    _sink = None

    def __setitem__(self, key, value):
        super().__setitem__(key, value)
        if self._sink is not None:
            self._sink.d = np.array(self, dtype=self._sink.DT)


class _NodeView:
    """a reference node store (KDTreeNode / QuadTreeNode) under the oracle's attribute names, numpy-valued"""

    def __init__(self, get):
        object.__setattr__(self, '_get', get)

    def __getattr__(self, name):
        node = self._get()
        if name in ('bbox_min', 'bbox_max'):
            return getattr(node.bbox, name[5:]).numpy().astype(F)
        v = getattr(node, name)
        if isinstance(v, dr._Flat):
            out = v.numpy().view(_WriteBack)
            out._sink = v
            return out
        return v

    def __setattr__(self, name, value):
        node = self._get()
        cur = getattr(node, name)
        setattr(node, name, type(cur)(np.asarray(value)))

    def getWidth(self):
        return self._get().getWidth()

    def split(self, idx):                                     # QuadTreeNode.split, src/quadtree.py:96-191
        self._get().split(mi.UInt32(np.asarray(idx, U)))

    def merge(self, idx):                                     # :194-213
        self._get().merge(mi.UInt32(np.asarray(idx, U)))

    def getAllLeafNodeIndex(self, rootIndex=None):            # :288-345
        r = None if rootIndex is None else mi.UInt32(np.asarray(rootIndex, U))
        return self._get().getAllLeafNodeIndex(r).numpy()


class QuadTree:
    def __init__(self, r):
        self.r = r
        self.quadTreeNode = _NodeView(lambda: self.r.quadTreeNode)

    maxDepth = property(lambda s: s.r.maxDepth, lambda s, v: setattr(s.r, 'maxDepth', v))
    isStoreNEERadiance = property(lambda s: s.r.isStoreNEERadiance, lambda s, v: setattr(s.r, 'isStoreNEERadiance', v))

    def validateQuadTreeNodeBBox(self):                       # src/quadtree.py:468-509
        return self.r.validateQuadTreeNodeBBox(self.r.quadTreeNode)


class KDTree:
    NPZ_KEYS = so.KDTree.NPZ_KEYS

    def __init__(self, max_leaf_size=1, maxDepth=10):         # src/kdtree.py:117-130
        self.r = R_KDTree(max_leaf_size, maxDepth)
        self.quadTree = QuadTree(self.r.quadTree)
        self.kdTreeNode = _NodeView(lambda: self.r.kdTreeNode)

    maxLeafSize = property(lambda s: s.r.maxLeafSize, lambda s, v: setattr(s.r, 'maxLeafSize', v))
    maxDepth = property(lambda s: s.r.maxDepth, lambda s, v: setattr(s.r, 'maxDepth', v))

    def _rebind(self):
        self.quadTree.r = self.r.quadTree

    def setup(self, bbox_min, bbox_max):                      # :133-138
        self.r.setup([float(F(x)) for x in bbox_min], [float(F(x)) for x in bbox_max])

    def copyFrom(self, o):                                    # :141-153
        self.r.copyFrom(o.r)

    def getAllLeafNodeIndex(self):                            # :173-177
        return self.r.getAllLeafNodeIndex().numpy()

    def addDataPropagate(self, rec, exact=False):             # :180-225
        if rec.position.shape[0] == 0:
            return            # scatterDataIntoSDTree returns before the call (src/path_guiding_integrator.py:482)
        old = dr._scatter_flat
        if exact:             # NOT reference behaviour: float64 accumulation, one rounding per call and node
            def exact_scatter(target, value, index, active, reduce_add):
                if not reduce_add or target.DT is not F:
                    return old(target, value, index, active, reduce_add)
                t64 = _F64(target.d.astype(np.float64))
                v = value.d if hasattr(value, 'd') else np.asarray(value)
                old(t64, _F64(np.asarray(v, np.float64).reshape(-1)), index, active, True)
                target.d = t64.d.astype(F)
            dr._scatter_flat = exact_scatter
        try:
            self.r.addDataPropagate(to_record(rec))
        finally:
            dr._scatter_flat = old

    def split(self, idx):                                     # :229-323
        self.r.split(mi.UInt32(np.asarray(idx, U)))

    def setRefinementThreshold(self, iteration):              # :327-330
        self.r.setRefinementThreshold(iteration)

    def refine(self):                                         # :333-358
        self.r.refine()

    def validateTreeNodeBBox(self):                           # :361-398
        return self.r.validateTreeNodeBBox()

    def resetTreeVertCount(self):                             # :401-432
        self.r.resetTreeVertCount()

    def getLeafNodeIndex(self, position, active=True):        # :435-470
        position = np.asarray(position, F)
        n = position.shape[0]
        return self.r.getLeafNodeIndex(mi.Vector3f(position), _mask(active, n)).numpy()

    def sample(self, position, sampler, active=True, return_debug=False):      # :473-486
        position = np.asarray(position, F)
        n = position.shape[0]
        p, a = mi.Vector3f(position), _mask(active, n)
        if not return_debug:
            d, pdf = self.r.sample(p, sampler.shim, a)
            return d.numpy(), pdf.numpy()
        # the four statements of KDTree.sample, with the node reached read back after each descent
        leaf = self.r.getLeafNodeIndex(p, a)
        root = dr.gather(mi.UInt32, self.r.kdTreeNode.quadTreeRootIndex, leaf, a)
        d = self.r.quadTree.sampleQuadTree(root, sampler.shim, a)
        snode, _, spos, _ = mi.Loop.finished['Sample QuadTree']          # loop state, src/quadtree.py:944
        pdf = self.r.quadTree.pdfQuadTree(root, d, a)
        pnode = mi.Loop.finished['PDF QuadTree'][0]                      # :1018
        ppos = _ref.common.dirToCanonical(d)                             # :1016
        dbg = dict(leaf=leaf.numpy(), root=root.numpy(), sample_node=_w(snode, n), sample_pos=_w(spos, n),
                   pdf_node=_w(pnode, n), pdf_pos=_w(ppos, n))
        return d.numpy(), pdf.numpy(), dbg

    def pdf(self, position, direction, active=True, return_debug=False):       # :489-496
        position = np.asarray(position, F)
        n = position.shape[0]
        p, a = mi.Vector3f(position), _mask(active, n)
        dv = mi.Vector3f(np.asarray(direction, F))
        out = self.r.pdf(p, dv, a).numpy()
        if return_debug:
            pnode = mi.Loop.finished['PDF QuadTree'][0]
            leaf = self.r.getLeafNodeIndex(p, a)
            root = dr.gather(mi.UInt32, self.r.kdTreeNode.quadTreeRootIndex, leaf, a)
            return out, dict(leaf=leaf.numpy(), root=root.numpy(), pdf_node=_w(pnode, n),
                             pdf_pos=_w(_ref.common.dirToCanonical(dv), n))
        return out

    def setQuadTreeRefinementThreshold(self):                 # :503-514
        _sync_div_mode()
        try:
            self.r.setQuadTreeRefinementThreshold()
        finally:
            dr.SCALAR_DIV_RECIPROCAL = False

    def refineAllQuadTree(self):                              # :517-524
        self.r.refineAllQuadTree()

    def cleanUnusedQuadTree(self):                            # :527-528
        self.r.cleanUnusedQuadTree()

    def resetAllQuadTreeIrradiance(self):                     # :531-532
        self.r.resetAllQuadTreeIrradiance()

    def to_arrays(self):                                      # the 23 keys of saveToFile, :539-602
        k, q = self.r.kdTreeNode, self.r.quadTree.quadTreeNode
        return dict(
            kdtree_maxLeafSize=np.asarray(self.r.maxLeafSize), kdtree_maxDepth=np.asarray(self.r.maxDepth),
            kdtree_bbox_min=k.bbox.min.numpy(), kdtree_bbox_max=k.bbox.max.numpy(),
            kdtree_depth=k.depth.numpy(), kdtree_vertCount=k.vertCount.numpy(), kdtree_isLeaf=k.isLeaf.numpy(),
            kdtree_quadTreeRootIndex=k.quadTreeRootIndex.numpy(),
            kdtree_child_left_index=k.child_left_index.numpy(), kdtree_child_right_index=k.child_right_index.numpy(),
            quadtree_maxDepth=np.asarray(self.r.quadTree.maxDepth),
            quadtree_isStoreNEERadiance=np.asarray(self.r.quadTree.isStoreNEERadiance),
            quadtree_rootNodeIndex=q.rootNodeIndex.numpy(),
            quadtree_bbox_min=q.bbox.min.numpy(), quadtree_bbox_max=q.bbox.max.numpy(),
            quadtree_depth=q.depth.numpy(), quadtree_irradiance=q.irradiance.numpy(), quadtree_isLeaf=q.isLeaf.numpy(),
            quadtree_refinementThreshold=q.refinementThreshold.numpy(),
            quadtree_child_1_index=q.child_1_index.numpy(), quadtree_child_2_index=q.child_2_index.numpy(),
            quadtree_child_3_index=q.child_3_index.numpy(), quadtree_child_4_index=q.child_4_index.numpy())

    def saveToFile(self, fileName):                           # :539-602, the reference's own writer
        self.r.saveToFile(fileName)

    def loadFromArrays(self, d):                              # :156-170, the reference's own reader
        self.r.loadFromFile(d)
        self._rebind()

    def loadFromFile(self, fileName):
        self.loadFromArrays(np.load(fileName))


class _F64(dr._Flat):
    DT = np.float64

    def __init__(self, d):
        self.d = d


# ------------------------------------------------------------------------- integrator pieces
def mis_weight(pdf_a, pdf_b):                                 # src/path_guiding_integrator.py:16-24
    return _ref.integrator.mis_weight(mi.Float(np.asarray(pdf_a, F)), mi.Float(np.asarray(pdf_b, F))).numpy()


def make_integrator(max_depth, num_rays, bbox_min=(0, 0, 0), bbox_max=(1, 1, 1), **setup):
    """PathGuidingIntegrator(props) + setup(), as main.py:49-64 does"""
    integ = _ref.integrator.PathGuidingIntegrator(mi.Properties(max_depth=max_depth))
    integ.setup(num_rays, [float(x) for x in bbox_min], [float(x) for x in bbox_max], **setup)
    return integ


def process_path_data(Lfinal, throughputRadiance, throughputBsdf, bsdf, max_depth):
    """PathGuidingIntegrator.processPathData, src/path_guiding_integrator.py:434-453"""
    slots = throughputRadiance.shape[0]
    integ = make_integrator(max_depth, slots // max_depth)
    r = integ.surfaceInteractionRecord
    r.throughputRadiance = mi.Color3f(np.asarray(throughputRadiance, F))
    r.throughputBsdf = mi.Color3f(np.asarray(throughputBsdf, F))
    r.bsdf = mi.Color3f(np.asarray(bsdf, F))
    integ.processPathData(mi.Color3f(np.asarray(Lfinal, F)))
    return r.product.numpy(), r.radiance.numpy()


def filter_records(active, radiance, radiance_nee, woPdf):
    """the filter + compaction of PathGuidingIntegrator.scatterDataIntoSDTree (:456-500), run on the
    reference with a capturing tree; the surviving lanes are recovered from a lane id carried in
    `bsdfPdf`, a field the reference compacts alongside (:496) and never reads"""
    n = radiance.shape[0]
    integ = make_integrator(1, n)
    r = integ.surfaceInteractionRecord
    r.active = mi.Bool(np.asarray(active, bool))
    r.radiance = mi.Float(np.asarray(radiance, F))
    r.radiance_nee = mi.Color3f(np.asarray(radiance_nee, F))
    r.woPdf = mi.Float(np.asarray(woPdf, F))
    r.bsdfPdf = mi.Float(np.arange(n, dtype=F))
    assert n < (1 << 24)
    got = {}

    class Capture:
        def addDataPropagate(self, rec):
            got['rec'] = rec
    integ.sdTree_current = Capture()
    integ.scatterDataIntoSDTree()
    keep = np.zeros(n, bool)
    rad = np.asarray(radiance, F).copy()
    nee = np.asarray(radiance_nee, F).copy()
    rad[np.isnan(rad)] = 0
    nee[np.isnan(nee)] = 0
    if 'rec' in got:
        ids = got['rec'].bsdfPdf.numpy().astype(np.int64)
        keep[ids] = True
        assert np.array_equal(got['rec'].radiance.numpy().view(U), rad[ids].view(U))
        assert np.array_equal(got['rec'].radiance_nee.numpy().view(U), nee[ids].view(U))
    return keep, rad, nee


def refine_and_prepare(current, prev, iteration):
    """PathGuidingIntegrator.refineAndPrepareSDTreeForNextIteration (:566-586) on the given pair"""
    integ = _ref.integrator.PathGuidingIntegrator(mi.Properties(max_depth=1))
    integ.sdTree_current, integ.sdTree_prev, integ.iteration = current.r, prev.r, iteration
    _sync_div_mode()
    try:
        integ.refineAndPrepareSDTreeForNextIteration()
    finally:
        dr.SCALAR_DIV_RECIPROCAL = False
