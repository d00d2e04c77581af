"""ORACLE (test infrastructure only -- never imported by the product path).

Literal numpy restatement of the reference's SD-tree (spatial binary tree +
directional quadtree forest): /root/reference/src/kdtree.py, src/quadtree.py,
src/common.py and the tree-facing parts of src/path_guiding_integrator.py.
Every method cites the reference lines it follows.  The code is deliberately
"wavefront literal": one numpy array op per Dr.Jit array op (gather / scatter /
scatter_reduce / compress / masked assignment), host `while` loops where the
reference has host loops, lane-masked loops where it has recorded `mi.Loop`s.

PINNED AGAINST THE REFERENCE'S OWN SOURCE (round 2): the reference ships no tests, golden vectors or saved
trees for this path and Mitsuba 3 + Dr.Jit are not installable here, but its
source files run UNMODIFIED on the numpy stand-ins for Dr.Jit / Mitsuba in oracle/refshim/
(load_reference()).  tests/test_reference_on_shim.py holds this oracle against that leg bit for bit:
every parity case of tests/sdt_cases.py, 40 randomised differential seeds, train / query sequences,
the reference's own __main__ self-tests and validators (src/kdtree.py:361-398,769-772;
src/quadtree.py:468-509,1205-1218).  What that does NOT pin is the semantics of the Dr.Jit
primitives themselves (listed below, and the CEPHES sincos / atan2 of oracle/drjit_math.py): they
are assumptions of the stand-ins.  tools/record_reference_fixtures.py records the same vectors from
a real Mitsuba installation when one is at hand.  Further anchors: (i) topology vectors hand-derived
from the reference source (tests/test_oracle_golden.py), (ii) conservation checks.

Third-party semantics assumed (Dr.Jit 0.4.x / Mitsuba 3.0-3.5):
  * BoundingBox.contains is inclusive on both ends; NaN is outside.
  * dr.compress returns ascending indices.
  * masked dr.gather returns 0 on inactive lanes.
  * scatter_reduce(Add) on fp32 = fp32 atomics in some order; here np.add.at
    (index order).  `exact=True` accumulates in float64 and rounds once, which is
    what tolerance checks of the CUDA splat are held against.
  * sampler.next_2d = two consecutive next_1d; the sampler is replaced by an
    explicit uniform stream u[lane, 3*level + {0,1,2}] = (u_x, u_y, u_select),
    consumed exactly as src/quadtree.py:956,980 consume the PCG32 stream
    (3 floats per visited node, the leaf included).
  * python-float thresholds are converted to fp32 before comparing with fp32 arrays.
"""
import math
import numpy as np

from . import drjit_math as dm

F = np.float32
U = np.uint32

# Open uncertainty: a Dr.Jit build may lower `irradiance / 100` (src/quadtree.py:519)
# to a multiplication by the fp32 reciprocal.  False = IEEE fp32 division (the default everywhere);
# tests flip it together with the library's "quad_thr_reciprocal" switch.
QUAD_THR_RECIPROCAL = False


# --------------------------------------------------------------------------- helpers
def gather(src, idx, active=None):
    """dr.gather with optional mask: inactive lanes read 0."""
    idx = np.asarray(idx)
    if active is None:
        return src[idx]
    out = np.zeros((idx.shape[0],) + src.shape[1:], dtype=src.dtype)
    if active.any():
        out[active] = src[idx[active]]
    return out


def compress(mask):
    """dr.compress: ascending indices of set lanes."""
    return np.flatnonzero(mask).astype(U)


def resize(arr, new_size, default_zero=True):
    """resizeDrJitArray, src/common.py:161-189."""
    old = arr.shape[0]
    if old <= new_size:
        fill = np.zeros if default_zero else np.ones
        tail = fill((new_size - old,) + arr.shape[1:], dtype=arr.dtype)
        return np.concatenate([arr, tail], axis=0)
    return arr[:new_size].copy()


def concat(a, b):
    """concatDrJitArray, src/common.py:192-225."""
    return np.concatenate([a, b], axis=0)


def bbox_contains(bmin, bmax, p):
    """mi.BoundingBox{2,3}f.contains: inclusive, all axes."""
    with np.errstate(invalid='ignore'):
        return np.all((p >= bmin) & (p <= bmax), axis=-1)


LCG_M, LCG_C = 747796405, 2891336453


def counter_uniform(seed, lane, idx):
    """Perf-mode uniform in [0,1) keyed (seed, lane, idx), stated identically in the CUDA library
    (the reference's PCG32 state is not reachable from Python).  h0 = murmur3 finaliser of
    (seed, lane).  idx = 3*level + k as QuadTree.sampleQuadTree consumes them: k == 2 (child selection)
    is the top 24 bits of the 32-bit LCG stream t_{level+1} = t_level*M + C started at t_0 = h0;
    k in (0, 1) (leaf position) is a second finaliser round of h0 ^ f(idx)."""
    def fmix(h):
        h = h ^ (h >> U(16))
        h = (h * U(0x85EBCA6B)).astype(U)
        h = h ^ (h >> U(13))
        h = (h * U(0xC2B2AE35)).astype(U)
        h = h ^ (h >> U(16))
        return h
    with np.errstate(over='ignore'):
        lane = np.asarray(lane, dtype=U)
        idx = np.asarray(idx, dtype=U)
        h0 = fmix((U(seed) + lane * U(0x9E3779B1)).astype(U))
        h = fmix((h0 ^ (idx * U(0x85EBCA77) + U(0x165667B1)).astype(U)).astype(U))
        sel = np.broadcast_to((idx % U(3)) == U(2), np.broadcast(h0, idx).shape)
        if sel.any():
            steps = np.broadcast_to(idx // U(3) + U(1), sel.shape)
            t = np.broadcast_to(h0, sel.shape).copy()
            for k in range(int(steps[sel].max())):
                go = sel & (steps > k)
                t[go] = (t[go].astype(np.uint64) * np.uint64(LCG_M) + np.uint64(LCG_C)).astype(U)
            h = np.where(sel, t, h)
    return ((h >> U(8)).astype(F) * F(2.0 ** -24)).astype(F)


class ExplicitSampler:
    """Stand-in for mi.Sampler inside QuadTree.sampleQuadTree: serves
    u[lane, cursor] and advances the cursor of the lanes that execute the call."""

    def __init__(self, u=None, seed=None, n=None, lane_offset=0):
        self.u = None if u is None else np.asarray(u, dtype=F)
        self.seed = seed
        self.n = self.u.shape[0] if self.u is not None else n
        self.lane = (np.arange(self.n, dtype=np.uint64) + lane_offset).astype(U)
        self.cursor = np.zeros(self.n, dtype=np.int64)

    def _take(self, lanes_mask):
        if self.u is not None:
            c = np.minimum(self.cursor, self.u.shape[1] - 1)
            v = self.u[np.arange(self.n), c]
        else:
            v = counter_uniform(self.seed, self.lane, self.cursor.astype(U))
        self.cursor = self.cursor + lanes_mask.astype(np.int64)
        return v.astype(F)

    def next_1d(self, executing):
        return self._take(executing)

    def next_2d(self, executing):
        x = self._take(executing)
        y = self._take(executing)
        return np.stack([x, y], axis=1)


class SurfaceInteractionRecord:
    """Fields of src/common.py:14-40 that the tree consumes."""

    def __init__(self, position, direction, radiance, woPdf,
                 radiance_nee=None, direction_nee=None):
        n = position.shape[0]
        self.position = np.asarray(position, dtype=F).reshape(n, 3)
        self.direction = np.asarray(direction, dtype=F).reshape(n, 2)
        self.radiance = np.asarray(radiance, dtype=F).reshape(n)
        self.woPdf = np.asarray(woPdf, dtype=F).reshape(n)
        self.radiance_nee = (np.zeros((n, 3), F) if radiance_nee is None
                             else np.asarray(radiance_nee, dtype=F).reshape(n, 3))
        self.direction_nee = (np.zeros((n, 2), F) if direction_nee is None
                              else np.asarray(direction_nee, dtype=F).reshape(n, 2))


# --------------------------------------------------------------------------- quadtree
class QuadTreeNode:
    """SoA node store shared by all trees, src/quadtree.py:12-345."""

    FIELDS = ('bbox_min', 'bbox_max', 'depth', 'irradiance', 'isLeaf',
              'refinementThreshold', 'child_1_index', 'child_2_index',
              'child_3_index', 'child_4_index')

    def __init__(self, size=0):
        # dr.zeros(QuadTreeNode, shape=size)
        self.rootNodeIndex = np.zeros(0, U)
        self.bbox_min = np.zeros((size, 2), F)
        self.bbox_max = np.zeros((size, 2), F)
        self.depth = np.zeros(size, U)
        self.irradiance = np.zeros(size, F)
        self.isLeaf = np.zeros(size, bool)
        self.refinementThreshold = np.zeros(size, F)
        self.child_1_index = np.zeros(size, U)
        self.child_2_index = np.zeros(size, U)
        self.child_3_index = np.zeros(size, U)
        self.child_4_index = np.zeros(size, U)
        # float64 shadow of `irradiance` used when exact=True splats are requested
        self.irradiance64 = None

    def copyFrom(self, o):                                   # :40-55
        self.rootNodeIndex = o.rootNodeIndex.copy()
        for f in self.FIELDS:
            setattr(self, f, getattr(o, f).copy())

    def getWidth(self):                                      # :74-75
        return self.depth.shape[0]

    def children(self):
        return (self.child_1_index, self.child_2_index, self.child_3_index, self.child_4_index)

    def addIrradiance(self, idx, irradiance, active):        # :88-93
        if active.any():
            if self.irradiance64 is not None:
                np.add.at(self.irradiance64, idx[active], irradiance[active].astype(np.float64))
            else:
                np.add.at(self.irradiance, idx[active], irradiance[active])

    def resize(self, new_size):                              # :216-247
        self.depth = resize(self.depth, new_size)
        self.irradiance = resize(self.irradiance, new_size)
        self.isLeaf = resize(self.isLeaf, new_size, default_zero=False)
        self.refinementThreshold = resize(self.refinementThreshold, new_size)
        self.child_1_index = resize(self.child_1_index, new_size)
        self.child_2_index = resize(self.child_2_index, new_size)
        self.child_3_index = resize(self.child_3_index, new_size)
        self.child_4_index = resize(self.child_4_index, new_size)
        self.bbox_min = resize(self.bbox_min, new_size)
        self.bbox_max = resize(self.bbox_max, new_size)

    def split(self, idx):                                    # :96-191
        idx = np.asarray(idx, dtype=U)
        num = idx.shape[0]
        old_size = self.getWidth()
        self.resize(old_size + num * 4)
        r = np.arange(num, dtype=U)
        c1 = r * U(4) + U(0) + U(old_size)
        c2 = r * U(4) + U(1) + U(old_size)
        c3 = r * U(4) + U(2) + U(old_size)
        c4 = r * U(4) + U(3) + U(old_size)
        self.child_1_index[idx] = c1
        self.child_2_index[idx] = c2
        self.child_3_index[idx] = c3
        self.child_4_index[idx] = c4
        self.isLeaf[idx] = False
        depth = self.depth[idx] + U(1)
        for c in (c1, c2, c3, c4):
            self.depth[c] = depth
        with np.errstate(all='ignore'):
            irr = (self.irradiance[idx] / F(4)).astype(F)        # :133-134
        for c in (c1, c2, c3, c4):
            self.irradiance[c] = irr
        thr = self.refinementThreshold[idx]
        for c in (c1, c2, c3, c4):
            self.refinementThreshold[c] = thr
        bmin = self.bbox_min[idx]
        bmax = self.bbox_max[idx]
        bmid = ((bmin + bmax) / F(2)).astype(F)                  # :151
        # quadrant 1: [mid, max]
        self.bbox_min[c1] = bmid
        self.bbox_max[c1] = bmax
        # quadrant 2: x in [min.x, mid.x], y in [mid.y, max.y]
        q2min = bmin.copy(); q2min[:, 1] = bmid[:, 1]
        q2max = bmax.copy(); q2max[:, 0] = bmid[:, 0]
        self.bbox_min[c2] = q2min
        self.bbox_max[c2] = q2max
        # quadrant 3: [min, mid]
        self.bbox_min[c3] = bmin
        self.bbox_max[c3] = bmid
        # quadrant 4: x in [mid.x, max.x], y in [min.y, mid.y]
        q4min = bmin.copy(); q4min[:, 0] = bmid[:, 0]
        q4max = bmax.copy(); q4max[:, 1] = bmid[:, 1]
        self.bbox_min[c4] = q4min
        self.bbox_max[c4] = q4max

    def merge(self, idx):                                    # :194-213
        if idx.shape[0] == 0:
            return
        self.child_1_index[idx] = 0
        self.child_2_index[idx] = 0
        self.child_3_index[idx] = 0
        self.child_4_index[idx] = 0
        self.isLeaf[idx] = True

    def createRootNode(self, num):                           # :250-285
        old_root = self.rootNodeIndex.shape[0]
        self.rootNodeIndex = resize(self.rootNodeIndex, old_root + num)
        old_size = self.getWidth()
        self.resize(old_size + num)
        new_node = np.arange(num, dtype=U) + U(old_size)
        new_root = np.arange(num, dtype=U) + U(old_root)
        self.rootNodeIndex[new_root] = new_node
        self.bbox_min[new_node] = 0
        self.bbox_max[new_node] = 1
        self.isLeaf[new_node] = True
        return new_root

    def getAllLeafNodeIndex(self, rootIndex=None):           # :288-345
        if rootIndex is None or rootIndex.shape[0] == 0:
            return compress(self.isLeaf)
        nodeIndex = self.rootNodeIndex[rootIndex]
        allLeaf = np.zeros(0, U)
        active = True
        while active:
            isLeaf = self.isLeaf[nodeIndex]
            leafNode = nodeIndex[compress(isLeaf)]
            if leafNode.shape[0] > 0:
                allLeaf = concat(allLeaf, leafNode)
            nonLeaf = nodeIndex[compress(~isLeaf)]
            nodeIndex = concat(concat(self.child_1_index[nonLeaf], self.child_2_index[nonLeaf]),
                               concat(self.child_3_index[nonLeaf], self.child_4_index[nonLeaf]))
            active = nodeIndex.shape[0] > 0
        return allLeaf


class QuadTree:
    """src/quadtree.py:348-1101."""

    def __init__(self, maxDepth=20, isStoreNEERadiance=False):   # :350-362
        q = QuadTreeNode(1)
        q.rootNodeIndex = np.zeros(1, U)
        q.refinementThreshold[:] = F(np.inf)
        q.isLeaf[:] = True
        q.bbox_min[:] = 0
        q.bbox_max[:] = 1
        self.quadTreeNode = q
        self.maxDepth = maxDepth
        self.isStoreNEERadiance = isStoreNEERadiance

    def createRootNode(self, num):
        return self.quadTreeNode.createRootNode(num)

    # ---- splat ---------------------------------------------------------- :389-464
    def addDataPropagate(self, rootIndex, rec):
        q = self.quadTreeNode

        def propagate(position, irradiance):
            nodeIndex = q.rootNodeIndex[rootIndex].copy()
            active = bbox_contains(q.bbox_min[nodeIndex], q.bbox_max[nodeIndex], position)
            while active.any():
                q.addIrradiance(nodeIndex, irradiance, active)
                isLeaf = gather(q.isLeaf, nodeIndex, active)
                active = active & ~isLeaf
                cis = [gather(child, nodeIndex, active) for child in q.children()]   # :419-422
                for cidx in cis:                                                      # :424-438
                    test = bbox_contains(q.bbox_min[cidx], q.bbox_max[cidx], position)
                    m = test & active
                    nodeIndex[m] = cidx[m]

        with np.errstate(all='ignore'):
            irr = np.where(rec.woPdf > 0, rec.radiance / rec.woPdf, F(0)).astype(F)      # :451
        propagate(rec.direction, irr)
        if self.isStoreNEERadiance:                                                      # :455-464
            with np.errstate(all='ignore'):
                lum = dm.luminance(rec.radiance_nee)
                irr_nee = np.where(rec.woPdf > 0, lum / rec.woPdf, F(0)).astype(F)
            propagate(rec.direction_nee, irr_nee)

    # ---- validators ------------------------------------------------------ :468-509
    def validateQuadTreeNodeBBox(self, q=None):
        q = self.quadTreeNode if q is None else q
        idx = np.arange(q.getWidth(), dtype=U)
        active = ~q.isLeaf
        ok = np.ones(q.getWidth(), bool)
        for child in q.children():
            c = gather(child, idx, active)
            ok &= np.all((q.bbox_min[c] >= q.bbox_min[idx]) & (q.bbox_max[c] <= q.bbox_max[idx]), axis=1)
        return not bool((active & ~ok).any())

    # ---- refine ---------------------------------------------------------- :512-637
    def setRefinementThreshold(self, rootIndex, total_flux_prev_quadtree):
        q = self.quadTreeNode
        with np.errstate(all='ignore'):
            e = np.asarray(total_flux_prev_quadtree, dtype=F)
            thr = (e * F(1.0 / 100) if QUAD_THR_RECIPROCAL else e / F(100)).astype(F)      # :519
        nodeIndex = q.rootNodeIndex[rootIndex]
        active = nodeIndex.shape[0] > 0
        while active:
            q.refinementThreshold[nodeIndex] = thr
            notLeaf = ~q.isLeaf[nodeIndex]
            active = bool(notLeaf.any())
            if active:
                sel = compress(notLeaf)
                nl = nodeIndex[sel]
                nthr = thr[sel]
                nodeIndex = concat(concat(q.child_1_index[nl], q.child_2_index[nl]),
                                   concat(q.child_3_index[nl], q.child_4_index[nl]))
                thr = concat(concat(nthr, nthr), concat(nthr, nthr))

    def refine(self, rootIndex):
        q = self.quadTreeNode
        # merge pass :574-611
        parent = q.rootNodeIndex[rootIndex]
        active = parent.shape[0] > 0
        while active:
            notLeaf = ~q.isLeaf[parent]
            irr = q.irradiance[parent]
            thr = q.refinementThreshold[parent]
            with np.errstate(invalid='ignore'):
                small = notLeaf & (irr < thr)
                normal = notLeaf & (irr >= thr)
            q.merge(parent[compress(small)])
            valid = parent[compress(normal)]
            parent = concat(concat(q.child_1_index[valid], q.child_2_index[valid]),
                            concat(q.child_3_index[valid], q.child_4_index[valid]))
            active = parent.shape[0] > 0
        # split pass :617-637
        active = True
        while active:
            leaf = q.getAllLeafNodeIndex(rootIndex)
            irr = q.irradiance[leaf]
            thr = q.refinementThreshold[leaf]
            depth = q.depth[leaf]
            with np.errstate(invalid='ignore'):
                cond = (irr > thr) & (depth < self.maxDepth)
            active = bool(cond.any())
            if active:
                q.split(leaf[compress(cond)])

    # ---- reset ----------------------------------------------------------- :640-683
    def resetTreeIrradiance(self, rootIndex):
        q = self.quadTreeNode
        nodeIndex = q.rootNodeIndex[rootIndex]
        active = nodeIndex.shape[0] > 0
        while active:
            q.irradiance[nodeIndex] = 0
            notLeaf = ~q.isLeaf[nodeIndex]
            active = bool(notLeaf.any())
            if active:
                nl = nodeIndex[compress(notLeaf)]
                nodeIndex = concat(concat(q.child_1_index[nl], q.child_2_index[nl]),
                                   concat(q.child_3_index[nl], q.child_4_index[nl]))

    def resetAllTreeIrradiance(self):
        self.resetTreeIrradiance(self.quadTreeNode.rootNodeIndex)

    def getAllLeafNodeIndex(self, rootIndex=None):
        return self.quadTreeNode.getAllLeafNodeIndex(rootIndex)

    # ---- copy / compact / append ------------------------------------------ :695-928
    def copyTree(self, rootIndex):
        src = self.quadTreeNode
        rootIndex = np.atleast_1d(np.asarray(rootIndex, dtype=U))
        out = QuadTreeNode(0)
        out.rootNodeIndex = np.arange(rootIndex.shape[0], dtype=U)
        nodeIndex = src.rootNodeIndex[rootIndex]
        parentIndex = np.zeros(0, U)
        parentChildIndex = np.zeros(0, U)
        active = nodeIndex.shape[0] > 0
        while active:
            num = nodeIndex.shape[0]
            old_size = out.getWidth()
            out.resize(old_size + num)
            newNodeIndex = np.arange(num, dtype=U) + U(old_size)
            if parentIndex.shape[0] > 0:
                for k, child in enumerate(out.children(), start=1):
                    m = parentChildIndex == k
                    child[parentIndex[m]] = newNodeIndex[m]
            out.bbox_min[newNodeIndex] = src.bbox_min[nodeIndex]
            out.bbox_max[newNodeIndex] = src.bbox_max[nodeIndex]
            out.depth[newNodeIndex] = src.depth[nodeIndex]
            out.irradiance[newNodeIndex] = src.irradiance[nodeIndex]
            isLeaf = src.isLeaf[nodeIndex]
            out.isLeaf[newNodeIndex] = isLeaf
            out.refinementThreshold[newNodeIndex] = src.refinementThreshold[nodeIndex]
            sel = compress(~isLeaf)
            nonLeaf = nodeIndex[sel]
            newNonLeaf = newNodeIndex[sel]
            if nonLeaf.shape[0] > 0:
                k = nonLeaf.shape[0]
                parentIndex = np.repeat(newNonLeaf, 4)
                parentChildIndex = np.tile(np.arange(1, 5, dtype=U), k)
                nodeIndex = np.zeros(k * 4, U)
                nodeIndex[0::4] = src.child_1_index[nonLeaf]
                nodeIndex[1::4] = src.child_2_index[nonLeaf]
                nodeIndex[2::4] = src.child_3_index[nonLeaf]
                nodeIndex[3::4] = src.child_4_index[nonLeaf]
                active = True
            else:
                active = False
        return out

    def copyFrom(self, o):                                   # :831-841
        self.quadTreeNode.copyFrom(o.quadTreeNode)
        self.maxDepth = o.maxDepth
        self.isStoreNEERadiance = o.isStoreNEERadiance

    def clearTreeUnusedNode(self):                           # :844-851
        n = self.quadTreeNode.rootNodeIndex.shape[0]
        self.quadTreeNode = self.copyTree(np.arange(n, dtype=U))

    def appendQuadTreeNode(self, other):                     # :854-928
        q = self.quadTreeNode
        old_root = q.rootNodeIndex.shape[0]
        in_root = other.rootNodeIndex.shape[0]
        q.rootNodeIndex = resize(q.rootNodeIndex, old_root + in_root)
        old_size = q.getWidth()
        in_size = other.depth.shape[0]
        q.resize(old_size + in_size)
        off = U(old_size)
        other.rootNodeIndex = other.rootNodeIndex + off
        notLeaf = ~other.isLeaf
        for child in other.children():
            child[notLeaf] += off
        root_slots = np.arange(in_root, dtype=U) + U(old_root)
        q.rootNodeIndex[root_slots] = other.rootNodeIndex
        sl = slice(old_size, old_size + in_size)
        for f in QuadTreeNode.FIELDS:
            getattr(q, f)[sl] = getattr(other, f)
        return root_slots

    # ---- sample ----------------------------------------------------------- :931-998
    def sampleQuadTree(self, rootIndex, sampler, active_sample=True, return_node=False):
        q = self.quadTreeNode
        n = rootIndex.shape[0]
        nodeIndex = q.rootNodeIndex[rootIndex].copy()
        pos = np.zeros((n, 2), F)
        active = np.broadcast_to(np.asarray(active_sample, bool), (n,)).copy()
        guard = 0
        while active.any():
            executing = active.copy()          # lanes that run this loop body
            isLeaf = gather(q.isLeaf, nodeIndex, active)
            bmin = gather(q.bbox_min, nodeIndex, active)
            bmax = gather(q.bbox_max, nodeIndex, active)
            u2 = sampler.next_2d(executing)                                        # :956
            m = active & isLeaf
            with np.errstate(all='ignore'):
                cand = (bmin + u2 * (bmax - bmin)).astype(F)
            pos[m] = cand[m]
            active = active & ~isLeaf
            ci = [gather(c, nodeIndex, active) for c in q.children()]
            ce = [gather(q.irradiance, c, active) for c in ci]
            with np.errstate(all='ignore'):
                e1 = ce[0]
                e2 = (ce[1] + e1).astype(F)
                e3 = (ce[2] + e2).astype(F)
                e4 = (ce[3] + e3).astype(F)
                s = (sampler.next_1d(executing) * e4).astype(F)                    # :980 (unmasked)
                pick = [s < e1, (e1 <= s) & (s < e2), (e2 <= s) & (s < e3), e3 <= s]
            moved = np.zeros(n, bool)
            for k in range(4):
                m = active & pick[k]
                nodeIndex[m] = ci[k][m]
                moved |= m
            # The reference would spin forever on a lane whose child energies are
            # NaN (no bin matches).  Defined behaviour here and in the CUDA library:
            # such a lane stops with position (0,0).
            active = active & moved
            guard += 1
            assert guard < 4096
        d = dm.canonical_to_dir(pos)                                               # :996
        if return_node:
            return d, nodeIndex, pos
        return d

    # ---- pdf -------------------------------------------------------------- :1001-1101
    def pdfQuadTree(self, rootIndex, direction, active=True, return_node=False):
        q = self.quadTreeNode
        n = rootIndex.shape[0]
        nodeIndex = q.rootNodeIndex[rootIndex].copy()
        pdf = np.ones(n, F)
        act = np.broadcast_to(np.asarray(active, bool), (n,)).copy()
        position = dm.dir_to_canonical(direction)                                  # :1016
        guard = 0
        while act.any():
            isLeaf = gather(q.isLeaf, nodeIndex, act)
            m = act & isLeaf
            pdf[m] = (pdf[m] * dm.INV_FOUR_PI).astype(F)                           # :1030
            act = act & ~isLeaf
            ci = [gather(c, nodeIndex, act) for c in q.children()]
            test = [bbox_contains(q.bbox_min[c], q.bbox_max[c], position) for c in ci]
            nodeE = gather(q.irradiance, nodeIndex, act)
            ce = [gather(q.irradiance, ci[k], act & test[k]) for k in range(4)]
            childE = np.where(test[0], ce[0],
                              np.where(test[1], ce[1],
                                       np.where(test[2], ce[2],
                                                np.where(test[3], ce[3], F(0))))).astype(F)
            with np.errstate(all='ignore'):
                ratio = ((F(4) * childE) / nodeE).astype(F)                        # :1084
                newpdf = (pdf * ratio).astype(F)
            pdf[act] = newpdf[act]
            isnan = np.isnan(pdf)
            m = act & isnan
            pdf[m] = 0                                                             # :1090-1092
            act = act & ~isnan
            moved = np.zeros(n, bool)
            for k in range(4):
                m = act & test[k]
                nodeIndex[m] = ci[k][m]
                moved |= m
            act = act & moved      # (cannot trigger for finite positions; loop guard)
            guard += 1
            assert guard < 4096
        if return_node:
            return pdf, nodeIndex, position
        return pdf


# --------------------------------------------------------------------------- kd-tree
class KDTreeNode:
    """src/kdtree.py:16-105."""

    FIELDS = ('bbox_min', 'bbox_max', 'depth', 'vertCount', 'isLeaf',
              'quadTreeRootIndex', 'child_left_index', 'child_right_index')

    def __init__(self, size=0):
        self.bbox_min = np.zeros((size, 3), F)
        self.bbox_max = np.zeros((size, 3), F)
        self.depth = np.zeros(size, U)
        self.vertCount = np.zeros(size, F)
        self.isLeaf = np.zeros(size, bool)
        self.quadTreeRootIndex = np.zeros(size, U)
        self.child_left_index = np.zeros(size, U)
        self.child_right_index = np.zeros(size, U)

    def copyFrom(self, o):                                   # :38-50
        for f in self.FIELDS:
            setattr(self, f, getattr(o, f).copy())

    def getWidth(self):
        return self.depth.shape[0]

    def resize(self, new_size):                              # :79-105
        self.depth = resize(self.depth, new_size)
        self.vertCount = resize(self.vertCount, new_size)
        self.isLeaf = resize(self.isLeaf, new_size, default_zero=False)
        self.quadTreeRootIndex = resize(self.quadTreeRootIndex, new_size)
        self.child_left_index = resize(self.child_left_index, new_size)
        self.child_right_index = resize(self.child_right_index, new_size)
        self.bbox_min = resize(self.bbox_min, new_size)
        self.bbox_max = resize(self.bbox_max, new_size)


class KDTree:
    """src/kdtree.py:108-663."""

    NPZ_KEYS = ('kdtree_maxLeafSize', 'kdtree_maxDepth', 'kdtree_bbox_min', 'kdtree_bbox_max',
                'kdtree_depth', 'kdtree_vertCount', 'kdtree_isLeaf', 'kdtree_quadTreeRootIndex',
                'kdtree_child_left_index', 'kdtree_child_right_index',
                'quadtree_maxDepth', 'quadtree_isStoreNEERadiance', 'quadtree_rootNodeIndex',
                'quadtree_bbox_min', 'quadtree_bbox_max', 'quadtree_depth', 'quadtree_irradiance',
                'quadtree_isLeaf', 'quadtree_refinementThreshold', 'quadtree_child_1_index',
                'quadtree_child_2_index', 'quadtree_child_3_index', 'quadtree_child_4_index')

    def __init__(self, max_leaf_size=1, maxDepth=10):        # :117-130
        k = KDTreeNode(1)
        k.isLeaf[:] = True
        k.bbox_min[:] = 0
        k.bbox_max[:] = 1
        self.kdTreeNode = k
        self.maxLeafSize = max_leaf_size
        self.maxDepth = maxDepth
        self.quadTree = QuadTree()

    def setup(self, bbox_min, bbox_max):                     # :133-138
        self.kdTreeNode.bbox_min = np.asarray(bbox_min, F).reshape(1, 3).copy()
        self.kdTreeNode.bbox_max = np.asarray(bbox_max, F).reshape(1, 3).copy()

    def copyFrom(self, o):                                   # :141-153
        self.kdTreeNode.copyFrom(o.kdTreeNode)
        self.maxLeafSize = o.maxLeafSize
        self.maxDepth = o.maxDepth
        self.quadTree.copyFrom(o.quadTree)

    def getAllLeafNodeIndex(self):                           # :173-177
        return compress(self.kdTreeNode.isLeaf)

    # ---- splat ------------------------------------------------------------ :180-225
    def addDataPropagate(self, rec, exact=False):
        k = self.kdTreeNode
        q = self.quadTree.quadTreeNode
        if exact and q.irradiance64 is None:
            q.irradiance64 = q.irradiance.astype(np.float64)
        p = rec.position
        n = p.shape[0]
        nodeIndex = np.zeros(n, U)
        active = bbox_contains(k.bbox_min[0], k.bbox_max[0], p)
        while active.any():
            np.add.at(k.vertCount, nodeIndex[active], F(1))                 # :199
            isLeaf = gather(k.isLeaf, nodeIndex, active)
            active = active & ~isLeaf
            cis = [gather(child, nodeIndex, active) for child in (k.child_left_index, k.child_right_index)]
            for cidx in cis:                                                # left first, right second
                test = bbox_contains(k.bbox_min[cidx], k.bbox_max[cidx], p)
                m = test & active
                nodeIndex[m] = cidx[m]
        quadTreeRoot = k.quadTreeRootIndex[nodeIndex]                       # :224 (unmasked)
        self.quadTree.addDataPropagate(quadTreeRoot, rec)
        if exact:
            q.irradiance = q.irradiance64.astype(F)

    # ---- split ------------------------------------------------------------ :229-323
    def split(self, idx):
        k = self.kdTreeNode
        idx = np.asarray(idx, dtype=U)
        num = idx.shape[0]
        old_size = k.getWidth()
        k.resize(old_size + num * 2)
        r = np.arange(num, dtype=U)
        left = r * U(2) + U(0) + U(old_size)
        right = r * U(2) + U(1) + U(old_size)
        k.child_left_index[idx] = left
        k.child_right_index[idx] = right
        k.isLeaf[idx] = False
        depth = k.depth[idx]
        k.depth[left] = depth + U(1)
        k.depth[right] = depth + U(1)
        vc = k.vertCount[idx].copy()
        pos = vc > 0
        vc[pos] = (vc[pos] / F(2)).astype(F)                                # :262
        k.vertCount[left] = vc
        k.vertCount[right] = vc
        bmin = k.bbox_min[idx]
        bmax = k.bbox_max[idx]
        bmid = ((bmin + bmax) / F(2)).astype(F)                             # :270
        axis = (depth % U(3)).astype(np.int64)                              # :277
        rows = np.arange(num)
        mid_c = bmid[rows, axis]
        lmax = bmax.copy(); lmax[rows, axis] = mid_c
        rmin = bmin.copy(); rmin[rows, axis] = mid_c
        k.bbox_min[left] = bmin
        k.bbox_max[left] = lmax
        k.bbox_min[right] = rmin
        k.bbox_max[right] = bmax
        qroot = k.quadTreeRootIndex[idx]                                    # :316-317
        k.quadTreeRootIndex[left] = qroot
        copy = self.quadTree.copyTree(qroot)                                # :320
        new_roots = self.quadTree.appendQuadTreeNode(copy)                  # :322
        k.quadTreeRootIndex[right] = new_roots

    def setRefinementThreshold(self, iteration):             # :327-330
        c = 12000
        self.maxLeafSize = c * math.sqrt(math.pow(2, iteration))

    def refine(self):                                        # :333-358
        k = self.kdTreeNode
        active = True
        while active:
            leaf = self.getAllLeafNodeIndex()
            vc = k.vertCount[leaf]
            depth = k.depth[leaf]
            cond = (vc > F(self.maxLeafSize)) & (depth < self.maxDepth)
            active = bool(cond.any())
            if active:
                self.split(leaf[compress(cond)])

    def validateTreeNodeBBox(self):                          # :361-398
        k = self.kdTreeNode
        idx = np.arange(k.getWidth(), dtype=U)
        active = ~k.isLeaf
        ok = np.ones(k.getWidth(), bool)
        for child in (k.child_left_index, k.child_right_index):
            c = gather(child, idx, active)
            ok &= np.all((k.bbox_min[c] >= k.bbox_min[idx]) & (k.bbox_max[c] <= k.bbox_max[idx]), axis=1)
        return not bool((active & ~ok).any())

    def resetTreeVertCount(self):                            # :401-432
        k = self.kdTreeNode
        nodeIndex = np.zeros(1, U)
        active = True
        while active:
            k.vertCount[nodeIndex] = 0
            notLeaf = ~k.isLeaf[nodeIndex]
            active = bool(notLeaf.any())
            if active:
                nl = nodeIndex[compress(notLeaf)]
                nodeIndex = concat(k.child_left_index[nl], k.child_right_index[nl])

    # ---- queries ---------------------------------------------------------- :435-496
    def getLeafNodeIndex(self, position, active=True):
        k = self.kdTreeNode
        position = np.asarray(position, F)
        n = position.shape[0]
        nodeIndex = np.zeros(n, U)
        act = bbox_contains(k.bbox_min[0], k.bbox_max[0], position) & np.broadcast_to(np.asarray(active, bool), (n,))
        while act.any():
            isLeaf = gather(k.isLeaf, nodeIndex, act)
            act = act & ~isLeaf
            cis = [gather(child, nodeIndex, act) for child in (k.child_left_index, k.child_right_index)]
            for cidx in cis:                                                # :462-468 left first, right second
                test = bbox_contains(k.bbox_min[cidx], k.bbox_max[cidx], position)
                m = test & act
                nodeIndex[m] = cidx[m]
        return nodeIndex

    def sample(self, position, sampler, active=True, return_debug=False):
        n = position.shape[0]
        act = np.broadcast_to(np.asarray(active, bool), (n,))
        leaf = self.getLeafNodeIndex(position, act)
        root = gather(self.kdTreeNode.quadTreeRootIndex, leaf, act)
        d, snode, spos = self.quadTree.sampleQuadTree(root, sampler, act, return_node=True)
        pdf, pnode, ppos = self.quadTree.pdfQuadTree(root, d, act, return_node=True)
        if return_debug:
            return d, pdf, dict(leaf=leaf, root=root, sample_node=snode, sample_pos=spos,
                                pdf_node=pnode, pdf_pos=ppos)
        return d, pdf

    def pdf(self, position, direction, active=True, return_debug=False):
        n = position.shape[0]
        act = np.broadcast_to(np.asarray(active, bool), (n,))
        leaf = self.getLeafNodeIndex(position, act)
        root = gather(self.kdTreeNode.quadTreeRootIndex, leaf, act)
        pdf, pnode, ppos = self.quadTree.pdfQuadTree(root, direction, act, return_node=True)
        if return_debug:
            return pdf, dict(leaf=leaf, root=root, pdf_node=pnode, pdf_pos=ppos)
        return pdf

    # ---- quadtree handlers -------------------------------------------------- :503-532
    def setQuadTreeRefinementThreshold(self):
        leaf = self.getAllLeafNodeIndex()
        roots = self.kdTreeNode.quadTreeRootIndex[leaf]
        q = self.quadTree.quadTreeNode
        rootNode = q.rootNodeIndex[roots]
        self.quadTree.setRefinementThreshold(roots, q.irradiance[rootNode])

    def refineAllQuadTree(self):
        leaf = self.getAllLeafNodeIndex()
        roots = self.kdTreeNode.quadTreeRootIndex[leaf]
        self.quadTree.refine(roots)

    def cleanUnusedQuadTree(self):
        self.quadTree.clearTreeUnusedNode()

    def resetAllQuadTreeIrradiance(self):
        self.quadTree.resetAllTreeIrradiance()
        self.quadTree.quadTreeNode.irradiance64 = None

    # ---- npz I/O ------------------------------------------------------------ :539-602, 156-170
    def to_arrays(self):
        k = self.kdTreeNode
        q = self.quadTree.quadTreeNode
        return dict(
            kdtree_maxLeafSize=np.asarray(self.maxLeafSize),
            kdtree_maxDepth=np.asarray(self.maxDepth),
            kdtree_bbox_min=k.bbox_min, kdtree_bbox_max=k.bbox_max,
            kdtree_depth=k.depth, kdtree_vertCount=k.vertCount, kdtree_isLeaf=k.isLeaf,
            kdtree_quadTreeRootIndex=k.quadTreeRootIndex,
            kdtree_child_left_index=k.child_left_index, kdtree_child_right_index=k.child_right_index,
            quadtree_maxDepth=np.asarray(self.quadTree.maxDepth),
            quadtree_isStoreNEERadiance=np.asarray(self.quadTree.isStoreNEERadiance),
            quadtree_rootNodeIndex=q.rootNodeIndex,
            quadtree_bbox_min=q.bbox_min, quadtree_bbox_max=q.bbox_max,
            quadtree_depth=q.depth, quadtree_irradiance=q.irradiance, quadtree_isLeaf=q.isLeaf,
            quadtree_refinementThreshold=q.refinementThreshold,
            quadtree_child_1_index=q.child_1_index, quadtree_child_2_index=q.child_2_index,
            quadtree_child_3_index=q.child_3_index, quadtree_child_4_index=q.child_4_index)

    def saveToFile(self, fileName):
        np.savez_compressed(fileName, **self.to_arrays())

    def loadFromArrays(self, d):
        # NB the reference truncates maxLeafSize with int() on load (:161)
        self.maxLeafSize = int(d['kdtree_maxLeafSize'])
        self.maxDepth = int(d['kdtree_maxDepth'])
        k = KDTreeNode(0)
        k.bbox_min = np.asarray(d['kdtree_bbox_min'], F).reshape(-1, 3).copy()
        k.bbox_max = np.asarray(d['kdtree_bbox_max'], F).reshape(-1, 3).copy()
        k.depth = np.asarray(d['kdtree_depth'], U).copy()
        k.vertCount = np.asarray(d['kdtree_vertCount'], F).copy()
        k.isLeaf = np.asarray(d['kdtree_isLeaf'], bool).copy()
        k.quadTreeRootIndex = np.asarray(d['kdtree_quadTreeRootIndex'], U).copy()
        k.child_left_index = np.asarray(d['kdtree_child_left_index'], U).copy()
        k.child_right_index = np.asarray(d['kdtree_child_right_index'], U).copy()
        self.kdTreeNode = k
        self.quadTree = QuadTree(int(d['quadtree_maxDepth']), bool(d['quadtree_isStoreNEERadiance']))
        q = QuadTreeNode(0)
        q.rootNodeIndex = np.asarray(d['quadtree_rootNodeIndex'], U).copy()
        q.bbox_min = np.asarray(d['quadtree_bbox_min'], F).reshape(-1, 2).copy()
        q.bbox_max = np.asarray(d['quadtree_bbox_max'], F).reshape(-1, 2).copy()
        q.depth = np.asarray(d['quadtree_depth'], U).copy()
        q.irradiance = np.asarray(d['quadtree_irradiance'], F).copy()
        q.isLeaf = np.asarray(d['quadtree_isLeaf'], bool).copy()
        q.refinementThreshold = np.asarray(d['quadtree_refinementThreshold'], F).copy()
        q.child_1_index = np.asarray(d['quadtree_child_1_index'], U).copy()
        q.child_2_index = np.asarray(d['quadtree_child_2_index'], U).copy()
        q.child_3_index = np.asarray(d['quadtree_child_3_index'], U).copy()
        q.child_4_index = np.asarray(d['quadtree_child_4_index'], U).copy()
        self.quadTree.quadTreeNode = q

    def loadFromFile(self, fileName):
        self.loadFromArrays(np.load(fileName))


# --------------------------------------------------------------------------- integrator pieces
def mis_weight(pdf_a, pdf_b):
    """src/path_guiding_integrator.py:16-24 (power heuristic, fma, NaN -> 0)."""
    a = np.asarray(pdf_a, F)
    b = np.asarray(pdf_b, F)
    with np.errstate(all='ignore'):
        a2 = (a * a).astype(F)
        # dr.fma(pdf_b, pdf_b, a2): single rounding of b*b + a2
        den = (b.astype(np.float64) * b.astype(np.float64) + a2.astype(np.float64)).astype(F)
        res = np.where(a > 0, a2 / den, F(0)).astype(F)
    res[np.isnan(res)] = 0
    return res


def nee_mis(bsdf_pdf_em, sdtree_pdf_em, pdf_with_delta, pdf_without_delta, ds_pdf, ds_delta,
            bsdfSamplingFraction, iteration):
    """src/path_guiding_integrator.py:241-253 -> (surface_pdf_em, mis_em)."""
    f = F(bsdfSamplingFraction)
    omf = F(1 - bsdfSamplingFraction)
    eps = F(0.00001)
    with np.errstate(all='ignore'):
        pdf_diffuse = ((pdf_with_delta + eps) / (pdf_without_delta + eps)).astype(F)
        surface = (f * bsdf_pdf_em + (omf * sdtree_pdf_em) * pdf_diffuse).astype(F)
    if iteration <= 1:
        surface = np.asarray(bsdf_pdf_em, F).copy()
    mis_em = np.where(ds_delta, F(1), mis_weight(ds_pdf, surface)).astype(F)
    return surface, mis_em


def mixture(bsdf_pdf, sdtree_pdf, bsdf_value, do_mis, bsdfSamplingFraction):
    """src/path_guiding_integrator.py:310-311 -> (woPdf, bsdf_weight) on do_mis lanes;
    other lanes keep (bsdf_pdf, bsdf_value / bsdf_pdf is NOT recomputed -> returned unchanged)."""
    f = F(bsdfSamplingFraction)
    omf = F(1 - bsdfSamplingFraction)
    with np.errstate(all='ignore'):
        mix = ((f * bsdf_pdf) + omf * sdtree_pdf).astype(F)
        woPdf = np.where(do_mis, mix, bsdf_pdf).astype(F)
        w = (bsdf_value / woPdf[:, None]).astype(F)
    return woPdf, w


def process_path_data(Lfinal, throughputRadiance, throughputBsdf, bsdf, max_depth):
    """src/path_guiding_integrator.py:434-453 -> (product (slots,3), radiance (slots,))."""
    slots = throughputRadiance.shape[0]
    ray = np.arange(slots, dtype=np.int64) // max_depth
    with np.errstate(all='ignore'):
        out = ((Lfinal[ray] - throughputRadiance) / throughputBsdf).astype(F)
        out[np.isnan(out)] = 0
        inc = (out / bsdf).astype(F)
        inc[np.isnan(inc)] = 0
        rad = dm.luminance(inc)
    return out, rad


def filter_records(active, radiance, radiance_nee, woPdf):
    """src/path_guiding_integrator.py:463-478 -> (keep mask, scrubbed radiance, scrubbed nee)."""
    radiance = np.asarray(radiance, F).copy()
    radiance_nee = np.asarray(radiance_nee, F).copy()
    radiance[np.isnan(radiance)] = 0
    radiance_nee[np.isnan(radiance_nee)] = 0
    both_zero = (radiance == 0) & (dm.luminance(radiance_nee) == 0)
    keep = active & ~both_zero & ~(woPdf == 0) & ~np.isnan(woPdf)
    return keep, radiance, radiance_nee


def refine_and_prepare(current, prev, iteration):
    """src/path_guiding_integrator.py:553-586."""
    current.setRefinementThreshold(iteration)
    current.refine()
    current.setQuadTreeRefinementThreshold()
    current.refineAllQuadTree()
    current.cleanUnusedQuadTree()
    prev.copyFrom(current)
    current.resetTreeVertCount()
    current.resetAllQuadTreeIrradiance()
