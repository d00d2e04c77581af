"""ORACLE (test infrastructure).  The frozen SD-tree of the synthetic microbench (BASELINE.json configs[1],
SURVEY.md 8d: "built by the oracle running splat + refine for K training iterations"), built ONCE by the numpy
restatement of the reference and used as INPUT DATA by both arms of bench.py and by the config-2-scale parity
test: the CUDA arm uploads it through sdt_upload (the npz schema of KDTree.saveToFile), the CPU arm hands it to
the C port.  Nothing here is timed or shipped; the tree is workload, like the synthetic records."""
import hashlib
import os
import tempfile

import numpy as np

from oracle import sdtree_oracle as so


def _schedule():
    from practical_path_guiding_lab_b200 import synthetic as syn        # seeded record generator + schedule (numpy only)
    return syn, syn.build_schedule()


def build():
    """-> (current, prev) oracle KDTrees after the build schedule"""
    syn, schedule = _schedule()
    cur = so.KDTree(maxDepth=20)
    cur.setup([0, 0, 0], [1, 1, 1])
    cur.quadTree.maxDepth = 20
    cur.quadTree.isStoreNEERadiance = False
    prev = so.KDTree(maxDepth=20)
    prev.copyFrom(cur)
    scene = syn.Scene()
    for it in schedule:
        r = scene.records(it['seed'], it['n'])
        cur.addDataPropagate(so.SurfaceInteractionRecord(r['position'], r['direction'], r['radiance'], r['wo_pdf']))
        cur.maxLeafSize = it['max_leaf_size']
        cur.refine()
        cur.setQuadTreeRefinementThreshold()
        cur.refineAllQuadTree()
        cur.cleanUnusedQuadTree()
        prev.copyFrom(cur)
        cur.resetTreeVertCount()
        cur.resetAllQuadTreeIrradiance()
    return cur, prev


def frozen_tree_arrays(cache=True):
    """the 23 npz arrays of the frozen tree; cached under the system temp directory (keyed by the schedule and
    the oracle source) so that the two arms of one bench run, and the ranks of one box, build it once"""
    syn, schedule = _schedule()
    key = hashlib.sha256()
    key.update(repr(schedule).encode())
    for f in (so.__file__, syn.__file__):
        key.update(open(f, 'rb').read())
    path = os.path.join(tempfile.gettempdir(), f"sdt_frozen_tree_{key.hexdigest()[:16]}.npz")
    if cache and os.path.exists(path):
        try:
            return dict(np.load(path))
        except Exception:
            pass
    _, prev = build()
    arrays = {k: np.asarray(v) for k, v in prev.to_arrays().items()}
    if cache:
        try:
            fd, tmp = tempfile.mkstemp(suffix=".npz", dir=os.path.dirname(path))
            os.close(fd)
            np.savez(tmp, **arrays)
            os.replace(tmp, path)
        except Exception:
            pass
    return arrays
