#!/usr/bin/env python
"""Records golden vectors from the REAL reference (Mitsuba 3 <= 3.5 / Dr.Jit 0.4, as the reference repo pins them):

    python tools/record_reference_fixtures.py --reference /path/to/practical_path_guiding_lab [--variant llvm_ad_rgb]

It replays the INPUTS stored in tests/golden/reference_on_shim.npz (the records of every training iteration, the query
positions, explicit uniforms and directions) through the reference's own KDTree (src/kdtree.py, imported unmodified)
running on real Dr.Jit, and writes tests/golden/reference_real.npz with the same keys.  tests/test_reference_golden.py
picks that file up when it exists, and the differences to the shim-produced file -- i.e. to the assumptions
oracle/refshim makes about Dr.Jit's primitives (gather / scatter_reduce / masked assignment / sincos / atan2 / the
lowering of `E / 100`) -- are printed key by key.  (SURVEY.md 8c item vii.)

NOT RUN in the build container: Mitsuba and Dr.Jit are not installable there (no network, not in the wheelhouse).
Only public Dr.Jit / Mitsuba API is used; the duck-typed sampler serves the recorded uniforms in the order the
reference draws them (next_2d at the leaf, next_1d per level: src/quadtree.py:956,980).
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F, U = np.float32, np.uint32


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", required=True, help="checkout of takkasila/practical_path_guiding_lab")
    ap.add_argument("--variant", default="llvm_ad_rgb")
    ap.add_argument("--shim-golden", default=os.path.join(ROOT, "tests", "golden", "reference_on_shim.npz"))
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "reference_real.npz"))
    a = ap.parse_args()
    import drjit as dr
    import mitsuba as mi
    mi.set_variant(a.variant)
    sys.path.insert(0, a.reference)
    from src.kdtree import KDTree                      # the reference's own classes
    from src.common import SurfaceInteractionRecord

    class TableSampler:
        """uniforms u[lane, k] in the order the reference consumes them: 3 per visited level (x, y, select)"""

        def __init__(self, u):
            self.u = mi.Float(np.ascontiguousarray(u.reshape(-1)))
            self.stride = u.shape[1]
            self.lane = dr.arange(mi.UInt32, u.shape[0])
            self.cursor = dr.zeros(mi.UInt32, u.shape[0])

        def _take(self, active=True):
            idx = dr.minimum(self.cursor, self.stride - 1)
            v = dr.gather(mi.Float, self.u, self.lane * self.stride + idx, active)
            self.cursor = dr.select(mi.Bool(active), self.cursor + 1, self.cursor)
            return v

        def next_1d(self, active=True):
            return self._take(active)

        def next_2d(self, active=True):
            x = self._take(active)
            return mi.Point2f(x, self._take(active))

    def record(z, p, it):
        n = z[f"{p}it{it}/position"].shape[0]
        r = dr.zeros(SurfaceInteractionRecord, n)
        r.position = mi.Vector3f(*[mi.Float(np.ascontiguousarray(z[f"{p}it{it}/position"][:, k])) for k in range(3)])
        r.direction = mi.Vector2f(*[mi.Float(np.ascontiguousarray(z[f"{p}it{it}/direction"][:, k])) for k in range(2)])
        r.radiance = mi.Float(z[f"{p}it{it}/radiance"])
        r.woPdf = mi.Float(z[f"{p}it{it}/woPdf"])
        r.radiance_nee = mi.Color3f(*[mi.Float(np.ascontiguousarray(z[f"{p}it{it}/radiance_nee"][:, k])) for k in range(3)])
        r.direction_nee = mi.Vector2f(*[mi.Float(np.ascontiguousarray(z[f"{p}it{it}/direction_nee"][:, k])) for k in range(2)])
        r.active = mi.Bool(np.ones(n, bool))
        return r

    def arrays(tree, tmp):
        tree.saveToFile(tmp)                           # the reference's own npz schema (src/kdtree.py:539-603)
        return dict(np.load(tmp))

    z = np.load(a.shim_golden)
    names = sorted({k.split("/")[0] for k in z.files})
    out = {}
    tmp = a.out + ".tmp.npz"
    for name in names:
        p = name + "/"
        kd, qd, nee, leaf, iters, refine_last = [int(v) for v in z[p + "cfg"]]
        cur = KDTree(maxDepth=kd)
        cur.setup(mi.Vector3f(*[float(v) for v in z[p + "lo"]]), mi.Vector3f(*[float(v) for v in z[p + "hi"]]))
        cur.quadTree.maxDepth = qd
        cur.quadTree.isStoreNEERadiance = bool(nee)
        prev = KDTree(maxDepth=kd)
        prev.copyFrom(cur)
        for it in range(iters):
            cur.addDataPropagate(record(z, p, it))
            if it == iters - 1 and not refine_last:
                break
            cur.maxLeafSize = leaf                     # refineAndPrepareSDTreeForNextIteration with an explicit threshold
            cur.refine()
            cur.setQuadTreeRefinementThreshold()
            cur.refineAllQuadTree()
            cur.cleanUnusedQuadTree()
            prev.copyFrom(cur)
            cur.resetTreeVertCount()
            cur.resetAllQuadTreeIrradiance()
        for tag, t in (("prev", prev), ("cur", cur)):
            for k, v in arrays(t, tmp).items():
                out[f"{p}{tag}/{k}"] = v
        pos, act, u, dirs = z[p + "q/pos"], z[p + "q/active"], z[p + "q/u"], z[p + "q/dirs"]
        P = mi.Vector3f(*[mi.Float(np.ascontiguousarray(pos[:, k])) for k in range(3)])
        A = mi.Bool(act)
        d, pdf = prev.sample(P, TableSampler(u), A)
        D = mi.Vector3f(*[mi.Float(np.ascontiguousarray(dirs[:, k])) for k in range(3)])
        pp = prev.pdf(P, D, A)
        out[p + "q/leaf"] = np.array(prev.getLeafNodeIndex(P, A)).astype(U)
        out[p + "q/sample_dir"] = np.stack([np.array(d[k]) for k in range(3)], 1).astype(F)
        out[p + "q/sample_pdf"] = np.array(pdf).astype(F)
        out[p + "q/pdf"] = np.array(pp).astype(F)
    if os.path.exists(tmp):
        os.remove(tmp)
    np.savez_compressed(a.out, **out)
    bad = 0
    for k in sorted(out):
        if k not in z.files:
            continue
        x, y = np.asarray(out[k]), np.asarray(z[k])
        same = x.shape == y.shape and bool(np.all((x == y) | ((x != x) & (y != y))))
        if not same:
            bad += 1
            print("DIFFERS from the shim leg:", k, x.shape, y.shape)
    print(a.out, "written;", "identical to the shim leg on every shared key" if not bad else f"{bad} keys differ")


if __name__ == "__main__":
    main()
