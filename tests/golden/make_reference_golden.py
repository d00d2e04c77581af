"""Writes tests/golden/reference_on_shim.npz: inputs and outputs of THE REFERENCE'S OWN SOURCE
(/root/reference/src/*.py, imported unmodified on the numpy Dr.Jit / Mitsuba stand-ins of oracle/refshim)
for a few splat / refine / query sequences.  /root/reference does not exist on the GPU box, so these
vectors are how the reference leg travels: tests/test_reference_golden.py replays the inputs through the
C ABI (libsdtree.so under `-m gpu`, the host emulation on CPU) and demands the stored outputs -- no oracle
in between.

    python tests/golden/make_reference_golden.py        (needs /root/reference)

Primitive semantics of the stand-ins are assumptions (oracle/refshim/drjit.py); everything above them --
control flow, tie rules, operation order, node numbering -- is the reference's code."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.refshim import as_oracle as ro          # noqa: E402
from oracle import drjit_math as dm                 # noqa: E402
import sdt_cases as cases                            # noqa: E402

F, U = np.float32, np.uint32

CONFIGS = [
    dict(name='unit_nee', lo=(0, 0, 0), hi=(1, 1, 1), kd=20, qd=20, nee=True, leaf=250, iters=3, n=4000, refine_last=False),
    dict(name='box_shallow', lo=(-3.5, 0.25, -1), hi=(2.25, 7, 0.5), kd=6, qd=5, nee=True, leaf=40, iters=3, n=3000, refine_last=True),
    dict(name='cube100', lo=(0, 0, 0), hi=(100, 100, 100), kd=12, qd=20, nee=False, leaf=60, iters=2, n=3000, refine_last=True),
]


def records(rng, c):
    """dyadic radiance / woPdf, and NEE radiance whose luminance is a power-of-two multiple of one constant: the
    fp32 sums do not depend on the order of the atomics, so the statistics and the post-refine topology are
    reproducible bit for bit -- with NON-ZERO NEE energy deposited"""
    n = c['n']
    lo, hi = np.asarray(c['lo'], F), np.asarray(c['hi'], F)
    ext = hi - lo
    pos = (lo + rng.random((n, 3)) ** 1.5 * ext).astype(F)
    d = rng.random((n, 2)).astype(F)
    m = rng.random(n) < 0.6
    d[m] = np.clip(np.stack([0.3 + 0.01 * rng.standard_normal(m.sum()), 0.7 + 0.003 * rng.standard_normal(m.sum())], 1), 0, 1).astype(F)
    rad = (rng.integers(0, 17, n) / 8.0).astype(F)
    wo = rng.choice(np.array([0.25, 0.5, 1.0, 2.0], F), n).astype(F)
    rec = ro.SurfaceInteractionRecord(pos, d, rad, wo)
    if c['nee']:
        rec.radiance_nee = cases.dyadic_nee(rng, n)      # non-zero, exactly dyadic luminance
        rec.direction_nee = rng.random((n, 2)).astype(F)
        k = rng.random(n) < 0.5
        rec.direction_nee[k] = np.clip(np.stack([0.8 + 0.02 * rng.standard_normal(k.sum()), 0.2 + 0.02 * rng.standard_normal(k.sum())], 1), 0, 1).astype(F)
    pos[:3] = [lo, hi, lo + ext * F(0.5)]
    pos[3] = hi + ext
    pos[4, 1] = np.nan
    d[5], d[6], d[7] = [0.5, 0.5], [1.0, 0.0], [1.5, 0.2]
    d[8, 0] = np.nan
    wo[9], wo[10], wo[11] = 0.0, -1.0, np.nan
    return rec


def main():
    out = {}
    rng = np.random.default_rng(20261018)
    for c in CONFIGS:
        p = c['name'] + '/'
        cur, prev = _pair(c)
        for k in ('lo', 'hi'):
            out[p + k] = np.asarray(c[k], F)
        out[p + 'cfg'] = np.array([c['kd'], c['qd'], int(c['nee']), c['leaf'], c['iters'], int(c['refine_last'])], np.int64)
        for it in range(c['iters']):
            rec = records(rng, c)
            for f in ('position', 'direction', 'radiance', 'woPdf', 'radiance_nee', 'direction_nee'):
                out[f'{p}it{it}/{f}'] = getattr(rec, f)
            cur.addDataPropagate(rec)
            if it == c['iters'] - 1 and not c['refine_last']:
                break
            cases.oracle_refine(cur, prev, c['leaf'])
        for tag, t in (('prev', prev), ('cur', cur)):
            for k, v in t.to_arrays().items():
                out[f'{p}{tag}/{k}'] = np.asarray(v)
        n = 2048
        lo, hi = np.asarray(c['lo'], F), np.asarray(c['hi'], F)
        ext = hi - lo
        pos = (lo - 0.01 * ext + rng.random((n, 3)) * ext * 1.02).astype(F)
        pos[:4] = [lo, hi, lo + ext * F(0.5), lo + ext * F(0.25)]
        pos[4, 2] = np.nan
        active = rng.random(n) < 0.9
        u = rng.random((n, 3 * (c['qd'] + 1))).astype(F)
        u[:64] = rng.integers(0, 5, (64, u.shape[1])) / F(4)
        d, pdf, dbg = prev.sample(pos, ro.ExplicitSampler(u=u), active, return_debug=True)
        dirs = rng.standard_normal((n, 3)).astype(F)
        dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
        dirs[:200] = dm.canonical_to_dir((rng.integers(0, 9, (200, 2)) / 8.0).astype(F))
        dirs[200:205] = [[0, 0, 1], [0, 0, -1], [np.nan, 0, 1], [0, 0, 0], [np.inf, 0, 0]]
        pp, pdbg = prev.pdf(pos, dirs, active, return_debug=True)
        out.update({p + 'q/pos': pos, p + 'q/active': active, p + 'q/u': u, p + 'q/dirs': dirs,
                    p + 'q/leaf': dbg['leaf'], p + 'q/root': dbg['root'], p + 'q/sample_node': dbg['sample_node'],
                    p + 'q/sample_dir': d, p + 'q/sample_pdf': pdf, p + 'q/pdf_node': pdbg['pdf_node'], p + 'q/pdf': pp})
        print(c['name'], 'kd nodes', prev.kdTreeNode.getWidth(), 'quad nodes', prev.quadTree.quadTreeNode.getWidth(),
              'max quad depth', int(prev.quadTree.quadTreeNode.depth.max()),
              'nee energy in cur', float(np.asarray(out[p + 'cur/quadtree_irradiance']).sum()))
    path = os.path.join(HERE, 'reference_on_shim.npz')
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), 'bytes')


def _pair(c):
    cur = ro.KDTree(maxDepth=c['kd'])
    cur.setup(c['lo'], c['hi'])
    cur.quadTree.maxDepth = c['qd']
    cur.quadTree.isStoreNEERadiance = c['nee']
    prev = ro.KDTree(maxDepth=c['kd'])
    prev.copyFrom(cur)
    return cur, prev


if __name__ == '__main__':
    main()
