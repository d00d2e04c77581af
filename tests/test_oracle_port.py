"""The C + OpenMP restatement of the reference's per-vertex operations (oracle/sdtree_port.c,
the CPU baseline of bench.py) against the numpy oracle: node ids, directions and pdfs bit-exact,
splatted statistics equal (dyadic energies)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sdt_cases as cases  # noqa: E402
from oracle import sdtree_oracle as so  # noqa: E402
from oracle.port import PortTree  # noqa: E402

F = np.float32


def _trained_oracle():
    cur, prev = cases.oracle_pair()
    lobes = [((0.3, 0.7, 0.02),), ((0.3, 0.7, 0.004), (0.8, 0.2, 0.05)), ((0.8, 0.2, 0.01),)]
    for it in range(3):
        cur.addDataPropagate(cases.dyadic_records(20000, 100 + it, lobes[it]))
        cases.oracle_refine(cur, prev, 600)
    return cur, prev


def test_port_matches_numpy_oracle():
    cur, prev = _trained_oracle()
    rng = np.random.default_rng(1)
    n = 5000
    pos = (rng.random((n, 3)) * 1.02 - 0.01).astype(F)
    pos[:3] = [[0.5, 0.5, 0.5], [np.nan, 0, 0], [1, 1, 1]]
    active = rng.random(n) < 0.9
    pt = PortTree(prev.to_arrays())
    d, p, dbg = pt.sample(pos, seed=11, lane_offset=5, active=active, debug=True)
    od, op, odbg = prev.sample(pos, so.ExplicitSampler(seed=11, n=n, lane_offset=5), active, return_debug=True)
    a = active
    assert np.array_equal(dbg[a, 0], odbg['leaf'][a]) and np.array_equal(dbg[a, 2], odbg['sample_node'][a])
    assert np.array_equal(dbg[a, 3], odbg['pdf_node'][a])
    assert cases.beq(d, od) and cases.beq(p, op)
    dirs = rng.standard_normal((n, 3)).astype(F)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    dirs[:4] = [[1, 0, 0], [0, 0, 1], [0, 0, 0], [np.nan, 0, 1]]
    pp = pt.pdf(pos, dirs, active)
    assert cases.beq(pp, prev.pdf(pos, dirs, active))
    # splat into a copy of current (zero statistics): every visited node, like the reference
    rec = cases.dyadic_records(30000, 9, ((0.3, 0.7, 0.02),))
    ct = PortTree(cur.to_arrays())
    ct.splat(rec.position, rec.direction, rec.radiance, rec.woPdf)
    cur.addDataPropagate(rec)
    assert np.array_equal(ct.a['kd_count'], cur.kdTreeNode.vertCount)
    assert np.array_equal(ct.a['q_energy'], cur.quadTree.quadTreeNode.irradiance)
    assert ct.threads() >= 1
