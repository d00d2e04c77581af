"""Generates tests/golden/sdtree_golden.npz with the ORACLE (oracle/sdtree_oracle.py): a small SD-tree
trained for 3 iterations on seeded dyadic records (post-refine arrays in the reference's npz schema),
plus seeded queries and the oracle's answers (leaf / root / node ids, directions, pdfs, splat
statistics).  The CPU and GPU parity suites replay the queries through libsdtree and compare with
these committed answers, so a drift of either the oracle or the kernels shows up against a fixed
file.  (The reference itself cannot be imported here -- no Mitsuba / Dr.Jit -- see SURVEY.md 8c.)

    python tests/golden/make_sdtree_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import sdt_cases as cases  # noqa: E402
from oracle import sdtree_oracle as so  # noqa: E402

F = np.float32


def main():
    cur, prev = cases.oracle_pair(store_nee=False)
    lobes = [((0.3, 0.7, 0.02),), ((0.3, 0.7, 0.004), (0.8, 0.2, 0.05)), ((0.8, 0.2, 0.01),)]
    for it in range(3):
        cur.addDataPropagate(cases.dyadic_records(6000, 500 + it, lobes[it]))
        cases.oracle_refine(cur, prev, 300)
    rng = np.random.default_rng(2024)
    n = 2048
    pos = (rng.random((n, 3)) * 1.02 - 0.01).astype(F)
    pos[:6] = [[0.5, 0.5, 0.5], [0, 0, 0], [1, 1, 1], [np.nan, 0.5, 0.5], [0.25, 0.5, 0.75], [1.5, 0.5, 0.5]]
    active = rng.random(n) < 0.9
    u = rng.random((n, 3 * 22)).astype(F)
    d, p, dbg = prev.sample(pos, so.ExplicitSampler(u=u), active, return_debug=True)
    dc, pc = prev.sample(pos, so.ExplicitSampler(seed=77, n=n, lane_offset=3), active)
    dirs = rng.standard_normal((n, 3)).astype(F)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    dirs[:6] = [[1, 0, 0], [0, 1, 0], [0, 0, 1], [0, 0, -1], [0, 0, 0], [np.inf, 0, 0]]
    pp, pdbg = prev.pdf(pos, dirs, active, return_debug=True)
    rec = cases.dyadic_records(5000, 900, ((0.55, 0.5, 0.01),))
    cur.addDataPropagate(rec)
    tree = {f"tree_{k}": np.asarray(v) for k, v in prev.to_arrays().items()}
    np.savez_compressed(os.path.join(HERE, "sdtree_golden.npz"), pos=pos, active=active, u=u, dirs=dirs,
                        leaf=dbg['leaf'], root=dbg['root'], sample_node=dbg['sample_node'], pdf_node=dbg['pdf_node'],
                        sample_dir=d, sample_pdf=p, counter_dir=dc, counter_pdf=pc, pdf=pp, pdf_query_node=pdbg['pdf_node'],
                        rec_position=rec.position, rec_direction=rec.direction, rec_radiance=rec.radiance, rec_wo_pdf=rec.woPdf,
                        splat_vert_count=cur.kdTreeNode.vertCount, splat_irradiance=cur.quadTree.quadTreeNode.irradiance, **tree)
    print("nodes", prev.kdTreeNode.getWidth(), prev.quadTree.quadTreeNode.getWidth())


if __name__ == "__main__":
    main()
