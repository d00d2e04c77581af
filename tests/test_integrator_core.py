"""Host logic of the drop-in integrator (practical_path_guiding_lab_b200/integrator.py)
against the oracle's restatement of src/path_guiding_integrator.py: record scatter at
ray*max_depth+depth, end-of-pass processPathData+filter+splat, the guided/BSDF choice, the
refine orchestration, npz/OBJ output.  Runs on the host emulation of the kernels (CPU)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sdt_cases as cases  # noqa: E402
from hostemu.build_hostemu import build as build_hostemu  # noqa: E402
from oracle import sdtree_oracle as so  # noqa: E402
from oracle import drjit_math as dm  # noqa: E402
from practical_path_guiding_lab_b200.integrator import PathGuidingCore  # noqa: E402

F = np.float32


@pytest.fixture(scope="module")
def lib():
    return build_hostemu()


def test_props_validation(lib):
    with pytest.raises(Exception):
        PathGuidingCore(max_depth=-2, lib_path=lib)
    with pytest.raises(Exception):
        PathGuidingCore(rr_depth=-1, lib_path=lib)
    assert PathGuidingCore(max_depth=-1, lib_path=lib).max_depth == -1


def test_pass_record_splat_refine(lib, tmp_path):
    rng = np.random.default_rng(2)
    rays, md = 4000, 4
    core = PathGuidingCore(max_depth=md, rr_depth=8, lib_path=lib, kd_capacity=1 << 12, quad_capacity=1 << 16)
    core.setup(rays, [-1e-4] * 3, [1 + 1e-4] * 3, sdTreeMaxDepth=20, quadTreeMaxDepth=20, isStoreNEERadiance=True)
    cur, prev = cases.oracle_pair([F(-1e-4)] * 3, [F(1 + 1e-4)] * 3, 20, 20, True)
    for iteration in range(3):
        core.setIteration(iteration, False)
        core.resetRayPathData(np.zeros(1, F))
        orec = {k: np.zeros_like(v[:core.array_size]) for k, v in core.record.items()}      # (the record has one spare slot)
        L = np.zeros((rays, 3), F)
        thr = np.ones((rays, 3), F)
        ray_index = np.arange(rays)
        alive = np.ones(rays, bool)
        for d in range(md):
            depth = np.full(rays, d)
            pos = rng.random((rays, 3)).astype(F)
            wo = rng.standard_normal((rays, 3)).astype(F)
            wo /= np.linalg.norm(wo, axis=1, keepdims=True)
            nee_d = rng.standard_normal((rays, 3)).astype(F)
            nee_d /= np.linalg.norm(nee_d, axis=1, keepdims=True)
            w = (rng.random((rays, 3)) * 1.5).astype(F)
            nee = (rng.random((rays, 3)) * (rng.random((rays, 1)) < 0.3)).astype(F)
            wopdf = (rng.random(rays) * 2).astype(F)
            wopdf[rng.random(rays) < 0.05] = 0
            L = (L + thr * nee + thr * (rng.random((rays, 3)) < 0.02) * 5).astype(F)
            core.store_vertex(ray_index, depth, alive, pos, wo, w, thr, L, nee, nee_d, wopdf)
            gi = (ray_index * md + d)[alive]
            orec['position'][gi] = pos[alive]
            orec['direction'][gi] = dm.dir_to_canonical(wo[alive])
            orec['active'][gi] = 1
            orec['bsdf'][gi] = w[alive]
            orec['throughputBsdf'][gi] = thr[alive]
            orec['throughputRadiance'][gi] = L[alive]
            orec['radiance_nee'][gi] = nee[alive]
            orec['direction_nee'][gi] = dm.dir_to_canonical(nee_d[alive])
            orec['woPdf'][gi] = wopdf[alive]
            thr = (thr * w).astype(F)
            alive = alive & (rng.random(rays) < 0.8)
        for k in orec:
            assert np.array_equal(core.record[k][:core.array_size], orec[k]), k
        core.end_of_pass(L)
        _, rad = so.process_path_data(L, orec['throughputRadiance'], orec['throughputBsdf'], orec['bsdf'], md)
        keep, rad, nee = so.filter_records(orec['active'].astype(bool), rad, orec['radiance_nee'], orec['woPdf'])
        cur.addDataPropagate(so.SurfaceInteractionRecord(orec['position'][keep], orec['direction'][keep], rad[keep],
                                                         orec['woPdf'][keep], nee[keep], orec['direction_nee'][keep]), exact=True)
        got = core.tree.download(1)
        np.testing.assert_array_equal(got['kdtree_vertCount'], cur.kdTreeNode.vertCount)
        np.testing.assert_allclose(got['quadtree_irradiance'], cur.quadTree.quadTreeNode.irradiance, rtol=1e-4, atol=1e-6)
        # refine from the oracle's (exactly rounded) statistics so that both sides see the same buffers
        core.tree.upload_stats(cur.quadTree.quadTreeNode.irradiance, cur.kdTreeNode.vertCount)
        core.tree.set_max_leaf_size(500)
        core.tree.refine()
        cases.oracle_refine(cur, prev, 500)
        cases.assert_tree_equal(core.tree.download(0), prev)
    # iteration threshold of the real entry point: 12000*sqrt(2^k)
    core.setIteration(3, False)
    core.refineAndPrepareSDTreeForNextIteration()
    assert np.float32(core.tree.download(0)['kdtree_maxLeafSize']) == np.float32(12000 * np.sqrt(2.0 ** 3))
    # guided / BSDF choice on the trained tree
    core.setIteration(2, False)
    n = 3000
    pos = rng.random((n, 3)).astype(F)
    wo = rng.standard_normal((n, 3)).astype(F)
    wo /= np.linalg.norm(wo, axis=1, keepdims=True)
    do_mis = rng.random(n) < 0.8
    cu = rng.random(n).astype(F)
    mode, d, sp = core.choose_and_sample(pos, wo, do_mis, cu, seed=5)
    assert np.array_equal(mode == 1, do_mis & (cu > 0.5)) and np.array_equal(mode == 2, do_mis & ~(cu > 0.5))
    pv = so.KDTree()
    pv.loadFromArrays(core.tree.download(0))
    od, op = pv.sample(pos, so.ExplicitSampler(seed=5, n=n), mode == 1)
    assert cases.beq(d[mode == 1], od[mode == 1]) and cases.beq(sp[mode == 1], op[mode == 1])
    assert cases.beq(sp[mode == 2], pv.pdf(pos, wo, mode == 2)[mode == 2])
    # files
    f = str(tmp_path / "t.npz")
    core.saveSDTreeToFile(f)
    assert set(np.load(f).files) == set(so.KDTree.NPZ_KEYS)
    core.loadSDTreeFromFile(f)
    obj = str(tmp_path / "scene_iter-1.obj")
    core.saveSDTreeOBJ(obj)
    lines = open(obj).read().splitlines()
    nk = core.tree.sizes()['n_kd']
    assert lines[0] == '# OBJ file of KDTree Bounding Boxes' and lines[1] == 'o scene_iter-1'
    assert sum(l.startswith('v ') for l in lines) == 8 * nk and sum(l.startswith('l ') for l in lines) == 6 * nk
