"""The Mitsuba-facing plugin class (integrator.PathGuidingIntegrator) cannot meet a real Mitsuba in
this image.  This test runs it against tests/mitsuba_stub (numpy stand-ins for the few Dr.Jit /
Mitsuba names it touches, a one-box glowing "scene") so that its bounce loop, its record scatter
and every call it makes into PathGuidingCore / the C ABI execute: plugin registration, props,
training passes with splat + refine, guided passes (iteration > 1), final pass, variance / MSE,
save / load.  It checks plumbing, not radiometry."""
import importlib
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from hostemu.build_hostemu import build as build_hostemu  # noqa: E402


@pytest.fixture()
def plugin():
    stub = os.path.join(HERE, "mitsuba_stub")
    sys.path.insert(0, stub)
    for m in ("drjit", "mitsuba"):
        sys.modules.pop(m, None)
    import practical_path_guiding_lab_b200.integrator as integ
    integ = importlib.reload(integ)
    assert integ._HAVE_MITSUBA
    yield integ
    sys.path.remove(stub)
    for m in ("drjit", "mitsuba"):
        sys.modules.pop(m, None)
    importlib.reload(integ)


def test_plugin_runs_against_stub(plugin, tmp_path):
    import drjit as dr
    import mitsuba as mi
    assert 'path_guiding_integrator' in mi._registered
    with pytest.raises(Exception):
        mi._registered['path_guiding_integrator']({'max_depth': -3})
    it = mi._registered['path_guiding_integrator']({'max_depth': 6, 'rr_depth': 3})
    assert it.max_depth == 6 and it.aov_names() == ["depth.Y"] and it.to_string() == "path_guiding_integrator"
    # the library under the plugin: the host emulation (no GPU here); same C ABI
    it.core = plugin.PathGuidingCore(6, 3, lib_path=build_hostemu(), kd_capacity=1 << 12, quad_capacity=1 << 18)
    n = 4096
    it.setup(n, [0, 0, 0], [1, 1, 1], sdTreeMaxDepth=8, quadTreeMaxDepth=10, isStoreNEERadiance=True, bsdfSamplingFraction=0.5)
    scene = mi.StubScene()
    ray = mi.Ray3f(o=np.zeros((n, 3)), d=np.tile([[0, 0, 1.0]], (n, 1)))
    sizes = []
    for iteration in range(4):
        it.setIteration(iteration, False)
        it.resetVarianceCounter()
        for p in range(2):
            L, valid, aov = it.sample(scene, mi.StubSampler(n, seed=10 * iteration + p), ray)
            assert L.v.shape == (n, 3) and np.isfinite(L.v).all() and (L.v >= 0).all() and aov == [1]
            assert valid.v.shape == (n,) and valid.v.any()
        cur = it.core.tree.download(1)
        assert cur['kdtree_vertCount'][0] > 0 and cur['quadtree_irradiance'][0] > 0        # the passes splatted records
        v = it.computeVariance(2)
        assert np.isfinite(v) and v >= 0
        it.core.tree.set_max_leaf_size(200)
        it.core.tree.refine()                           # small threshold so that the tiny stub run refines at all
        sizes.append(it.core.tree.sizes())
    assert sizes[-1]['n_kd'] > 1 and sizes[-1]['n_quad'] > sizes[-1]['n_roots'] and sizes[-1]['error'] == 0
    # guided passes ran (iteration 2, 3 > 1): the sampled directions came from the tree
    f = str(tmp_path / "tree.npz")
    it.saveSDTreeToFile(f)
    it.saveSDTreeOBJ(str(tmp_path / "tree.obj"))
    it.loadSDTreeFromFile(f)
    # final iteration: no records, 2 spp per pass
    it.setIteration(4, True)
    it.resetVarianceCounter()
    ray2 = mi.Ray3f(o=np.zeros((2 * n, 3)), d=np.tile([[0, 0, 1.0]], (2 * n, 1)))
    L, valid, _ = it.sample(scene, mi.StubSampler(2 * n, spp=2, seed=99), ray2)
    assert L.v.shape == (2 * n, 3) and np.isfinite(L.v).all()
    assert it.sumL.v.shape == (n, 3)
    gt = mi.Spectrum(dr.Float(np.full(n, 0.3)))
    assert np.isfinite(it.computeMSE(2, gt)) and np.isfinite(it.computeVariance(2, gt)) and np.isfinite(it.computeVariance(2))
