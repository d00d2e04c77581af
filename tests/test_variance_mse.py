"""computeVariance / computeMSE (src/path_guiding_integrator.py:503-550) against hand-computed numbers.

Film of three pixels, two samples per pixel:
  A: (1,0,0), (3,0,0)     sumL (4,0,0)   sumL2 (10,0,0)      mean 2, mean of squares 5
  B: (0,2,0), (0,2,0)     sumL (0,4,0)   sumL2 (0,8,0)       mean 2, mean of squares 4
  C: (0,0,0), (1000,0,0)  sumL (1e3,0,0) sumL2 (1e6,0,0)     mean 500, mean of squares 5e5   -> above the 10000 cut-off
luminance weights (0.212671, 0.715160, 0.072169).
  variance, no ground truth:  A 5-4 = 1 -> 0.212671;  B 4-4 = 0;  C 5e5-2.5e5 -> clamped to 10000
                              mean = (0.212671 + 0 + 10000) / 3, divided by spp-1 = 1
  ground truth A (2,0,0), B (0,1,0), C (500,0,0):
  variance vs ground truth:   A 5-4 = 1 -> 0.212671;  B 4-1 = 3 -> 2.14548;  C 5e5-2.5e5 -> 10000;  mean / spp
  MSE vs ground truth:        A 0;  B (2-1)^2 = 1 -> 0.71516;  C 0;  mean = 0.71516 / 3
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

SUM_L = np.array([[4, 0, 0], [0, 4, 0], [1000, 0, 0]], np.float32)
SUM_L2 = np.array([[10, 0, 0], [0, 8, 0], [1e6, 0, 0]], np.float32)
GT = np.array([[2, 0, 0], [0, 1, 0], [500, 0, 0]], np.float32)
SPP = 2
WANT_VAR = (0.212671 + 0.0 + 10000.0) / 3 / (SPP - 1)
WANT_VAR_GT = (0.212671 + 3 * 0.715160 + 10000.0) / 3 / SPP
WANT_MSE = 0.715160 / 3
# the same sums read as ONE sample per pixel: mean = sumL, no division by spp - 1 (:545-547); C: 1e6 - 1000^2 = 0
WANT_VAR_1SPP = (0.212671 * (10 - 16) + 0.715160 * (8 - 16) + 0.0) / 3


def test_cornell_stand_in_accumulators():
    import torch
    from hostemu.build_hostemu import build as build_hostemu
    from practical_path_guiding_lab_b200.cornell import CornellBox
    r = CornellBox(3, 1, max_depth=4, device="cpu", lib_path=build_hostemu(), kd_capacity=64, quad_capacity=256)
    r.setup(sdTreeMaxDepth=4, quadTreeMaxDepth=4)
    r.sumL[:] = torch.from_numpy(SUM_L)
    r.sumL2[:] = torch.from_numpy(SUM_L2)
    gt = torch.from_numpy(GT)
    assert r.computeVariance(SPP) == pytest.approx(WANT_VAR, rel=1e-6)
    assert r.computeVariance(SPP, gt) == pytest.approx(WANT_VAR_GT, rel=1e-6)
    assert r.computeMSE(SPP, gt) == pytest.approx(WANT_MSE, rel=1e-6)
    assert r.computeVariance(1) == pytest.approx(WANT_VAR_1SPP, rel=1e-6)


def test_plugin_accumulators_on_stub():
    import importlib
    stub = os.path.join(HERE, "mitsuba_stub")
    sys.path.insert(0, stub)
    for m in ("drjit", "mitsuba"):
        sys.modules.pop(m, None)
    import practical_path_guiding_lab_b200.integrator as integ
    integ = importlib.reload(integ)
    try:
        import drjit as dr
        import mitsuba as mi
        it = mi._registered['path_guiding_integrator']({'max_depth': 4, 'rr_depth': 2})
        it.sumL = mi.Spectrum(dr.Float(SUM_L[:, 0]), dr.Float(SUM_L[:, 1]), dr.Float(SUM_L[:, 2]))
        it.sumL2 = mi.Spectrum(dr.Float(SUM_L2[:, 0]), dr.Float(SUM_L2[:, 1]), dr.Float(SUM_L2[:, 2]))
        gt = mi.Spectrum(dr.Float(GT[:, 0]), dr.Float(GT[:, 1]), dr.Float(GT[:, 2]))
        assert float(it.computeVariance(SPP)) == pytest.approx(WANT_VAR, rel=1e-6)
        assert float(it.computeVariance(SPP, gt)) == pytest.approx(WANT_VAR_GT, rel=1e-6)
        assert float(it.computeMSE(SPP, gt)) == pytest.approx(WANT_MSE, rel=1e-6)
        assert float(it.computeVariance(1)) == pytest.approx(WANT_VAR_1SPP, rel=1e-6)
    finally:
        sys.path.remove(stub)
        for m in ("drjit", "mitsuba"):
            sys.modules.pop(m, None)
        importlib.reload(integ)


def test_reference_methods_give_the_hand_computed_numbers():
    """the reference's own computeVariance / computeMSE (unmodified source on the numpy stand-ins) on the same film"""
    from oracle import refshim
    ref = refshim.load_reference()
    if ref is None:
        pytest.skip("reference tree not present")
    mi, dr = ref.mi, ref.dr
    cls = ref.integrator.PathGuidingIntegrator
    it = cls.__new__(cls)                                    # the accumulators are all these methods read
    col = lambda a: mi.Color3f(dr.cuda.ad.Float(a[:, 0]) if hasattr(dr, "cuda") and hasattr(dr.cuda, "ad") else mi.Float(a[:, 0]),
                               mi.Float(a[:, 1]), mi.Float(a[:, 2]))
    it.sumL, it.sumL2 = col(SUM_L), col(SUM_L2)
    gt = col(GT)
    assert float(it.computeVariance(SPP)) == pytest.approx(WANT_VAR, rel=1e-6)
    assert float(it.computeVariance(SPP, gt)) == pytest.approx(WANT_VAR_GT, rel=1e-6)
    assert float(it.computeMSE(SPP, gt)) == pytest.approx(WANT_MSE, rel=1e-6)
    assert float(it.computeVariance(1)) == pytest.approx(WANT_VAR_1SPP, rel=1e-6)
