"""fp32 restatements of the third-party math (Dr.Jit sincos/atan2, Mitsuba luminance)
checked against float64 libm; see oracle/drjit_math.py for what is and is not claimed."""
import numpy as np

from oracle import drjit_math as dm

F = np.float32


def _ulp_err(approx, exact):
    exact32 = exact.astype(F)
    ulp = np.spacing(np.maximum(np.abs(exact32), F(1e-30)))
    return np.abs(approx.astype(np.float64) - exact) / ulp


def test_sincos_accuracy():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.random(200000) * 2 * np.pi, rng.random(1000) * 200 - 100,
                        [0.0, np.pi / 2, np.pi, 2 * np.pi]]).astype(F)
    s, c = dm.sincos(x)
    xd = x.astype(np.float64)
    # absolute error (results near zero cannot be relative-accurate in fp32)
    assert np.max(np.abs(s - np.sin(xd))) < 3e-7
    assert np.max(np.abs(c - np.cos(xd))) < 3e-7


def test_atan2_accuracy_and_quadrants():
    rng = np.random.default_rng(1)
    y = (rng.random(200000) * 2 - 1).astype(F)
    x = (rng.random(200000) * 2 - 1).astype(F)
    r = dm.atan2(y, x)
    exact = np.arctan2(y.astype(np.float64), x.astype(np.float64))
    assert np.max(np.abs(r - exact)) < 6e-7
    assert dm.atan2(F(0), F(0)) == 0
    assert dm.atan2(F(0), F(-1)) == dm.PI
    assert dm.atan2(F(-0.0), F(-1)) == dm.PI        # comparison-based sign, like Dr.Jit
    assert dm.atan2(F(1), F(0)) == dm.HALF_PI
    assert dm.atan2(F(-1), F(0)) == -dm.HALF_PI


def test_direction_roundtrip():
    rng = np.random.default_rng(2)
    p = rng.random((100000, 2)).astype(F)
    d = dm.canonical_to_dir(p)
    assert np.max(np.abs(np.linalg.norm(d.astype(np.float64), axis=1) - 1)) < 1e-6
    p2 = dm.dir_to_canonical(d)
    # phi is ill-conditioned at the poles; compare away from them
    ok = (p[:, 1] > 0.01) & (p[:, 1] < 0.99)
    dphi = np.abs(p2[ok, 0] - p[ok, 0])
    dphi = np.minimum(dphi, 1 - dphi)
    assert np.max(dphi) < 2e-6
    assert np.max(np.abs(p2[:, 1] - p[:, 1])) < 2e-7
    # src/common.py:270-279 self-test vector
    assert np.allclose(dm.dir_to_canonical(np.array([[0, 1, 0]], F)), [[0.25, 0.5]])
    assert dm.dir_to_canonical(np.array([[np.nan, 0, 0], [0, np.inf, 0]], F)).tolist() == [[0, 0], [0, 0]]
    assert dm.canonical_to_dir(np.zeros((1, 2), F)).tolist() == [[0, 0, -1]]
    # x can round to exactly 1.0 (SURVEY 8a A4')
    assert dm.dir_to_canonical(np.array([[1, -1e-9, 0]], F))[0, 0] <= 1.0


def test_luminance():
    assert abs(float(dm.luminance(np.array([1, 1, 1], F))) - 1.0) < 1e-6
