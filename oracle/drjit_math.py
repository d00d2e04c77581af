"""ORACLE (test infrastructure only -- never imported by the product path).

fp32 restatement of the third-party math the reference's hot path calls and
that is absent from /root/reference:

  * Dr.Jit 0.4.x (un-vendored, un-pinned dependency; README.md:7-11 only names
    the package): ``dr.sincos``, ``dr.atan2``, ``dr.sqrt``, fp32 ``/``
    -- call sites src/common.py:110,115,142,152-153.
  * Mitsuba 3.0-3.5 ``mi.luminance(Color3f)`` -- call sites
    src/path_guiding_integrator.py:452,471 and src/quadtree.py:461.

Dr.Jit's sincos/atan2 are the CEPHES single-precision kernels (octant range
reduction with a three-part pi/4, degree-3 minimax polynomials).  That published
algorithm is restated here with every operation a separately rounded IEEE fp32
multiply/add (numpy has no fused multiply-add), so the CUDA library -- built with
-fmad=false, IEEE div/sqrt -- reproduces these functions BIT FOR BIT.  They are
NOT claimed bit-identical to Dr.Jit itself (Dr.Jit contracts to FMA and, on its
CUDA backend, may use approximate div/sqrt): this file is the one part of the
oracle the reference-on-stand-ins leg cannot pin (the stand-ins call it); the north_star tolerance (1e-5 relative on directions/pdfs) is what a real
Dr.Jit run would be held to.  tests/test_oracle_math.py checks these against
float64 libm to a few ulp.
"""
import numpy as np

F = np.float32

# CEPHES sinf/cosf constants
FOPI = F(1.27323954473516)          # 4/pi
DP1 = F(0.78515625)
DP2 = F(2.4187564849853515625e-4)
DP3 = F(3.77489497744594108e-8)
S0, S1, S2 = F(-1.6666654611e-1), F(8.3321608736e-3), F(-1.9515295891e-4)
C0, C1, C2 = F(4.166664568298827e-2), F(-1.388731625493765e-3), F(2.443315711809948e-5)
# CEPHES atanf constants
T3P8 = F(0.4142135623730950)        # tan(pi/8)
A0, A1, A2, A3 = F(8.05374449538e-2), F(-1.38776856032e-1), F(1.99777106478e-1), F(-3.33329491539e-1)

PI = F(np.pi)
TWO_PI = F(2.0 * np.pi)             # dr.two_pi  (== f32(2.0 * dr.pi))
HALF_PI = F(0.5 * np.pi)
QUARTER_PI = F(0.25 * np.pi)
INV_FOUR_PI = F(1.0 / (4.0 * np.pi))  # dr.inv_four_pi

# mi.luminance Rec.709 weights (Mitsuba constant)
LUM_R, LUM_G, LUM_B = F(0.212671), F(0.715160), F(0.072169)


def _f(x):
    return np.asarray(x, dtype=F)


def sincos(x):
    """(sin x, cos x), fp32, CEPHES octant reduction.  Accurate for |x| < 8192."""
    x = _f(x)
    with np.errstate(all='ignore'):
        xa = np.abs(x)
        j = (xa * FOPI).astype(np.uint32)           # truncation
        j = (j + np.uint32(1)) & np.uint32(0xFFFFFFFE)
        y = j.astype(F)
        xr = ((xa - y * DP1) - y * DP2) - y * DP3
        z = xr * xr
        ps = (((S2 * z + S1) * z + S0) * z) * xr + xr
        pc = ((((C2 * z + C1) * z + C0) * z) * z - F(0.5) * z) + F(1.0)
        swap = (j & np.uint32(2)) != 0
        s = np.where(swap, pc, ps)
        c = np.where(swap, ps, pc)
        neg_s = ((j & np.uint32(4)) != 0) ^ (x < 0)
        neg_c = ((j + np.uint32(2)) & np.uint32(4)) != 0
        s = np.where(neg_s, -s, s)
        c = np.where(neg_c, -c, c)
    return s.astype(F), c.astype(F)


def atan2(y, x):
    """atan2 in (-pi, pi], fp32.  Structure of Dr.Jit's atan2 (min/max ratio,
    quadrant fix-up by comparisons, (0,0) -> 0) around the CEPHES atanf kernel."""
    y = _f(y)
    x = _f(x)
    with np.errstate(all='ignore'):
        ax = np.abs(x)
        ay = np.abs(y)
        mn = np.minimum(ax, ay)
        mx = np.maximum(ax, ay)
        a = mn / mx
        big = a > T3P8
        t = np.where(big, (a - F(1.0)) / (a + F(1.0)), a)
        base = np.where(big, QUARTER_PI, F(0.0))
        z = t * t
        p = ((((A0 * z + A1) * z + A2) * z + A3) * z) * t + t
        r = base + p
        r = np.where(ay > ax, HALF_PI - r, r)
        r = np.where(x < 0, PI - r, r)
        r = np.where(y < 0, -r, r)
        r = np.where(mx == 0, F(0.0), r)
    return r.astype(F)


def luminance(rgb):
    """mi.luminance(Color3f): ((r*wr) + (g*wg)) + (b*wb), fp32."""
    rgb = _f(rgb)
    with np.errstate(all='ignore'):
        return ((rgb[..., 0] * LUM_R + rgb[..., 1] * LUM_G) + rgb[..., 2] * LUM_B).astype(F)


def canonical_to_dir(p):
    """src/common.py:100-129.  p (n,2) -> d (n,3)."""
    p = _f(p)
    with np.errstate(all='ignore'):
        cos_theta = F(2.0) * p[:, 1] - F(1.0)
        sin_theta = np.sqrt(F(1.0) - cos_theta * cos_theta)
        phi = TWO_PI * p[:, 0]
        sin_phi, cos_phi = sincos(phi)
        d = np.empty((p.shape[0], 3), dtype=F)
        d[:, 0] = sin_theta * cos_phi
        d[:, 1] = sin_theta * sin_phi
        d[:, 2] = cos_theta
    return d


def dir_to_canonical(d):
    """src/common.py:132-158.  d (n,3) -> p (n,2); non-finite d -> (0,0)."""
    d = _f(d)
    with np.errstate(all='ignore'):
        cos_theta = np.minimum(np.maximum(d[:, 2], F(-1.0)), F(1.0))   # dr.clip
        phi = atan2(d[:, 1], d[:, 0])
        # loop "rotate phi" (src/common.py:148-150)
        while True:
            neg = phi < 0
            if not neg.any():
                break
            phi = np.where(neg, phi + TWO_PI, phi).astype(F)
        p = np.empty((d.shape[0], 2), dtype=F)
        p[:, 0] = phi / TWO_PI
        p[:, 1] = (cos_theta + F(1.0)) / F(2.0)
        flag = np.isfinite(d[:, 0]) & np.isfinite(d[:, 1]) & np.isfinite(d[:, 2])
        p[~flag] = 0
    return p
