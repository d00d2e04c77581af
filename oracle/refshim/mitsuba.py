"""ORACLE SUPPORT (test infrastructure only -- never imported by the product path).

numpy-backed stand-in for the **Mitsuba 3.0-3.5** names the reference's SD-tree path touches (see
drjit.py in this directory for the purpose and the list of assumed primitive semantics).

Assumed here (Mitsuba 3 as documented upstream; not checkable in this image):
  * BoundingBox{2,3}f.contains(p) is inclusive on both ends, all axes; NaN is outside;
  * mi.luminance(c) = (c.r*0.212671 + c.g*0.715160) + c.b*0.072169 in fp32;
  * mi.Loop: see drjit.py (per-lane while loop);
  * Sampler.next_2d(active) = (next_1d(active), next_1d(active)), x first; a sampler call advances
    the stream of the lanes that execute it.  `ExplicitSampler` serves a given table
    u[lane, k] so that "the same uniform random numbers" can be fed to every implementation;
    `load_dict({'type': 'independent'})` gives a PCG32 stream (seeding recalled from upstream,
    only used to let the reference's print-only self-tests run).
The scene-facing classes (Ray3f, SurfaceInteraction3f, BSDF ...) are in scene_stub.py.
"""
import builtins

import numpy as _np

import drjit as dr
from drjit import (Bool, Color3f, Float, Int32, UInt32, Vector2f, Vector3f,  # noqa: F401
                   _Flat, _Nested, _bc, _leaves, _mask_stack)

Spectrum = Color3f
Point2f, Point3f, Normal3f = Vector2f, Vector3f, Vector3f
ScalarVector3f = Vector3f
_variant = [None]
_registered = {}


def set_variant(name, *more):
    _variant[0] = name


def variant():
    return _variant[0]


def register_integrator(name, factory):
    _registered[name] = factory


def luminance(c):
    return (c.c[0] * 0.212671 + c.c[1] * 0.715160) + c.c[2] * 0.072169


class _BBox:
    P = Vector3f

    def __init__(self, mn=None, mx=None):
        if mn is None:
            self.min, self.max = self.P(), self.P()
        else:
            self.min = self.P(mn)
            self.max = self.P(mx if mx is not None else mn)

    def contains(self, p, strict=False):
        r = None
        for k in range(self.P.K):
            t = (p.c[k] >= self.min.c[k]) & (p.c[k] <= self.max.c[k])
            r = t if r is None else (r & t)
        return r

    def __repr__(self):
        return f"{type(self).__name__}[min={self.min}, max={self.max}]"


class BoundingBox3f(_BBox):
    P = Vector3f
    DRJIT_STRUCT = {'min': Vector3f, 'max': Vector3f}


class BoundingBox2f(_BBox):
    P = Vector2f
    DRJIT_STRUCT = {'min': Vector2f, 'max': Vector2f}


class Loop:
    """Recorded loop == per-lane `while cond: body` (drjit.py, assumed semantics).
    `Loop.finished[name]` keeps the state tuple of the last loop of that name that ran to its end, so a
    test can read the node a descent stopped at without touching the reference's source."""
    finished = {}
    MAX_ITERATIONS = 512     # NOT reference behaviour: a lane that can never leave (NaN child energies in
    diverged = {}            # sampleQuadTree spin forever in the reference) is stopped here and reported

    def __init__(self, name='', state=None):
        self.name, self.state = name, state
        self.mask = None
        self.snap = None
        self.iterations = 0

    def set_max_iterations(self, n):
        pass

    def _blend(self):
        now = _leaves(self.state())
        assert len(now) == len(self.snap), f'loop "{self.name}": state changed shape'
        m = self.mask
        for leaf, old in zip(now, self.snap):
            n = builtins.max(leaf.d.shape[0], old.shape[0], m.shape[0])
            leaf.d = _np.where(_bc(m, n), _bc(leaf.d, n), _bc(old, n)).astype(leaf.DT)

    def __call__(self, cond):
        if self.mask is not None:
            self._blend()
            _mask_stack.pop()
        c = cond.d if isinstance(cond, _Flat) else _np.asarray(bool(cond)).reshape(-1)
        if self.mask is not None:
            n = builtins.max(c.shape[0], self.mask.shape[0])
            c = _bc(c, n) & _bc(self.mask, n)
        if c.any() and self.iterations >= Loop.MAX_ITERATIONS:
            Loop.diverged[self.name] = c.copy()
            c = _np.zeros_like(c)
        if not c.any():
            self.mask = self.snap = None
            Loop.finished[self.name] = self.state() if self.state is not None else None
            return False
        self.mask = c.copy()
        self.snap = [leaf.d.copy() for leaf in _leaves(self.state())]
        _mask_stack.append(self.mask)
        self.iterations += 1
        return True


# ----------------------------------------------------------------------------------- samplers
class Sampler:
    def _loop_state(self):
        return []

    def schedule_state(self):
        pass


class ExplicitSampler(Sampler):
    """Serves u[lane, cursor[lane]] and advances the cursor of the lanes that execute the call."""

    def __init__(self, u, spp=1):
        self.u = _np.ascontiguousarray(u, dtype=_np.float32)
        self.n = self.u.shape[0]
        self.cursor = UInt32._of(_np.zeros(self.n, _np.uint32))
        self._spp = spp

    def _loop_state(self):
        return [self.cursor]

    def seed(self, seed, wavefront_size=None):
        self.cursor = UInt32._of(_np.zeros(self.n, _np.uint32))

    def sample_count(self):
        return self._spp

    def next_1d(self, active=True):
        c = _np.minimum(self.cursor.d.astype(_np.int64), self.u.shape[1] - 1)
        v = self.u[_np.arange(self.n), c]
        a = active.d if isinstance(active, _Flat) else _np.asarray(bool(active)).reshape(-1)
        self.cursor = UInt32._of(self.cursor.d + _bc(a, self.n).astype(_np.uint32))
        return Float._of(v)

    def next_2d(self, active=True):
        x = self.next_1d(active)
        y = self.next_1d(active)
        return Vector2f._of([x, y])


class PCG32Sampler(Sampler):
    """`independent` sampler: PCG32 (published algorithm); per-lane seeding via TEA as recalled from
    Mitsuba 3's PCG32Sampler::seed.  Used only by the reference's print-only __main__ blocks."""
    MULT = _np.uint64(0x5851F42D4C957F2D)

    class _State(_Flat):
        DT = _np.uint64

    def __init__(self):
        self.state = PCG32Sampler._State._of(_np.zeros(0, _np.uint64))
        self.inc = None
        self.n = 0

    @staticmethod
    def _tea(v0, v1, rounds=4):
        v0, v1 = v0.astype(_np.uint32), v1.astype(_np.uint32)
        s = 0
        with _np.errstate(over='ignore'):
            for _ in range(rounds):
                s = (s + 0x9E3779B9) & 0xFFFFFFFF
                s32 = _np.uint32(s)
                v0 = v0 + ((((v1 << _np.uint32(4)) + _np.uint32(0xA341316C)) ^ (v1 + s32)) ^ ((v1 >> _np.uint32(5)) + _np.uint32(0xC8013EA4)))
                v1 = v1 + ((((v0 << _np.uint32(4)) + _np.uint32(0xAD90777D)) ^ (v0 + s32)) ^ ((v0 >> _np.uint32(5)) + _np.uint32(0x7E95761E)))
        return v0, v1

    def _next_u32(self, a):
        old = self.state.d
        with _np.errstate(over='ignore'):
            new = old * self.MULT + self.inc
        self.state.d = _np.where(a, new, old)
        xs = (((old >> _np.uint64(18)) ^ old) >> _np.uint64(27)).astype(_np.uint32)
        rot = (old >> _np.uint64(59)).astype(_np.uint32)
        with _np.errstate(over='ignore'):
            return (xs >> rot) | (xs << ((~rot + _np.uint32(1)) & _np.uint32(31)))

    def seed(self, seed, wavefront_size=1):
        self.n = int(wavefront_size)
        idx = _np.arange(self.n, dtype=_np.uint32)
        v0, v1 = self._tea(_np.full(self.n, seed, _np.uint32), idx)
        self.state.d = _np.zeros(self.n, _np.uint64)
        self.inc = (v1.astype(_np.uint64) << _np.uint64(1)) | _np.uint64(1)
        t = _np.ones(self.n, bool)
        self._next_u32(t)
        with _np.errstate(over='ignore'):
            self.state.d = self.state.d + v0.astype(_np.uint64)
        self._next_u32(t)

    def _loop_state(self):
        return [self.state]

    def sample_count(self):
        return 1

    def next_1d(self, active=True):
        a = active.d if isinstance(active, _Flat) else _np.asarray(bool(active)).reshape(-1)
        u = self._next_u32(_bc(a, self.n))
        f = ((u >> _np.uint32(9)) | _np.uint32(0x3F800000)).view(_np.float32) - _np.float32(1.0)
        return Float._of(f)

    def next_2d(self, active=True):
        x = self.next_1d(active)
        y = self.next_1d(active)
        return Vector2f._of([x, y])


def load_dict(d):
    if d.get('type') == 'independent':
        return PCG32Sampler()
    raise NotImplementedError(f"shim: load_dict({d!r})")


# -------------------------------------------------------------------- integrator-side plumbing
class Properties(dict):
    pass


class SamplingIntegrator:
    def __init__(self, props=None):
        self.props = props


class Scene:
    pass


class Medium:
    pass


class BSDFFlags:
    Delta = 0x0F0
    Smooth = 0x00F


class RayFlags:
    All = 0xFFFF


class BSDFContext:
    def __init__(self):
        self.component = 0xFFFFFFFF


def has_flag(flags, f):
    return dr.neq(flags & int(f), 0)


try:                                                                    # scene-facing classes
    from scene_stub import (DirectionSample3f, Ray3f, RayDifferential3f,    # noqa: E402,F401
                            StubScene, SurfaceInteraction3f)
except ImportError:                                                     # pragma: no cover
    pass
