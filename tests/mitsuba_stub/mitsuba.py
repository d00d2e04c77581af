"""TEST STAND-IN for the Mitsuba 3 names the plugin class touches (see drjit.py in this directory).
The "scene" is a box of side 1 whose every ray hits a grey diffuse surface at a pseudo-random point
and that glows uniformly: enough for the plugin's bounce loop, its record scatter and its calls
into PathGuidingCore to execute.  Nothing here models Mitsuba's behaviour."""
import numpy as np

import drjit as dr
from drjit import Bool, Float, Int32, UInt32, Vector2f, Vector3f  # noqa: F401

Spectrum = dr.Color3f
Color3f = dr.Color3f
_registered = {}


def register_integrator(name, factory):
    _registered[name] = factory


class SamplingIntegrator:
    def __init__(self, props):
        self.props = props


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


def luminance(c):
    return Float._of(c.v[:, 0] * np.float32(0.212671) + c.v[:, 1] * np.float32(0.715160) + c.v[:, 2] * np.float32(0.072169))


class Ray3f:
    def __init__(self, other=None, o=None, d=None):
        if other is not None:
            o, d = other.o, other.d
        self.o, self.d = Vector3f(o), Vector3f(d)

    def width(self):
        return self.o.v.shape[0]


class _BSDFSample:
    pass


class _Diffuse:
    """grey Lambertian surface, local frame == world frame"""

    def __init__(self, n):
        self.n = n

    def flags(self):
        return UInt32(np.full(self.n, BSDFFlags.Smooth))

    def _cos(self, wo):
        return np.maximum(wo.v[:, 2], 0.0)

    def eval_pdf(self, ctx, si, wo, active=True):
        c = self._cos(wo) / np.float32(np.pi)
        m = np.broadcast_to(np.asarray(dr.Arr._raw(active)), c.shape)
        return Spectrum(Float._of(np.where(m, 0.5 * c, 0))), Float._of(np.where(m, c, 0))

    def pdf(self, ctx, si, wo, active=True):
        return self.eval_pdf(ctx, si, wo, active)[1]

    def sample(self, ctx, si, u1, u2, active=True):
        u = np.broadcast_to(u2.v, (self.n, 2))
        r, phi = np.sqrt(u[:, 0]), 2 * np.pi * u[:, 1]
        wo = np.stack([r * np.cos(phi), r * np.sin(phi), np.sqrt(np.maximum(1 - u[:, 0], 0))], 1)
        bs = _BSDFSample()
        bs.wo = Vector3f._of(wo)
        bs.pdf = Float._of(wo[:, 2] / np.pi)
        bs.sampled_type = UInt32(np.full(self.n, BSDFFlags.Smooth))
        bs.eta = Float(np.ones(self.n))
        return bs, Spectrum(Float._of(np.full(self.n, 0.5)))


class SurfaceInteraction3f:
    def __init__(self, p, valid):
        self.p, self._valid = Vector3f._of(p), Bool._of(valid)

    @classmethod
    def zeros_(cls, shape=1):
        return cls(np.zeros((shape, 3), np.float32), np.zeros(shape, bool))

    def is_valid(self): return self._valid
    def bsdf(self): return _Diffuse(self.p.v.shape[0])
    def to_local(self, v): return Vector3f(v)
    def to_world(self, v): return Vector3f(v)
    def spawn_ray(self, d): return Ray3f(o=self.p, d=d)


class _Emitter:
    def eval(self, si):
        return Spectrum(Float._of(np.where(si.is_valid().v, 0.25, 0.0)))


class DirectionSample3f:
    def __init__(self, scene=None, si=None, ref=None):
        n = si.p.v.shape[0] if si is not None else 1
        self.emitter = _Emitter()
        self.d = Vector3f(np.zeros((n, 3)))
        self.pdf = Float(np.zeros(n))
        self.delta = Bool(np.zeros(n, bool))


class StubScene:
    def __init__(self, seed=1):
        self.rng = np.random.default_rng(seed)

    def ray_intersect(self, ray, ray_flags=None, coherent=None, active=True):
        n = ray.width()
        m = np.broadcast_to(np.asarray(dr.Arr._raw(active)), (n,))
        return SurfaceInteraction3f(self.rng.random((n, 3)).astype(np.float32), m & (self.rng.random(n) < 0.9))

    def pdf_emitter_direction(self, ref, ds, active=True):
        return Float(np.full(ds.d.v.shape[0], 1 / (4 * np.pi)))

    def sample_emitter_direction(self, si, u, test_visibility=True, active=True):
        n = si.p.v.shape[0]
        z = 1 - 2 * u.v[:, 0]
        r, phi = np.sqrt(np.maximum(1 - z * z, 0)), 2 * np.pi * u.v[:, 1]
        m = np.broadcast_to(np.asarray(dr.Arr._raw(active)), (n,))
        ds = DirectionSample3f(si=si)
        ds.d = Vector3f._of(np.stack([r * np.cos(phi), r * np.sin(phi), z], 1))
        ds.pdf = Float._of(np.where(m, 1 / (4 * np.pi), 0))
        return ds, Spectrum(Float._of(np.where(m, 0.25 * 4 * np.pi, 0)))


class StubSampler:
    def __init__(self, n, spp=1, seed=2):
        self.n, self.spp, self.rng = n, spp, np.random.default_rng(seed)

    def next_1d(self, active=True): return Float._of(self.rng.random(self.n))
    def next_2d(self, active=True): return Vector2f._of(self.rng.random((self.n, 2)))
    def sample_count(self): return self.spp
