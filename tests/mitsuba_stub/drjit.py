"""TEST STAND-IN for the handful of Dr.Jit calls the Mitsuba-facing plugin class makes
(practical_path_guiding_lab_b200/integrator.py).  numpy-backed eager arrays with the array
semantics the plugin relies on: elementwise operators with scalar / width-1 broadcasting, masked
assignment `x[mask] = v`, `.torch()` (here: the numpy array itself).  It exists so that the
plugin's control flow and its calls into PathGuidingCore run at all in an image without Mitsuba --
it says nothing about Mitsuba's own behaviour."""
import numpy as np


class Arr:
    K = 0                   # 0: flat (n,), else (n, K)
    DT = np.float32

    def __init__(self, *args):
        if len(args) > 1:                       # one value per component
            assert self.K == len(args)
            cols = [np.asarray(Arr._raw(c), self.DT).reshape(-1) for c in args]
            n = np.max([c.shape[0] for c in cols])
            self.v = np.stack([np.broadcast_to(c, (n,)) for c in cols], 1).astype(self.DT)
            return
        a = args[0] if args else 0
        v = np.asarray(Arr._raw(a))
        if self.K:
            if v.ndim == 0:
                v = np.full((1, self.K), v)
            elif v.ndim == 1:                   # one value per lane, repeated over the components
                v = np.repeat(v[:, None], self.K, 1)
            assert v.shape[1] == self.K
        else:
            v = v.reshape(-1)
        self.v = np.array(v, dtype=self.DT)

    # -- helpers
    @staticmethod
    def _raw(o):
        return o.v if isinstance(o, Arr) else o

    @classmethod
    def _of(cls, v):
        out = cls.__new__(cls)
        out.v = np.asarray(v, dtype=cls.DT)
        return out

    def _bin(self, o, f, rev=False, boolean=False):
        a, b = self.v, Arr._raw(o)
        cls = type(self)
        if isinstance(o, Arr):
            if a.ndim == 2 and b.ndim == 1:
                b = b[:, None]
            elif a.ndim == 1 and b.ndim == 2:
                a = a[:, None]
                cls = type(o)
            if cls.DT == np.uint32 and o.DT == np.float32 or cls.DT == np.int32 and o.DT == np.float32:
                cls = type(o)
        with np.errstate(all='ignore'):
            r = f(b, a) if rev else f(a, b)
        if boolean or r.dtype == np.bool_:
            return Bool._of(r)
        return cls._of(r)

    def __add__(self, o): return self._bin(o, np.add)
    def __radd__(self, o): return self._bin(o, np.add, True)
    def __sub__(self, o): return self._bin(o, np.subtract)
    def __rsub__(self, o): return self._bin(o, np.subtract, True)
    def __mul__(self, o): return self._bin(o, np.multiply)
    def __rmul__(self, o): return self._bin(o, np.multiply, True)
    def __truediv__(self, o): return self._bin(o, np.divide)
    def __rtruediv__(self, o): return self._bin(o, np.divide, True)
    def __pow__(self, o): return self._bin(o, np.power)
    def __neg__(self): return type(self)._of(-self.v)
    def __lt__(self, o): return self._bin(o, np.less, boolean=True)
    def __le__(self, o): return self._bin(o, np.less_equal, boolean=True)
    def __gt__(self, o): return self._bin(o, np.greater, boolean=True)
    def __ge__(self, o): return self._bin(o, np.greater_equal, boolean=True)
    def __and__(self, o): return self._bin(o, np.logical_and if self.DT == np.bool_ else np.bitwise_and)
    def __rand__(self, o): return self.__and__(o)
    def __or__(self, o): return self._bin(o, np.logical_or if self.DT == np.bool_ else np.bitwise_or)
    def __ror__(self, o): return self.__or__(o)
    def __invert__(self): return type(self)._of(~self.v)

    def __getitem__(self, key):
        if isinstance(key, Arr):                # masked read: the array itself (the mask matters on assignment)
            return self
        if isinstance(key, (int, np.integer)) and not self.K:
            return self.v[key].item()
        raise TypeError("stub: only x[mask] and flat x[i]")

    def __setitem__(self, key, value):
        assert isinstance(key, Arr) and key.DT == np.bool_, "stub: only masked assignment"
        m = key.v
        val = np.asarray(Arr._raw(value), self.DT)
        n = int(np.max([self.v.shape[0], m.shape[0], val.shape[0] if val.ndim else 1]))
        cur = np.broadcast_to(self.v, (n,) + self.v.shape[1:])
        if self.K:
            m = m[:, None]
            if val.ndim == 1:
                val = val[:, None]
        self.v = np.where(m, val, cur).astype(self.DT)

    def __bool__(self):
        assert self.v.size == 1, "stub: truth value of an array with more than one lane"
        return bool(self.v.reshape(-1)[0])

    def torch(self):
        return np.ascontiguousarray(self.v)

    @property
    def x(self): return Float._of(self.v[:, 0])
    @property
    def y(self): return Float._of(self.v[:, 1])
    @property
    def z(self): return Float._of(self.v[:, 2])

    def __repr__(self):
        return f"{type(self).__name__}({self.v!r})"


class Float(Arr):
    pass


class UInt32(Arr):
    DT = np.uint32


class Int32(Arr):
    DT = np.int32


class Bool(Arr):
    DT = np.bool_


class Vector2f(Arr):
    K = 2


class Vector3f(Arr):
    K = 3


class Color3f(Vector3f):
    pass


# ---- free functions (names are Dr.Jit's) ---------------------------------------------------------
def sqr(a): return a * a
def fma(a, b, c): return a * b + c
def isnan(a): return Bool._of(np.isnan(a.v))
def eq(a, b): return a._bin(b, np.equal, boolean=True)
def neq(a, b): return a._bin(b, np.not_equal, boolean=True)
def minimum(a, b): return a._bin(b, np.minimum) if isinstance(a, Arr) else b._bin(a, np.minimum)
def rcp(a): return 1.0 / a


def select(m, a, b):
    ref = a if isinstance(a, Arr) else (b if isinstance(b, Arr) else Float(0))
    if isinstance(a, Arr) and isinstance(b, Arr) and b.K and not a.K:
        ref = b
    mv = np.asarray(Arr._raw(m))
    av, bv = np.asarray(Arr._raw(a)), np.asarray(Arr._raw(b))
    if ref.K:
        mv = mv.reshape(-1, 1)
        if av.ndim == 1: av = av[:, None]
        if bv.ndim == 1: bv = bv[:, None]
    return type(ref)._of(np.where(mv, av, bv))


def any(m): return Bool._of(np.array([bool(np.any(m.v))]))       # noqa: A001
def width(a): return int(a.width()) if hasattr(a, "width") else int(a.v.shape[0])
def arange(cls, n): return cls(np.arange(n))
def max(a): return Float._of(a.v.max(axis=1))                    # noqa: A001  horizontal maximum of a vector / colour
def mean(a): return Float._of(np.array([a.v.mean()]))
def gather(cls, src, index): return cls._of(src.v[index.v])
def unravel(cls, flat): return cls._of(flat.v.reshape(-1, cls.K))
def eval(*a): return None                                        # noqa: A001
def sync_thread(): return None


def zeros(cls, shape=1):
    return cls.zeros_(shape) if hasattr(cls, "zeros_") else cls(np.zeros(shape))
