"""ORACLE SUPPORT (test infrastructure only -- never imported by the product path).

A numpy-backed, eager stand-in for the part of **Dr.Jit 0.4.x** that the reference's SD-tree path
touches, written so that /root/reference/src/common.py, src/quadtree.py, src/kdtree.py and
src/path_guiding_integrator.py import and run UNMODIFIED (see oracle/refshim/__init__.py for the
loader).  With it the reference's own control flow, tie rules, operation order and node numbering
execute here; what stays ASSUMED is only the semantics of each primitive, listed below.  This is
how oracle/sdtree_oracle.py is pinned: tests/test_reference_on_shim.py demands identical arrays
from "reference source on this shim" and from the oracle on every parity / fuzz case, and
tests/golden/make_reference_golden.py freezes outputs of that leg for the GPU suite.

Primitive semantics assumed (Dr.Jit 0.4.x as documented upstream; not checkable in this image):
  * arrays are fp32 / u32 / i32 / bool; a Python scalar combined with an array is first converted
    to the ARRAY's element type (so `f32_array > 16970.56` compares in fp32);
  * width-1 arrays broadcast; `x[mask] = v` == `x = select(mask, v, x)`; `x[mask] op= v` changes the
    masked lanes only; `x[mask]` as an rvalue is the whole array;
  * dr.gather(T, src, idx, active): 0 on masked lanes; nested source/target gather per component;
  * dr.scatter / dr.scatter_reduce(Add): masked lanes do nothing; Add is sequential fp32 addition in
    lane order here (the real one is an atomic in unspecified order);
  * dr.compress: ascending indices of the set lanes;
  * mi.Loop (recorded loop) == per-lane `while cond: body`: a lane whose condition is false at the
    top of an iteration keeps ALL its loop state (sampler included) and performs no side effect;
    every operation of the body runs for a lane that was active at the top, also after the lane
    cleared its own flag in mid-body;
  * `array / python_scalar`: IEEE fp32 division by default, or multiplication by the fp32-rounded
    reciprocal when SCALAR_DIV_RECIPROCAL is set (open uncertainty);
  * sqrt / division IEEE; sincos / atan2 = the CEPHES kernels restated in oracle/drjit_math.py;
    no FMA contraction except the explicit dr.fma.
"""
import builtins as _b

import numpy as _np

from oracle import drjit_math as _dm

F32, U32, I32, B8 = _np.float32, _np.uint32, _np.int32, _np.bool_

SCALAR_DIV_RECIPROCAL = False

pi = float(_np.pi)
two_pi = float(2.0 * _np.pi)
inv_four_pi = float(1.0 / (4.0 * _np.pi))
inf = float('inf')

_mask_stack = []          # pushed by mitsuba.Loop: side effects of exited lanes are suppressed


def _top_mask(n):
    if _mask_stack and _mask_stack[-1].shape[0] == n:
        return _mask_stack[-1]
    return None


class ArrayBase:
    """Common base, as in Dr.Jit (src/common.py annotates with dr.ArrayBase)."""
    K = 0                   # 0 = flat, otherwise number of components
    DT = F32
    IsFloat = True


# ------------------------------------------------------------------------------------------- flat
class _Flat(ArrayBase):
    def __init__(self, value=None):
        if value is None:
            self.d = _np.zeros(0, self.DT)
        elif isinstance(value, _Flat):
            self.d = value.d.astype(self.DT, copy=True)
        else:
            with _np.errstate(all='ignore'):
                self.d = _np.array(value, dtype=self.DT).reshape(-1).copy()

    @classmethod
    def _of(cls, d):
        o = cls.__new__(cls)
        o.d = _np.ascontiguousarray(d, dtype=cls.DT)
        return o

    # -- conversions
    def numpy(self):
        return self.d.copy()

    def __len__(self):
        return self.d.shape[0]

    def __repr__(self):
        return f"{type(self).__name__}({self.d.tolist()!r})"

    def __bool__(self):
        assert self.d.shape[0] == 1, "truth value of an array with more than one lane"
        return bool(self.d[0])

    # -- operators
    def _coerce(self, o):
        """other operand -> (numpy array, result class)"""
        if isinstance(o, _Flat):
            if type(o) is type(self):
                return o.d, type(self)
            # mixed element types: float wins over int, int over bool
            rank = {B8: 0, U32: 1, I32: 1, F32: 2}
            cls = type(self) if rank[self.DT] >= rank[o.DT] else type(o)
            return o.d, cls
        if isinstance(o, _Nested):
            return NotImplemented, None
        if isinstance(o, (bool, _np.bool_)) and self.DT is not B8:
            o = int(o)
        with _np.errstate(all='ignore'):
            return _np.asarray(o).astype(self.DT).reshape(-1), type(self)     # python scalar -> element type

    def _bin(self, o, f, rev=False, out_bool=False):
        b, cls = self._coerce(o)
        if b is NotImplemented:
            return NotImplemented
        a = self.d
        if cls.DT is not self.DT:
            a = a.astype(cls.DT)
        if b.dtype != cls.DT:
            b = b.astype(cls.DT)
        with _np.errstate(all='ignore'):
            r = f(b, a) if rev else f(a, b)
        return Bool._of(r) if out_bool else cls._of(r)

    def __add__(self, o): return self._bin(o, _np.add)
    def __radd__(self, o): return self._bin(o, _np.add, True)
    def __sub__(self, o): return self._bin(o, _np.subtract)
    def __rsub__(self, o): return self._bin(o, _np.subtract, True)
    def __mul__(self, o): return self._bin(o, _np.multiply)
    def __rmul__(self, o): return self._bin(o, _np.multiply, True)

    def __truediv__(self, o):
        assert self.DT is F32, 'use // for integer arrays'
        if SCALAR_DIV_RECIPROCAL and isinstance(o, (int, float)) and not isinstance(o, bool):
            return self * (1.0 / o)
        return self._bin(o, _np.divide)

    def __rtruediv__(self, o): return self._bin(o, _np.divide, True)

    def __floordiv__(self, o):
        assert self.DT is not F32
        return self._bin(o, _np.floor_divide)

    def __mod__(self, o):
        assert self.DT is not F32
        return self._bin(o, _np.remainder)

    def __pow__(self, o):
        if isinstance(o, int) and o >= 0:           # integer powers by repeated multiplication
            r = type(self)._of(_np.ones_like(self.d))
            for _ in range(o):
                r = r * self
            return r
        return self._bin(o, _np.power)

    def __neg__(self): return type(self)._of(-self.d)
    def __lt__(self, o): return self._bin(o, _np.less, out_bool=True)
    def __le__(self, o): return self._bin(o, _np.less_equal, out_bool=True)
    def __gt__(self, o): return self._bin(o, _np.greater, out_bool=True)
    def __ge__(self, o): return self._bin(o, _np.greater_equal, out_bool=True)
    __hash__ = None

    def _logic(self, o, fb, fi):
        return self._bin(o, fb if self.DT is B8 else fi)

    def __and__(self, o): return self._logic(o, _np.logical_and, _np.bitwise_and)
    def __rand__(self, o): return self.__and__(o)
    def __or__(self, o): return self._logic(o, _np.logical_or, _np.bitwise_or)
    def __ror__(self, o): return self.__or__(o)
    def __xor__(self, o): return self._logic(o, _np.logical_xor, _np.bitwise_xor)

    def __invert__(self):
        return type(self)._of(_np.logical_not(self.d) if self.DT is B8 else ~self.d)

    # in-place forms keep the object (loop state, attributes of structs)
    def _assign(self, r):
        if r is NotImplemented:
            raise TypeError('unsupported in-place operand')
        assert type(r) is type(self), (type(r), type(self))
        self.d = r.d
        return self

    def __iadd__(self, o): return self._assign(self + o)
    def __isub__(self, o): return self._assign(self - o)
    def __imul__(self, o): return self._assign(self * o)
    def __itruediv__(self, o): return self._assign(self / o)
    def __iand__(self, o): return self._assign(self & o)
    def __ior__(self, o): return self._assign(self | o)

    # -- indexing
    def __getitem__(self, key):
        if isinstance(key, Bool):
            return type(self)._of(self.d.copy())          # rvalue of a masked expression: the array
        if isinstance(key, (int, _np.integer)):
            return self.d[key].item()
        raise TypeError('shim: only x[mask] and x[int]')

    def __setitem__(self, key, value):
        if isinstance(key, (int, _np.integer)):
            self.d[key] = value
            return
        assert isinstance(key, Bool), 'shim: only masked assignment'
        v = value.d if isinstance(value, _Flat) else _np.asarray(value)
        with _np.errstate(all='ignore'):
            v = v.astype(self.DT).reshape(-1)
        n = _b.max(self.d.shape[0], key.d.shape[0], v.shape[0])
        self.d = _np.where(_bc(key.d, n), _bc(v, n), _bc(self.d, n)).astype(self.DT)


def _bc(a, n):
    if a.shape[0] == n:
        return a
    if n == 0:
        return a[:0]
    assert a.shape[0] == 1, f'width mismatch: {a.shape[0]} vs {n}'
    return _np.broadcast_to(a, (n,) + a.shape[1:])


class Float(_Flat):
    DT = F32


class UInt32(_Flat):
    DT = U32
    IsFloat = False


class Int32(_Flat):
    DT = I32
    IsFloat = False


class Bool(_Flat):
    DT = B8
    IsFloat = False


# ----------------------------------------------------------------------------------------- nested
class _Nested(ArrayBase):
    """Static array of K flat arrays (structure of arrays, like Dr.Jit's Array3f<Float>)."""
    K = 3
    LEAF = Float
    NAMES = 'xyz'

    def __init__(self, *args):
        L, K = self.LEAF, self.K
        if len(args) == 0:
            self.c = [L() for _ in range(K)]
        elif len(args) == K and K > 1:
            self.c = [L(a) for a in args]
        else:
            assert len(args) == 1, args
            a = args[0]
            if isinstance(a, _Nested):
                assert a.K == K
                self.c = [L(x) for x in a.c]
            elif isinstance(a, _Flat):
                self.c = [L(a) for _ in range(K)]
            else:
                v = _np.asarray(a)
                if v.ndim == 0:
                    self.c = [L(v) for _ in range(K)]
                elif v.ndim == 1:
                    assert v.shape[0] == K, 'one value per component expected'
                    self.c = [L(v[i]) for i in range(K)]
                else:
                    assert v.ndim == 2 and v.shape[1] == K, v.shape
                    self.c = [L(v[:, i]) for i in range(K)]

    @classmethod
    def _of(cls, comps):
        o = cls.__new__(cls)
        o.c = list(comps)
        return o

    def numpy(self):
        n = width(self)
        return _np.stack([_bc(x.d, n) for x in self.c], axis=1)

    def __repr__(self):
        return f"{type(self).__name__}({self.numpy().tolist()!r})"

    def __len__(self):
        return self.K

    def _comp(i):                                   # noqa: N805
        def get(self): return self.c[i]

        def set_(self, v):
            self.c[i] = v if isinstance(v, self.LEAF) else self.LEAF(v)
        return property(get, set_)

    x, y, z = _comp(0), _comp(1), _comp(2)

    def _other(self, o, i):
        return o.c[i] if isinstance(o, _Nested) else o

    def _map(self, o, name, out_mask=False):
        if isinstance(o, _Nested):
            assert o.K == self.K
        comps = [getattr(self.c[i], name)(self._other(o, i)) for i in range(self.K)]
        cls = type(self)
        if comps and isinstance(comps[0], Bool) and self.LEAF is not Bool:
            cls = _mask_type(self.K)
        elif isinstance(o, _Nested) and self.LEAF is not Float and o.LEAF is Float:
            cls = type(o)
        return cls._of(comps)

    def __add__(self, o): return self._map(o, '__add__')
    def __radd__(self, o): return self._map(o, '__radd__')
    def __sub__(self, o): return self._map(o, '__sub__')
    def __rsub__(self, o): return self._map(o, '__rsub__')
    def __mul__(self, o): return self._map(o, '__mul__')
    def __rmul__(self, o): return self._map(o, '__rmul__')
    def __truediv__(self, o): return self._map(o, '__truediv__')
    def __rtruediv__(self, o): return self._map(o, '__rtruediv__')
    def __pow__(self, o): return self._map(o, '__pow__')
    def __lt__(self, o): return self._map(o, '__lt__')
    def __le__(self, o): return self._map(o, '__le__')
    def __gt__(self, o): return self._map(o, '__gt__')
    def __ge__(self, o): return self._map(o, '__ge__')
    def __and__(self, o): return self._map(o, '__and__')
    def __or__(self, o): return self._map(o, '__or__')
    def __neg__(self): return type(self)._of([-x for x in self.c])
    def __invert__(self): return type(self)._of([~x for x in self.c])
    __hash__ = None

    def _assign(self, r):
        for i in range(self.K):
            self.c[i].d = r.c[i].d
        return self

    def __iadd__(self, o): return self._assign(self + o)
    def __isub__(self, o): return self._assign(self - o)
    def __imul__(self, o): return self._assign(self * o)
    def __itruediv__(self, o): return self._assign(self / o)

    def __getitem__(self, key):
        if isinstance(key, (Bool, _Nested)):
            return type(self)._of([type(x)._of(x.d.copy()) for x in self.c])
        if isinstance(key, (int, _np.integer)):
            return self.c[key]
        raise TypeError('shim: only v[mask] and v[int]')

    def __setitem__(self, key, value):
        if isinstance(key, (int, _np.integer)):
            self.c[key] = value if isinstance(value, self.LEAF) else self.LEAF(value)
            return
        for i in range(self.K):
            k = key.c[i] if isinstance(key, _Nested) else key
            self.c[i][k] = self._other(value, i)


class Vector2f(_Nested):
    K = 2


class Vector3f(_Nested):
    K = 3


class Color3f(_Nested):
    K = 3


class Vector2b(_Nested):
    K = 2
    LEAF = Bool


class Vector3b(_Nested):
    K = 3
    LEAF = Bool


def _mask_type(k):
    return {2: Vector2b, 3: Vector3b}[k]


# -------------------------------------------------------------------------------------- structs
def _struct_fields(obj):
    s = getattr(type(obj), 'DRJIT_STRUCT', None)
    return None if s is None else list(s.keys())


def _leaves(obj, out=None):
    """flat arrays reachable from a loop-state / struct object, in a fixed order"""
    if out is None:
        out = []
    if isinstance(obj, _Flat):
        out.append(obj)
    elif isinstance(obj, _Nested):
        out.extend(obj.c)
    elif isinstance(obj, (tuple, list)):
        for o in obj:
            _leaves(o, out)
    elif hasattr(obj, '_loop_state'):
        _leaves(obj._loop_state(), out)
    elif _struct_fields(obj) is not None:
        for k in _struct_fields(obj):
            _leaves(getattr(obj, k), out)
    return out


# ------------------------------------------------------------------------------------ functions
def width(a):
    if isinstance(a, _Flat):
        return int(a.d.shape[0])
    if isinstance(a, (int, float, bool)):
        return 1
    ls = _leaves(a)
    return _b.max((int(x.d.shape[0]) for x in ls), default=0)


def zeros(cls, shape=1):
    if isinstance(cls, type) and issubclass(cls, _Flat):
        return cls._of(_np.zeros(shape, cls.DT))
    if isinstance(cls, type) and issubclass(cls, _Nested):
        return cls._of([zeros(cls.LEAF, shape) for _ in range(cls.K)])
    s = getattr(cls, 'DRJIT_STRUCT', None)
    assert s is not None, f'dr.zeros: unsupported type {cls}'
    obj = cls()
    for k, t in s.items():
        setattr(obj, k, zeros(t, shape))
    if hasattr(obj, 'zero_'):
        obj.zero_(shape)
    return obj


def full(cls, value, shape=1):
    if issubclass(cls, _Flat):
        return cls._of(_np.full(shape, value, cls.DT))
    return cls._of([full(cls.LEAF, value, shape) for _ in range(cls.K)])


def ones(cls, shape=1):
    return full(cls, 1, shape)


def arange(cls, start, stop=None, step=1):
    if stop is None:
        start, stop = 0, start
    return cls._of(_np.arange(start, stop, step).astype(cls.DT))


def _idx(index):
    if isinstance(index, _Flat):
        return index.d.astype(_np.int64)
    return _np.asarray(index, dtype=_np.int64).reshape(-1)


def _act(active, n):
    if isinstance(active, _Flat):
        a = active.d
    else:
        a = _np.asarray(bool(active)).reshape(-1)
    return a, (n if a.shape[0] == 1 else _b.max(n, a.shape[0]))


def gather(dtype, source, index, active=True):
    if isinstance(source, _Nested):
        assert issubclass(dtype, _Nested) and dtype.K == source.K
        return dtype._of([gather(dtype.LEAF, c, index, active) for c in source.c])
    if not isinstance(source, _Flat) and _struct_fields(source) is not None:
        obj = dtype()
        for k, t in dtype.DRJIT_STRUCT.items():
            setattr(obj, k, gather(t, getattr(source, k), index, active))
        return obj
    assert isinstance(source, _Flat), type(source)
    if issubclass(dtype, _Nested):                         # flat AoS source -> nested
        i = _idx(index)
        return dtype._of([gather(dtype.LEAF, source, UInt32._of(i * dtype.K + k), active) for k in range(dtype.K)])
    i = _idx(index)
    a, n = _act(active, i.shape[0])
    i, a = _bc(i, n), _bc(a, n)
    tm = _top_mask(n)
    if tm is not None:
        a = a & tm
    out = _np.zeros(n, dtype.DT)
    if source.d.shape[0] == 1:                  # a width-1 source is a scalar: every index reads it
        i = _np.zeros_like(i)
    if a.all():
        out = source.d[i].astype(dtype.DT)
    elif a.any():
        out[a] = source.d[i[a]]
    return dtype._of(out)


def _scatter_flat(target, value, index, active, reduce_add):
    i = _idx(index)
    v = value.d if isinstance(value, _Flat) else _np.asarray(value).reshape(-1)
    with _np.errstate(all='ignore'):
        v = v.astype(target.DT)
    a, n = _act(active, i.shape[0] if v.shape[0] == 1 else _b.max(i.shape[0], v.shape[0]))
    i, v, a = _bc(i, n), _bc(v, n), _bc(a, n)
    tm = _top_mask(n)
    if tm is not None:
        a = a & tm
    if not a.all():
        i, v = i[a], v[a]
    if i.shape[0] == 0:
        return
    if reduce_add:
        with _np.errstate(all='ignore'):
            _np.add.at(target.d, i, v)               # sequential fp32 adds in lane order
    else:
        target.d[i] = v


def scatter(target, value, index, active=True):
    if isinstance(target, _Nested):
        for k in range(target.K):
            scatter(target.c[k], value.c[k] if isinstance(value, _Nested) else value, index, active)
        return
    _scatter_flat(target, value, index, active, False)


class ReduceOp:
    Add = 'add'


def scatter_reduce(op, target, value, index, active=True):
    assert op == ReduceOp.Add
    if isinstance(target, _Nested):
        for k in range(target.K):
            scatter_reduce(op, target.c[k], value.c[k] if isinstance(value, _Nested) else value, index, active)
        return
    _scatter_flat(target, value, index, active, True)


def compress(mask):
    return UInt32._of(_np.flatnonzero(mask.d).astype(U32))


def select(m, a, b):
    ref = a if isinstance(a, ArrayBase) else (b if isinstance(b, ArrayBase) else None)
    if isinstance(a, _Flat) and isinstance(b, _Nested):
        ref = b
    if isinstance(ref, _Nested):
        return type(ref)._of([select(m.c[k] if isinstance(m, _Nested) else m,
                                     a.c[k] if isinstance(a, _Nested) else a,
                                     b.c[k] if isinstance(b, _Nested) else b) for k in range(ref.K)])
    if ref is None:
        ref = Float(0) if isinstance(a, float) or isinstance(b, float) else Int32(0)
    cls = type(ref)
    if isinstance(a, _Flat) and isinstance(b, _Flat) and type(a) is not type(b):
        cls = Float if Float in (type(a), type(b)) else cls
    mv = m.d if isinstance(m, _Flat) else _np.asarray(bool(m)).reshape(-1)
    with _np.errstate(all='ignore'):
        av = (a.d if isinstance(a, _Flat) else _np.asarray(a).reshape(-1)).astype(cls.DT)
        bv = (b.d if isinstance(b, _Flat) else _np.asarray(b).reshape(-1)).astype(cls.DT)
    n = _b.max(mv.shape[0], av.shape[0], bv.shape[0])
    return cls._of(_np.where(_bc(mv, n), _bc(av, n), _bc(bv, n)))


def eq(a, b):
    if isinstance(a, _Nested):
        return a._map(b, '_eq')
    return a._bin(b, _np.equal, out_bool=True)


def neq(a, b):
    if isinstance(a, _Nested):
        return a._map(b, '_ne')
    return a._bin(b, _np.not_equal, out_bool=True)


_Flat._eq = lambda self, o: self._bin(o, _np.equal, out_bool=True)          # noqa: E731
_Flat._ne = lambda self, o: self._bin(o, _np.not_equal, out_bool=True)      # noqa: E731


def _unary(f, out_bool=False):
    def g(a):
        if isinstance(a, _Nested):
            comps = [g(x) for x in a.c]
            return (_mask_type(a.K) if out_bool else type(a))._of(comps)
        if not isinstance(a, _Flat):
            a = Float(a)
        with _np.errstate(all='ignore'):
            r = f(a.d)
        return Bool._of(r) if out_bool else type(a)._of(r)
    return g


isnan = _unary(_np.isnan, True)
isfinite = _unary(_np.isfinite, True)
isinf = _unary(_np.isinf, True)
sqrt = _unary(_np.sqrt)
abs = _unary(_np.abs)                                   # noqa: A001


def sqr(a):
    return a * a


def rcp(a):
    return 1.0 / a


def fma(a, b, c):
    """single rounding of a*b + c (the product of two fp32 values is exact in float64)"""
    if isinstance(a, _Nested) or isinstance(b, _Nested) or isinstance(c, _Nested):
        ref = next(v for v in (a, b, c) if isinstance(v, _Nested))
        pick = lambda v, k: v.c[k] if isinstance(v, _Nested) else v            # noqa: E731
        return type(ref)._of([fma(pick(a, k), pick(b, k), pick(c, k)) for k in range(ref.K)])
    a, b, c = (v if isinstance(v, _Flat) else Float(v) for v in (a, b, c))
    n = _b.max(width(a), width(b), width(c))
    with _np.errstate(all='ignore'):
        r = _bc(a.d, n).astype(_np.float64) * _bc(b.d, n).astype(_np.float64) + _bc(c.d, n).astype(_np.float64)
    return Float._of(r.astype(F32))


def minimum(a, b):
    if isinstance(a, _Nested):
        return a._map(b, '_min')
    if not isinstance(a, _Flat):
        a, b = b, a
    return a._bin(b, _np.minimum)


def maximum(a, b):
    if isinstance(a, _Nested):
        return a._map(b, '_max')
    if not isinstance(a, _Flat):
        a, b = b, a
    return a._bin(b, _np.maximum)


_Flat._min = lambda self, o: self._bin(o, _np.minimum)     # noqa: E731
_Flat._max = lambda self, o: self._bin(o, _np.maximum)     # noqa: E731


def clip(a, lo, hi):
    return minimum(maximum(a, lo), hi)


clamp = clip


def sincos(a):
    s, c = _dm.sincos(a.d)
    return Float._of(s), Float._of(c)


def sin(a): return sincos(a)[0]
def cos(a): return sincos(a)[1]


def atan2(y, x):
    n = _b.max(width(y), width(x))
    return Float._of(_dm.atan2(_bc(y.d, n), _bc(x.d, n)))


def any(m):                                             # noqa: A001
    if isinstance(m, _Nested):
        r = m.c[0]
        for x in m.c[1:]:
            r = r | x
        return r
    if isinstance(m, (bool, _np.bool_)):
        return bool(m)
    return bool(m.d.any())


def all(m):                                             # noqa: A001
    if isinstance(m, _Nested):
        r = m.c[0]
        for x in m.c[1:]:
            r = r & x
        return r
    if isinstance(m, (bool, _np.bool_)):
        return bool(m)
    return bool(m.d.all())


def max(a):                                             # noqa: A001
    """horizontal maximum: over the components of a nested array, over the lanes of a flat one"""
    if isinstance(a, _Nested):
        r = a.c[0]
        for x in a.c[1:]:
            r = maximum(r, x)
        return r
    return type(a)._of(a.d.max(keepdims=True))


def sum(a):                                             # noqa: A001
    if isinstance(a, _Nested):
        r = a.c[0]
        for x in a.c[1:]:
            r = r + x
        return r
    return type(a)._of(a.d.sum(keepdims=True, dtype=a.DT))


def mean(a):
    if isinstance(a, _Nested):
        return type(a)._of([mean(x) for x in a.c])
    with _np.errstate(all='ignore'):
        return Float._of(_np.asarray([a.d.astype(_np.float64).mean()]).astype(F32))


def ravel(a):
    if isinstance(a, _Flat):
        return a
    return a.LEAF._of(a.numpy().reshape(-1))


def unravel(cls, flat):
    v = flat.d.reshape(-1, cls.K)
    return cls._of([cls.LEAF._of(v[:, k].copy()) for k in range(cls.K)])


def repeat(a, count):
    if isinstance(a, _Nested):
        return type(a)._of([repeat(x, count) for x in a.c])
    return type(a)._of(_np.repeat(a.d, count))


def tile(a, count):
    if isinstance(a, _Nested):
        return type(a)._of([tile(x, count) for x in a.c])
    return type(a)._of(_np.tile(a.d, count))


def eval(*a, **k): return None                          # noqa: A001
def schedule(*a, **k): return None
def sync_thread(): return None


printf_calls = []


def printf_async(fmt, *args, active=True):
    """the reference only prints from its validators, on the lanes that FAIL: keep the count"""
    a = active.d if isinstance(active, _Flat) else _np.asarray(bool(active)).reshape(-1)
    if a.any():
        printf_calls.append((fmt, int(a.sum())))
