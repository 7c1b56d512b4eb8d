"""oracle/ti_shim.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A strict-IEEE-fp32 INTERPRETER for the subset of `taichi` / `taichi_glsl` that the reference's
differender/volume_raycaster.py uses, so that the reference's OWN SOURCE (read from /root/reference at run time, never copied)
can be executed in this container, where the real Taichi JIT is not installable (no network).

What it is: two stand-in modules (`taichi`, `taichi_glsl`) plus replacements for the builtins that Taichi re-interprets inside
kernels (`float`, `int`, `max`, `min`, `pow`).  With them the reference's source runs UNMODIFIED (its `ti.random()` jitter is served
from a supplied tensor, set_random_source) -- the `VolumeRaycaster` class (:56-389), and the `Raycaster` / `RaycastFunction` wrappers on
top of it -- as plain Python:
every scalar is an fp32 value (`F`), every operator rounds once to fp32 in SOURCE ORDER (no contraction, no fast-math), struct-for
loops iterate the field indices, fields are numpy arrays, and `kernel.grad()` is a reverse sweep over a tape of the scalar operations
the forward call recorded (adjoints accumulated in float64; the select-style adjoints of max/min and the unconditional
`adjoint x partial` products of Taichi's autodiff are kept, so 0 x inf = NaN poisoning -- SURVEY 7.3 H4 -- shows up here as it would
there).

What it pins: the COMPOSITION of the algorithm -- order of operations, control flow, indices, clamps, the ERT rule, the jitter and
sample-position formulas, which values are differentiated -- against the reference's source rather than against a reading of it.
The oracle's `source_order` build (oracle/cpu_ref.c with every contraction switched off) must agree with it BIT FOR BIT on the
image, the sample counts and the early-termination counts, and to float64-accumulation accuracy on both gradients
(tests/test_shim_pin.py, tests/golden/shim_*.npz, oracle/taichi_probe.py --shim).

What it does NOT pin: the real Taichi compiler's rounding choices (FMA contraction, fast-math division / pow) -- that is what
tools/rounding_envelope.py bounds and what oracle/taichi_probe.py checks on the first box that has Taichi -- and the PRIMITIVES,
which are restated here from the published definitions of the two packages exactly as in oracle/cpu_ref.c:
    taichi_glsl:  mix(x, y, a) = x*(1-a) + y*a;  clamp(x, lo, hi) = min(hi, max(lo, x));  reflect(I, N) = I - 2*N.dot(I)*N;
                  cross; vec3 / vec4 constructors (a single scalar broadcasts)
    taichi:       Vector.dot = sum of products left to right;  norm = sqrt(norm_sqr + eps);  normalized = (1/(norm + eps)) * v;
                  max(a, b) = a > b ? a : b and min(a, b) = a < b ? a : b (so max(NaN, 0) = 0);  float -> int casts truncate;
                  a Python / numpy constant stays a double until it meets an fp32 value (ti.tan of a Python float is math.tan);
                  adjoints: max/min route the whole adjoint to the selected operand (ties: the right-hand one), floor has none,
                  d pow(a, b) / da = b * pow(a, b - 1), d sqrt(a) = 0.5 / sqrt(a), d (a / b) = (1 / b, -a / b^2).
Two reference quirks are resolved the way SURVEY 7.3 defines them: a NEGATIVE field index (tape[i, j, -1], H1) reads zeros, and
an index past the end raises (H2: parity runs choose max_samples >= the longest ray).
"""
import ctypes
import ctypes.util
import math
import operator
import sys
import types

import numpy as np

f32 = np.float32
f64 = np.float64
_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.powf.restype = ctypes.c_float
_libm.powf.argtypes = [ctypes.c_float, ctypes.c_float]
_libm.tanf.restype = ctypes.c_float
_libm.tanf.argtypes = [ctypes.c_float]

_builtin_float, _builtin_int, _builtin_max, _builtin_min, _builtin_pow = float, int, max, min, pow


# ----------------------------------------------------------------------------------------------------------- the tape
class Tape:
    """Scalar operations recorded by the forward calls: out = op(a, b) with the local partials (float64)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.out, self.a, self.da, self.b, self.db = [], [], [], [], []
        self.n_nodes = 0
        self.leaf = {}            # node -> (grad array, flat index)
        self.seg = None           # the running kernel's segment

    def node(self):
        self.n_nodes += 1
        return self.n_nodes - 1

    def rec(self, v, ia, da, ib=-1, db=0.0):
        o = self.node()
        self.out.append(o); self.a.append(ia); self.da.append(_builtin_float(da)); self.b.append(ib); self.db.append(_builtin_float(db))
        return F(v, o)


TAPE = Tape()


class Segment:
    def __init__(self, tape):
        self.t0 = len(tape.out)
        self.n0 = tape.n_nodes
        self.t1 = None
        self.stores = {}          # (id(field), flat index incl. component) -> (field, flat, node): the LAST store wins
        self.reads = {}           # node -> (grad array, flat)


# ------------------------------------------------------------------------------------------------------- fp32 scalars
class F:
    """One fp32 value; `i` is its tape node (-1: a constant w.r.t. every differentiable field)."""
    __slots__ = ("v", "i")

    def __init__(self, v, i=-1):
        self.v = v
        self.i = i

    def __repr__(self):
        return f"F({self.v!r}{'' if self.i < 0 else ', #%d' % self.i})"

    # arithmetic: one IEEE fp32 rounding per operator
    def __add__(s, o):
        if type(o) is Vec:
            return o.__radd__(s)
        o = _c(o); v = s.v + o.v
        return F(v) if (s.i < 0 and o.i < 0) else TAPE.rec(v, s.i, 1.0, o.i, 1.0)

    def __radd__(s, o):
        return _c(o).__add__(s)

    def __sub__(s, o):
        if type(o) is Vec:
            return o.__rsub__(s)
        o = _c(o); v = s.v - o.v
        return F(v) if (s.i < 0 and o.i < 0) else TAPE.rec(v, s.i, 1.0, o.i, -1.0)

    def __rsub__(s, o):
        return _c(o).__sub__(s)

    def __mul__(s, o):
        if type(o) is Vec:
            return o.__rmul__(s)
        o = _c(o); v = s.v * o.v
        return F(v) if (s.i < 0 and o.i < 0) else TAPE.rec(v, s.i, f64(o.v), o.i, f64(s.v))

    def __rmul__(s, o):
        return _c(o).__mul__(s)

    def __truediv__(s, o):
        if type(o) is Vec:
            return o.__rtruediv__(s)
        o = _c(o); v = s.v / o.v
        if s.i < 0 and o.i < 0:
            return F(v)
        a, b = f64(s.v), f64(o.v)
        return TAPE.rec(v, s.i, f64(1.0) / b, o.i, -a / (b * b))

    def __rtruediv__(s, o):
        return _c(o).__truediv__(s)

    def __neg__(s):
        return F(-s.v) if s.i < 0 else TAPE.rec(-s.v, s.i, -1.0)

    def __pow__(s, k):
        return ti_pow(s, k)

    def __pos__(s):
        return s

    # comparisons: plain fp32 comparisons (anything with a NaN is False)
    def __lt__(s, o): return bool(s.v < _c(o).v)
    def __le__(s, o): return bool(s.v <= _c(o).v)
    def __gt__(s, o): return bool(s.v > _c(o).v)
    def __ge__(s, o): return bool(s.v >= _c(o).v)
    def __eq__(s, o): return bool(s.v == _c(o).v)
    def __ne__(s, o): return bool(s.v != _c(o).v)
    __hash__ = None

    def __float__(s): return _builtin_float(s.v)
    def __int__(s): return _builtin_int(s.v)
    def __bool__(s): return bool(s.v != 0)


def _c(x):
    if type(x) is F:
        return x
    if isinstance(x, (bool, _builtin_int, _builtin_float, np.floating, np.integer)):
        return F(f32(x))
    raise TypeError(f"ti_shim: cannot use {type(x).__name__} as an fp32 scalar")


def _is_num(x):
    return isinstance(x, (bool, _builtin_int, _builtin_float, np.floating, np.integer))


# ----------------------------------------------------------------------------------------------------------- vectors
class Vec:
    """taichi.Vector (column vectors only): elementwise arithmetic, scalars broadcast.  Elements are F inside kernels and plain
    Python numbers in Python scope (ti.static expressions), where everything stays a double as in real Taichi."""
    __slots__ = ("e",)

    def __init__(self, e):
        self.e = list(e)

    def __repr__(self):
        return f"Vec({self.e})"

    @property
    def n(self): return len(self.e)
    def __len__(self): return len(self.e)
    def __iter__(self): return iter(self.e)
    def __getitem__(self, k): return self.e[k]
    def __setitem__(self, k, v): self.e[k] = v

    def _bin(self, o, op, swap=False):
        oe = o.e if type(o) is Vec else [o] * len(self.e)
        if len(oe) != len(self.e):
            raise ValueError("ti_shim: vector sizes differ")
        return Vec([op(y, x) if swap else op(x, y) for x, y in zip(self.e, oe)])

    def __add__(s, o): return s._bin(o, operator.add)
    def __radd__(s, o): return s._bin(o, operator.add, True)
    def __sub__(s, o): return s._bin(o, operator.sub)
    def __rsub__(s, o): return s._bin(o, operator.sub, True)
    def __mul__(s, o): return s._bin(o, operator.mul)
    def __rmul__(s, o): return s._bin(o, operator.mul, True)
    def __truediv__(s, o): return s._bin(o, operator.truediv)
    def __rtruediv__(s, o): return s._bin(o, operator.truediv, True)
    def __neg__(s): return Vec([-x for x in s.e])
    def __pow__(s, k): return Vec([ti_pow(x, k) for x in s.e])

    # swizzles the reference uses
    x = property(lambda s: s.e[0], lambda s, v: s.e.__setitem__(0, v))
    y = property(lambda s: s.e[1], lambda s, v: s.e.__setitem__(1, v))
    z = property(lambda s: s.e[2], lambda s, v: s.e.__setitem__(2, v))
    w = property(lambda s: s.e[3], lambda s, v: s.e.__setitem__(3, v))

    @property
    def xyz(self): return Vec(self.e[:3])

    @xyz.setter
    def xyz(self, v): self.e[:3] = list(v.e)

    def sum(self):
        r = self.e[0]
        for x in self.e[1:]:
            r = r + x
        return r

    def dot(self, o): return (self * o).sum()
    def norm_sqr(self): return (self * self).sum()
    def norm(self, eps=0): return sqrt(self.norm_sqr() + eps)

    def normalized(self, eps=0):
        invlen = 1 / (self.norm() + eps)
        return invlen * self

    def cross(self, o): return cross(self, o)


# ------------------------------------------------------------------------------------------ scalar functions (ti.*)
def sqrt(x):
    if type(x) is Vec:
        return Vec([sqrt(e) for e in x.e])
    if type(x) is not F:
        return math.sqrt(x)
    v = np.sqrt(x.v)
    return F(v) if x.i < 0 else TAPE.rec(v, x.i, f64(0.5) / f64(v))


def floor(x):
    if type(x) is Vec:
        return Vec([floor(e) for e in x.e])
    if type(x) is not F:
        return _builtin_float(math.floor(x))
    return F(np.floor(x.v))                          # no adjoint


def tan(x):
    if type(x) is not F:
        return math.tan(x)                           # a Python-scope constant: evaluated by Python, in double
    if x.i >= 0:
        raise NotImplementedError("ti_shim: tan of a differentiated value")
    return F(f32(_libm.tanf(x.v)))


def ti_pow(a, b):
    if type(a) is Vec:
        return Vec([ti_pow(e, b) for e in a.e])
    if type(a) is not F and type(b) is not F:
        return _builtin_pow(a, b)
    a, b = _c(a), _c(b)
    v = f32(_libm.powf(a.v, b.v))
    if a.i < 0 and b.i < 0:
        return F(v)
    a64, b64 = f64(a.v), f64(b.v)
    da = b64 * np.power(a64, b64 - 1.0)
    db = f64(v) * np.log(a64) if b.i >= 0 else 0.0
    return TAPE.rec(v, a.i, da, b.i, db)


def _max2(a, b):
    if type(a) is Vec or type(b) is Vec:
        av = a.e if type(a) is Vec else [a] * len(b.e)
        bv = b.e if type(b) is Vec else [b] * len(a.e)
        return Vec([_max2(x, y) for x, y in zip(av, bv)])
    if type(a) is not F and type(b) is not F:
        return a if a > b else b
    a, b = _c(a), _c(b)
    return a if a.v > b.v else b                     # the selected operand itself: its adjoint is the whole adjoint, the other gets none


def _min2(a, b):
    if type(a) is Vec or type(b) is Vec:
        av = a.e if type(a) is Vec else [a] * len(b.e)
        bv = b.e if type(b) is Vec else [b] * len(a.e)
        return Vec([_min2(x, y) for x, y in zip(av, bv)])
    if type(a) is not F and type(b) is not F:
        return a if a < b else b
    a, b = _c(a), _c(b)
    return a if a.v < b.v else b


def ti_max(*args):
    r = args[0]
    for x in args[1:]:
        r = _max2(r, x)
    return r


def ti_min(*args):
    r = args[0]
    for x in args[1:]:
        r = _min2(r, x)
    return r


def ti_float(x):
    if type(x) is F:
        return x
    if type(x) is Vec:
        return Vec([ti_float(e) for e in x.e])
    return F(f32(x))


def ti_int(x):
    if type(x) is F:
        return _builtin_int(x.v)                     # fp32 -> i32 truncates
    return _builtin_int(x)


BUILTINS = {"float": ti_float, "int": ti_int, "max": ti_max, "min": ti_min, "pow": ti_pow}


# ----------------------------------------------------------------------------------------------------------- fields
class _Axes:
    def __init__(self, ax): self.ax = ax


class SNode:
    def __init__(self, shape=()):
        self.shape = tuple(shape)

    def dense(self, axes, dims):
        dims = (dims,) * len(axes.ax) if _is_num(dims) else tuple(dims)
        shape = list(self.shape) + [1] * _builtin_max(0, _builtin_max(axes.ax) + 1 - len(self.shape))
        for a, d in zip(axes.ax, dims):
            shape[a] *= _builtin_int(d)
        return SNode(shape)

    def place(self, *fields):
        for f in fields:
            f._place(self.shape)

    def lazy_grad(self):
        pass


class GradField:
    def __init__(self, parent):
        self.parent = parent
        self.data = None

    def _place(self, shape):
        pass

    def _alloc(self):
        p = self.parent
        self.data = np.zeros(p.data.shape, np.float64)

    def fill(self, v):
        self.data[...] = v

    def from_torch(self, t):
        a = t.detach().cpu().numpy()
        if a.shape != self.data.shape:
            raise ValueError(f"ti_shim: grad.from_torch shape {a.shape} != {self.data.shape}")
        self.data[...] = a

    def to_torch(self, device=None):
        import torch
        return torch.from_numpy(self.data.astype(np.float32))

    def to_numpy64(self):
        return self.data.copy()

    def to_numpy(self):
        return self.data.astype(np.float32)

    def from_numpy(self, a):
        self.data[...] = a

    def __getitem__(self, idx):                      # a kernel READING a gradient (the example's apply_grad): plain fp32 constants
        idx = self.parent._idx(idx)
        if self.parent.n:
            return Vec([F(f32(self.data[idx + (c,)])) for c in range(self.parent.n)])
        return F(f32(self.data[idx]))

    def __setitem__(self, idx, val):
        idx = self.parent._idx(idx)
        if self.parent.n:
            vals = val.e if type(val) is Vec else [val] * self.parent.n
            for c, x in enumerate(vals):
                self.data[idx + (c,)] = _builtin_float(_c(x).v)
        else:
            self.data[idx] = _builtin_float(_c(val).v)


class Field:
    def __init__(self, dtype, n=0, needs_grad=False, shape=None):
        self.dtype, self.n, self.needs_grad = dtype, n, needs_grad
        self.shape = None
        self.data = None
        self.grad = GradField(self) if needs_grad else None
        self.cur = {}
        if shape is not None:
            self._place((shape,) if _is_num(shape) else tuple(shape))

    def _place(self, shape):
        self.shape = tuple(shape)
        full = self.shape + ((self.n,) if self.n else ())
        self.data = np.zeros(full, np.int32 if self.dtype == "i32" else np.float32)
        self.cur = {}
        if self.grad is not None:
            self.grad._alloc()

    # host side
    def from_torch(self, t):
        a = t.detach().cpu().numpy()
        if a.shape != self.data.shape:
            raise ValueError(f"ti_shim: from_torch shape {a.shape} != field shape {self.data.shape}")
        self.data[...] = a
        self.cur = {}

    def to_torch(self, device=None):
        import torch
        return torch.from_numpy(self.data.copy())

    def from_numpy(self, a):
        a = np.asarray(a)
        if a.shape != self.data.shape:
            raise ValueError(f"ti_shim: from_numpy shape {a.shape} != field shape {self.data.shape}")
        self.data[...] = a
        self.cur = {}

    def to_numpy(self):
        return self.data.copy()

    def fill(self, v):
        self.data[...] = v
        self.cur = {}

    def __iter__(self):
        return iter(np.ndindex(*self.shape))

    def _idx(self, idx):
        if idx is None:
            idx = ()
        elif not isinstance(idx, tuple):
            idx = (idx,)
        idx = tuple(ti_int(k) for k in idx)
        if len(idx) != len(self.shape):
            raise IndexError(f"ti_shim: {len(idx)} indices for a {len(self.shape)}-d field")
        return idx

    def _leaf(self, flat, v):
        """The tape node of one stored fp32 element (created on first use)."""
        x = self.cur.get(flat)
        if x is None:
            x = F(v, TAPE.node()) if self.needs_grad else F(v)
            if self.needs_grad:
                TAPE.leaf[x.i] = (self.grad.data, flat)
            self.cur[flat] = x
        if x.i >= 0 and TAPE.seg is not None and x.i not in TAPE.seg.reads:
            TAPE.seg.reads[x.i] = (self.grad.data, flat)
        return x

    def __getitem__(self, idx):
        idx = self._idx(idx)
        if any(k < 0 for k in idx):                                       # tape[i, j, -1] is the cleared tape
            if self.dtype == "i32":
                return 0
            return Vec([F(f32(0))] * self.n) if self.n else F(f32(0))
        for k, s in zip(idx, self.shape):
            if k >= s:
                raise IndexError(f"ti_shim: index {idx} outside a field of shape {self.shape} (SURVEY 7.3 H2: choose max_samples >= the longest ray)")
        if self.dtype == "i32":
            return _builtin_int(self.data[idx])
        if self.n:
            return Vec([self._leaf(idx + (c,), self.data[idx + (c,)]) for c in range(self.n)])
        return self._leaf(idx, self.data[idx])

    def __setitem__(self, idx, val):
        idx = self._idx(idx)
        if self.dtype == "i32":
            self.data[idx] = ti_int(val)
            return
        if self.n:
            vals = val.e if type(val) is Vec else [val] * self.n
            if len(vals) != self.n:
                raise ValueError("ti_shim: vector width mismatch in a field store")
            for c, x in enumerate(vals):
                self._store(idx + (c,), _c(x))
        else:
            self._store(idx, _c(val))

    def _store(self, flat, x):
        self.data[flat] = x.v
        self.cur[flat] = x
        if self.needs_grad and TAPE.seg is not None:
            TAPE.seg.stores[(id(self), flat)] = (self.grad.data, flat, x.i)


class _VectorNS:
    @staticmethod
    def field(n, dtype=None, shape=None, needs_grad=False):
        return Field(dtype, n=n, needs_grad=needs_grad, shape=shape)

    def __call__(self, e):
        return Vec(e)


def field(dtype, shape=None, needs_grad=False):
    return Field(dtype, n=0, needs_grad=needs_grad, shape=shape)


# ---------------------------------------------------------------------------------------------------------- kernels
class _BoundKernel:
    def __init__(self, k, obj):
        self.k, self.obj = k, obj

    def _args(self, args):
        import inspect
        params = list(inspect.signature(self.k.fn).parameters.values())[1:]
        out = []
        for p, a in zip(params, args):
            if p.annotation is ti_float or p.annotation is _builtin_float:
                a = ti_float(a)
            elif p.annotation is ti_int or p.annotation is _builtin_int:
                a = _builtin_int(a)
            out.append(a)
        return out

    def __call__(self, *args):
        seg = Segment(TAPE)
        TAPE.seg = seg
        try:
            with np.errstate(all="ignore"):          # 1/0, 0*inf, inf-inf are ordinary fp32 results here, not warnings
                self.k.fn(self.obj, *self._args(args))
        finally:
            TAPE.seg = None
        seg.t1 = len(TAPE.out)
        self.obj.__dict__.setdefault("_shim_segments", {})[self.k.fn.__name__] = seg

    def grad(self, *args):
        """Reverse sweep over the operations the LAST forward call of this kernel recorded: seeds come from the .grad of the
        elements it stored, results are added to the .grad of the elements it read."""
        seg = self.obj.__dict__.get("_shim_segments", {}).get(self.k.fn.__name__)
        if seg is None:
            raise RuntimeError(f"ti_shim: {self.k.fn.__name__}.grad() before the forward call")
        adj = [0.0] * TAPE.n_nodes
        for g, flat, node in seg.stores.values():
            if node >= 0:
                adj[node] += _builtin_float(g[flat])
        out, a, da, b, db = TAPE.out, TAPE.a, TAPE.da, TAPE.b, TAPE.db
        for k in range(seg.t1 - 1, seg.t0 - 1, -1):
            gk = adj[out[k]]                         # NOT skipped when zero: Taichi's adjoint code multiplies unconditionally (0 x inf = NaN)
            ia = a[k]
            if ia >= 0:
                adj[ia] += gk * da[k]
            ib = b[k]
            if ib >= 0:
                adj[ib] += gk * db[k]
        for node, (g, flat) in seg.reads.items():
            if node < seg.n0 or node in TAPE.leaf:   # not produced inside this call: a global load, its adjoint goes to the field's .grad
                g[flat] += adj[node]


class _Kernel:
    def __init__(self, fn):
        self.fn = fn

    def __get__(self, obj, cls=None):
        return self if obj is None else _BoundKernel(self, obj)


# ---------------------------------------------------------------------------------------------- taichi_glsl functions
def _flatten(args):
    e = []
    for a in args:
        if type(a) is Vec:
            e.extend(a.e)
        elif isinstance(a, (list, tuple)):
            e.extend(_flatten(a))
        else:
            e.append(a)
    return e


def _vecn(n):
    def make(*args):
        e = _flatten(args)
        if len(e) == 1:
            e = e * n
        if len(e) != n:
            raise ValueError(f"ti_shim: vec{n} from {len(e)} components")
        return Vec(e)
    return make


def mix(x, y, a):
    return x * (1 - a) + y * a


def clamp(x, xmin=0, xmax=1):
    return ti_min(xmax, ti_max(xmin, x))


def cross(a, b):
    return Vec([a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x])


def dot(a, b):
    return a.dot(b)


def reflect(I, N):
    return I - 2 * N.dot(I) * N


def normalize(v):
    return v.normalized()


# --------------------------------------------------------------------------------------------------- the two modules
_random_source = []


def set_random_source(values):
    """The numbers the next ti.random() calls return, in call order (the reference draws one per pixel in struct-for order, so a
    jitter image in the raw (w, h) layout, flattened, reproduces a supplied jitter tensor without touching the source)."""
    global _random_source
    _random_source = [f32(v) for v in np.asarray(values, np.float32).reshape(-1)][::-1]


def ti_random(dtype=None):
    if not _random_source:
        raise RuntimeError("ti_shim: ti.random() called with no numbers left (set_random_source)")
    return F(_random_source.pop())


def make_modules():
    """(taichi, taichi_glsl) stand-ins."""
    ti = types.ModuleType("taichi")
    ti.__shim__ = True
    ti.f32, ti.i32, ti.f64 = "f32", "i32", "f64"
    ti.cpu, ti.cuda, ti.gpu = "cpu", "cuda", "gpu"
    ti.i, ti.j, ti.k = _Axes((0,)), _Axes((1,)), _Axes((2,))
    ti.ij, ti.ijk = _Axes((0, 1)), _Axes((0, 1, 2))
    ti.init = lambda *a, **k: None
    ti.func = lambda fn: fn
    ti.kernel = _Kernel
    ti.data_oriented = lambda cls: cls
    ti.static = lambda *x: x[0] if len(x) == 1 else x
    ti.field = field
    ti.Vector = _VectorNS()
    ti.root = SNode(())
    ti.max, ti.min, ti.floor, ti.tan, ti.pow, ti.sqrt, ti.random = ti_max, ti_min, floor, tan, ti_pow, sqrt, ti_random
    ti.cast = lambda x, t: ti_float(x) if t == "f32" else ti_int(x)
    tl = types.ModuleType("taichi_glsl")
    tl.__shim__ = True
    tl.vec2, tl.vec3, tl.vec4 = _vecn(2), _vecn(3), _vecn(4)
    tl.mix, tl.clamp, tl.cross, tl.dot, tl.reflect, tl.normalize = mix, clamp, cross, dot, reflect, normalize
    tl.summation = lambda v: v.sum()
    return ti, tl


def load_reference(source, path="<reference>", extra_modules=None):
    """Executes the reference module's SOURCE TEXT (as read from the reference tree by the caller) against the stand-in modules, with
    the kernel builtins re-interpreted; returns the module object.  `extra_modules`: {name: module} stubs for imports of the file that
    are not on the path computed here (plotting, torchvtk)."""
    ti, tl = make_modules()
    mods = {"taichi": ti, "taichi_glsl": tl}
    mods.update(extra_modules or {})
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    try:
        mod = types.ModuleType("differender_reference_on_ti_shim")
        mod.__dict__.update(BUILTINS)
        exec(compile(source, path + " [on oracle/ti_shim.py]", "exec"), mod.__dict__)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    TAPE.reset()
    return mod


def reset():
    TAPE.reset()
