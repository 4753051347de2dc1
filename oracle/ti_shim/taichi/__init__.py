"""Minimal stand-in for the `taichi` package -- TEST INFRASTRUCTURE ONLY.

Purpose: Taichi is not installable in this environment (no network, not on
disk), yet the reference (AntonioFerreras/Digital-Earth) is pure Python whose
device code is written as `@ti.func`s.  This shim supplies just enough of the
`taichi` surface for the reference's *own, unmodified source files* under
/root/reference to be imported and executed one scalar path at a time, so that
golden vectors for the oracle can be generated from the reference itself
(tests/golden/gen_golden.py).  It is never imported by the product.

Semantics fixed here (documented in DESIGN.md "oracle semantics"):
  * default_fp = f32, default_ip = i32.  Every Taichi *variable* (assignment
    target, ti.func argument, return value, struct / vector / field element)
    is rounded to f32 (or truncated to i32); Python literals and Python-scope
    constants stay "weak" Python values until they meet a variable, so
    constant-only sub-expressions fold in f64 exactly like Taichi's AST
    builder does (`np.pi*2`, `3.0/(16.0*np.pi)`, `pow(np.pi, 3.0)` ...).
  * +,-,*,/ are IEEE binary32, evaluated in source order, no FMA contraction.
  * transcendental functions are glibc's float versions (expf, logf, powf,
    sinf, cosf, atan2f, asinf ...) called through ctypes, so a C restatement
    compiled with -ffp-contract=off reproduces results bit for bit.
  * pow(x, 2.0) is x*x (the fold every backend compiler performs); other exponents
    call powf.
  * max/min are IEEE maxNum/minNum (LLVM maxnum/minnum == C fmaxf/fminf).
  * ti.random() = (u32 >> 8) * 2^-24 from a pluggable source.
  * Texture.sample_lod = manual FP32 bilinear, texel centres at (i+.5)/N,
    clamp-to-edge, lerp(a,b,f) = a + f*(b-a), x first then y.
  * ti.func arguments are passed by value (vectors are copied on entry).
  * range() bounds are truncated to i32 (`for x in range(0, log2(441))` -> 8).
"""
import ast as _ast
import ctypes as _ct
import ctypes.util as _ctu
import functools as _ft
import inspect as _inspect
import math as _pm
import textwrap as _tw

import numpy as _np

_np.seterr(all="ignore")

f32 = _np.float32
f64 = _np.float64
f16 = _np.float16
i32 = _np.int32
u8 = _np.uint8
u32 = _np.uint32

vulkan = "vulkan"
cpu = "cpu"
gpu = "gpu"

_libm = _ct.CDLL(_ctu.find_library("m") or "libm.so.6")
for _n in ("expf", "logf", "sinf", "cosf", "tanf", "asinf", "acosf", "sqrtf", "floorf", "ceilf", "tanhf", "fabsf"):
    getattr(_libm, _n).restype = _ct.c_float
    getattr(_libm, _n).argtypes = [_ct.c_float]
for _n in ("powf", "atan2f", "fmaxf", "fminf"):
    getattr(_libm, _n).restype = _ct.c_float
    getattr(_libm, _n).argtypes = [_ct.c_float, _ct.c_float]


# --------------------------------------------------------------------------
# scalar typing
# --------------------------------------------------------------------------
class I32(int):
    """A Taichi i32 variable.  int op int -> i32, int op float -> f32."""

    def _bin(self, other, fi, ff, rev=False):
        if isinstance(other, (bool, _np.bool_)):
            other = int(other)
        if isinstance(other, _np.floating) or isinstance(other, float):
            a, b = f32(int(self)), f32(other)
            return f32(ff(b, a) if rev else ff(a, b))
        if isinstance(other, int):
            r = fi(int(other), int(self)) if rev else fi(int(self), int(other))
            return I32(_wrap32(r))
        return NotImplemented

    def __add__(self, o): return self._bin(o, lambda a, b: a + b, lambda a, b: a + b)
    def __radd__(self, o): return self._bin(o, lambda a, b: a + b, lambda a, b: a + b, True)
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b, lambda a, b: a - b)
    def __rsub__(self, o): return self._bin(o, lambda a, b: a - b, lambda a, b: a - b, True)
    def __mul__(self, o): return self._bin(o, lambda a, b: a * b, lambda a, b: a * b)
    def __rmul__(self, o): return self._bin(o, lambda a, b: a * b, lambda a, b: a * b, True)

    def __truediv__(self, o):
        if isinstance(o, Vector):
            return NotImplemented
        return f32(f32(int(self)) / f32(o))

    def __rtruediv__(self, o):
        return f32(f32(o) / f32(int(self)))

    def __neg__(self): return I32(-int(self))


def _wrap32(v):
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v & 0x80000000 else v


def _is_py(x):
    """Python-scope (compile-time) value: folds in f64 like Taichi's AST builder."""
    if isinstance(x, (I32, _np.float32, _np.float16, _np.integer, _np.bool_)):
        return False
    return isinstance(x, (int, float))  # includes np.float64 and bool


def _var(x):
    """Materialise a Taichi variable from a value (expr_init)."""
    if isinstance(x, (bool, _np.bool_)):
        return x
    if isinstance(x, I32):
        return x
    if isinstance(x, _np.float32):
        return x
    if isinstance(x, (_np.floating, float)):
        return f32(x)
    if isinstance(x, (int, _np.integer)):
        return I32(_wrap32(int(x)))
    if isinstance(x, Vector):
        return Vector(x.e)
    if isinstance(x, Matrix):
        return Matrix([list(r) for r in x.m])
    if isinstance(x, tuple):
        return tuple(_var(v) for v in x)
    return x


def _f(x):
    """Cast one scalar to an f32 variable value."""
    if isinstance(x, _np.float32):
        return x
    return f32(x)


def _range(*a):
    return range(*[int(v) for v in a])


# --------------------------------------------------------------------------
# vectors / matrices
# --------------------------------------------------------------------------
_SWZ = {c: i for i, c in enumerate("xyzw")}
_SWZ.update({c: i for i, c in enumerate("rgba")})


def _bcast(a, b):
    if isinstance(a, Vector) and isinstance(b, Vector):
        assert len(a.e) == len(b.e), "vector size mismatch"
        return a.e, b.e
    if isinstance(a, Vector):
        return a.e, [b] * len(a.e)
    return [a] * len(b.e), b.e


def _sc(v):
    """Scalar that enters a vector op: weak Python values become f32 (vector
    entries are always variables in Taichi)."""
    if isinstance(v, (I32, _np.float32, bool, _np.bool_)):
        return v
    if isinstance(v, (float, _np.floating)):
        return f32(v)
    if isinstance(v, (int, _np.integer)):
        return I32(int(v))
    return v


class Vector:
    __slots__ = ("e",)
    __array_ufunc__ = None  # make numpy scalars defer to our reflected ops

    def __init__(self, entries, dt=None):
        out = []
        for v in entries:
            if isinstance(v, Vector):
                out.extend(v.e)
            else:
                out.append(v)
        if dt is i32 or (len(out) and all(isinstance(v, I32) for v in out)):
            out = [I32(int(v)) for v in out]
        else:
            out = [_f(v) for v in out]
        object.__setattr__(self, "e", out)

    @staticmethod
    def field(n, dtype, shape=None):
        return _Field(dtype, shape, n)

    # -- element access
    def __len__(self): return len(self.e)
    def __iter__(self): return iter(self.e)
    def __getitem__(self, i): return self.e[int(i)]

    def __setitem__(self, i, v):
        self.e[int(i)] = _f(v) if not isinstance(self.e[int(i)], I32) else I32(int(v))

    def __getattr__(self, name):
        try:
            idx = [_SWZ[c] for c in name]
        except KeyError:
            raise AttributeError(name)
        if len(idx) == 1:
            return self.e[idx[0]]
        return Vector([self.e[i] for i in idx])

    def __setattr__(self, name, value):
        idx = [_SWZ[c] for c in name]
        if len(idx) == 1:
            self.e[idx[0]] = _f(value)
        else:
            for k, i in enumerate(idx):
                self.e[i] = _f(value[k])

    @property
    def n(self): return len(self.e)

    # -- arithmetic (elementwise, IEEE f32 via numpy scalars)
    def _op(self, o, fn, rev=False):
        if isinstance(o, Matrix):
            return NotImplemented
        o = o if isinstance(o, Vector) else _sc(o)
        a, b = _bcast(self, o)
        return Vector([fn(y, x) if rev else fn(x, y) for x, y in zip(a, b)])

    def __add__(self, o): return self._op(o, lambda a, b: a + b)
    def __radd__(self, o): return self._op(o, lambda a, b: a + b, True)
    def __sub__(self, o): return self._op(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._op(o, lambda a, b: a - b, True)
    def __mul__(self, o): return self._op(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._op(o, lambda a, b: a * b, True)
    def __truediv__(self, o): return self._op(o, lambda a, b: a / b)
    def __rtruediv__(self, o): return self._op(o, lambda a, b: a / b, True)
    def __neg__(self): return Vector([-v for v in self.e])
    def __pow__(self, o): return pow(self, o)

    def __matmul__(self, o):
        if isinstance(o, Matrix):  # row vector times matrix
            cols = len(o.m[0])
            out = []
            for j in range(cols):
                acc = self.e[0] * o.m[0][j]
                for i in range(1, len(self.e)):
                    acc = acc + self.e[i] * o.m[i][j]
                out.append(acc)
            return Vector(out)
        return NotImplemented

    # -- reductions (left-to-right, as taichi.lang.matrix does)
    def sum(self):
        acc = self.e[0]
        for v in self.e[1:]:
            acc = acc + v
        return acc

    def dot(self, o): return (self * o).sum()
    def norm_sqr(self): return (self * self).sum()
    def norm(self, eps=0): return sqrt(self.norm_sqr() + eps) if eps else sqrt(self.norm_sqr())

    def normalized(self, eps=0):
        invlen = 1 / (self.norm() + eps) if eps else 1 / self.norm()
        return invlen * self

    def cross(self, o):
        a, b = self.e, o.e
        return Vector([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])

    def __repr__(self): return "Vector(" + ", ".join(repr(float(v)) for v in self.e) + ")"


class Matrix:
    __slots__ = ("m",)
    __array_ufunc__ = None

    def __init__(self, rows):
        object.__setattr__(self, "m", [[_f(v) for v in r] for r in rows])

    def __getitem__(self, ij):
        i, j = ij
        return self.m[int(i)][int(j)]

    def __setitem__(self, ij, v):
        i, j = ij
        self.m[int(i)][int(j)] = _f(v)

    def transpose(self):
        return Matrix([list(c) for c in zip(*self.m)])

    def __matmul__(self, o):
        if isinstance(o, Vector):
            out = []
            for r in self.m:
                acc = r[0] * o.e[0]
                for k in range(1, len(r)):
                    acc = acc + r[k] * o.e[k]
                out.append(acc)
            return Vector(out)
        if isinstance(o, Matrix):
            ot = list(zip(*o.m))
            res = []
            for r in self.m:
                row = []
                for c in ot:
                    acc = r[0] * c[0]
                    for k in range(1, len(r)):
                        acc = acc + r[k] * c[k]
                    row.append(acc)
                res.append(row)
            return Matrix(res)
        return NotImplemented

    def __repr__(self): return "Matrix(%r)" % ([[float(v) for v in r] for r in self.m],)


class _VecType:
    """vecN: usable as constructor and as annotation."""

    def __init__(self, n): self.n = n

    def __call__(self, *args):
        flat = []
        for a in args:
            if isinstance(a, (Vector, list, tuple)):
                flat.extend(list(a))
            else:
                flat.append(a)
        if len(flat) == 1:
            flat = flat * self.n
        assert len(flat) == self.n, "vec%d built from %d entries" % (self.n, len(flat))
        return Vector(flat)


class _MatType:
    def __init__(self, n): self.n = n

    def __call__(self, *args):
        n = self.n
        flat = []
        for a in args:
            if isinstance(a, (Vector, list, tuple)):
                flat.extend(list(a))
            else:
                flat.append(a)
        if len(flat) == 1:
            flat = flat * (n * n)
        assert len(flat) == n * n
        return Matrix([flat[i * n:(i + 1) * n] for i in range(n)])


# --------------------------------------------------------------------------
# ops (taichi.lang.ops): fold in Python when every operand is a Python value
# --------------------------------------------------------------------------
def _unary(cname, pyfn):
    cfn = getattr(_libm, cname)

    def op(x):
        if isinstance(x, Vector):
            return Vector([op(c) for c in x.e])
        if _is_py(x):
            return pyfn(x)
        return f32(cfn(float(f32(x))))
    op.__name__ = cname[:-1]
    return op


def _py_sqrt(x): return _pm.sqrt(x) if x >= 0 else float("nan")
def _py_log(x): return _pm.log(x) if x > 0 else (float("-inf") if x == 0 else float("nan"))


exp = _unary("expf", _pm.exp)
log = _unary("logf", _py_log)
sin = _unary("sinf", _pm.sin)
cos = _unary("cosf", _pm.cos)
tan = _unary("tanf", _pm.tan)
asin = _unary("asinf", _pm.asin)
acos = _unary("acosf", _pm.acos)
tanh = _unary("tanhf", _pm.tanh)
floor = _unary("floorf", lambda v: float(_pm.floor(v)))
ceil = _unary("ceilf", lambda v: float(_pm.ceil(v)))


def sqrt(x):
    if isinstance(x, Vector):
        return Vector([sqrt(c) for c in x.e])
    if _is_py(x):
        return _py_sqrt(x)
    return _np.sqrt(f32(x))  # IEEE correctly rounded


def _binary(fn_f32, fn_py):
    def op(a, b):
        if isinstance(a, Vector) or isinstance(b, Vector):
            a = a if isinstance(a, Vector) else _sc(a)
            b = b if isinstance(b, Vector) else _sc(b)
            x, y = _bcast(a, b)
            return Vector([op(p, q) for p, q in zip(x, y)])
        if _is_py(a) and _is_py(b):
            return fn_py(a, b)
        return fn_f32(a, b)
    return op


def _powf(a, b):
    a, b = f32(a), f32(b)
    if b == 2.0:  # LLVM/NVVM/gcc fold pow(x, 2.0) -> x*x even without fast-math
        return a * a
    return f32(_libm.powf(float(a), float(b)))


def _pypow(a, b):
    try:
        return float(a) ** float(b)
    except (ValueError, ZeroDivisionError, OverflowError):
        return float("nan")


pow = _binary(_powf, _pypow)
atan2 = _binary(lambda a, b: f32(_libm.atan2f(float(f32(a)), float(f32(b)))), _pm.atan2)


def _max2_f32(a, b):
    if isinstance(a, I32) and isinstance(b, I32):
        return I32(int.__int__(a) if int(a) > int(b) else int(b))
    return f32(_libm.fmaxf(float(f32(a)), float(f32(b))))


def _min2_f32(a, b):
    if isinstance(a, I32) and isinstance(b, I32):
        return I32(int(a) if int(a) < int(b) else int(b))
    return f32(_libm.fminf(float(f32(a)), float(f32(b))))


_max2 = _binary(_max2_f32, lambda a, b: a if a > b else b)
_min2 = _binary(_min2_f32, lambda a, b: a if a < b else b)


def max(*args):
    return _ft.reduce(_max2, args)


def min(*args):
    return _ft.reduce(_min2, args)


def abs(x):
    if isinstance(x, Vector):
        return Vector([abs(c) for c in x.e])
    if isinstance(x, I32):
        return I32(-int(x) if int(x) < 0 else int(x))
    if _is_py(x):
        return -x if x < 0 else x
    return f32(_np.abs(f32(x)))


def cast(x, dt):
    if isinstance(x, Vector):
        return Vector([cast(c, dt) for c in x.e])
    if dt in (i32, int):
        return I32(_wrap32(int(x)))  # float -> int truncates toward zero
    if dt in (f32, float):
        return f32(x)
    if dt is u8:
        return I32(int(x) & 0xFF)
    if dt is f16:
        return f32(_np.float16(x))
    raise TypeError(dt)


def select(c, a, b):
    return _var(a) if c else _var(b)


def static(x):
    return x


def ndrange(*a):
    import itertools
    return itertools.product(*[range(int(v)) for v in a])


def loop_config(**kw):
    pass


# -- random ------------------------------------------------------------------
_random_source = None


def _set_random_source(fn):
    """fn() -> uint32.  Installed by the golden-vector generator."""
    global _random_source
    _random_source = fn


def random(dtype=f32):
    assert _random_source is not None, "no random source installed"
    bits = int(_random_source()) & 0xFFFFFFFF
    return f32(bits >> 8) * f32(1.0 / 16777216.0)


# --------------------------------------------------------------------------
# fields / textures / structs
# --------------------------------------------------------------------------
class _Field:
    def __init__(self, dtype, shape=None, n=0):
        self.dtype, self.n = dtype, n
        if shape is None:
            shape = None
        elif isinstance(shape, int):
            shape = (shape,)
        self.shape = shape
        self.arr = None
        if shape is not None:
            self._alloc()

    def _alloc(self):
        npdt = {f32: _np.float32, i32: _np.int32, u8: _np.uint8, f16: _np.float16, float: _np.float32, int: _np.int32}[self.dtype]
        full = tuple(self.shape) + ((self.n,) if self.n else ())
        self.arr = _np.zeros(full, dtype=npdt)

    def from_numpy(self, a):
        self.arr[...] = a

    def to_numpy(self):
        return self.arr.copy()

    def fill(self, v):
        self.arr[...] = v

    def _key(self, k):
        if k is None:
            return ()
        if isinstance(k, tuple):
            return tuple(int(v) for v in k)
        return (int(k),)

    def __getitem__(self, k):
        v = self.arr[self._key(k)]
        if self.n:
            if self.dtype in (i32, int):
                return Vector([I32(int(c)) for c in v], dt=i32)
            return Vector([f32(c) for c in v])
        if self.dtype in (i32, int, u8):
            return I32(int(v))
        return f32(v)

    def __setitem__(self, k, v):
        if self.n:
            self.arr[self._key(k)] = [float(c) for c in v]
        else:
            self.arr[self._key(k)] = v


def field(dtype, shape=None):
    return _Field(dtype, shape)


class Format:
    r8 = "r8"; rgba8 = "rgba8"; rgba16f = "rgba16f"; rgba32f = "rgba32f"; r32f = "r32f"


class Texture:
    """2-D texture, storage [x][y][c] float32 (already /255 or f16-rounded by
    whoever fills it, as the reference's copy_* kernels do)."""

    def __init__(self, fmt, shape):
        self.fmt = fmt
        self.shape = tuple(shape)
        self.data = _np.zeros(self.shape + (4,), dtype=_np.float32)

    def set_data(self, arr):
        """arr: [x][y] or [x][y][c<=4] float32 texel values."""
        a = _np.asarray(arr, dtype=_np.float32)
        if a.ndim == 2:
            a = a[:, :, None]
        self.data[...] = 0
        self.data[:, :, :a.shape[2]] = a
        if self.fmt == Format.rgba16f:
            self.data = self.data.astype(_np.float16).astype(_np.float32)

    def sample_lod(self, uv, lod):
        w, h = self.shape
        u, v = _f(uv[0]), _f(uv[1])
        x = u * f32(w) - f32(0.5)
        y = v * f32(h) - f32(0.5)
        x0f = floor(x)
        y0f = floor(y)
        fx = x - x0f
        fy = y - y0f
        x0 = int(x0f); y0 = int(y0f)
        x1 = x0 + 1; y1 = y0 + 1
        x0 = 0 if x0 < 0 else (w - 1 if x0 > w - 1 else x0)
        x1 = 0 if x1 < 0 else (w - 1 if x1 > w - 1 else x1)
        y0 = 0 if y0 < 0 else (h - 1 if y0 > h - 1 else y0)
        y1 = 0 if y1 < 0 else (h - 1 if y1 > h - 1 else y1)
        d = self.data
        out = []
        for c in range(4):
            t00 = f32(d[x0, y0, c]); t10 = f32(d[x1, y0, c])
            t01 = f32(d[x0, y1, c]); t11 = f32(d[x1, y1, c])
            a = t00 + fx * (t10 - t00)
            b = t01 + fx * (t11 - t01)
            out.append(a + fy * (b - a))
        return Vector(out)


class _TexAnn:
    def __call__(self, *a, **k): return self


class types:
    texture = _TexAnn()
    rw_texture = _TexAnn()
    vector = staticmethod(lambda n, dt: _VecType(n))


class _Template:
    pass


def template():
    return _Template()


class _Dense:
    def dense(self, *a, **k): return self
    def place(self, *fields): return self


class _Root(_Dense):
    pass


root = _Root()
ij = "ij"


class tools:
    @staticmethod
    def imread(path):
        raise RuntimeError("ti.tools.imread is not available in the shim")

    class image:
        @staticmethod
        def imwrite(img, path):
            raise RuntimeError("not available in the shim")


class ui:
    RMB = CTRL = SPACE = SHIFT = None


def init(**kw):
    pass


def data_oriented(cls):
    return cls


def dataclass(cls):
    ann = dict(cls.__annotations__)

    def _zero(t):
        if isinstance(t, _VecType):
            return Vector([f32(0)] * t.n)
        if t in (int, i32):
            return I32(0)
        return f32(0)

    def __init__(self, **kw):
        for k, t in ann.items():
            object.__setattr__(self, k, _zero(t))
        for k, v in kw.items():
            setattr(self, k, v)

    def __setattr__(self, k, v):
        t = ann[k]
        if isinstance(t, _VecType):
            v = Vector(list(v))
        elif t in (int, i32):
            v = I32(int(v))
        else:
            v = f32(v)
        object.__setattr__(self, k, v)

    return type(cls.__name__, (), {"__init__": __init__, "__setattr__": __setattr__, "_ann": ann})


# --------------------------------------------------------------------------
# @ti.func / @ti.kernel: source-level rewrite to Taichi variable semantics
# --------------------------------------------------------------------------
class _Rewriter(_ast.NodeTransformer):
    def __init__(self, assigned):
        self.assigned = assigned

    @staticmethod
    def _wrap(node):
        return _ast.Call(func=_ast.Name(id="__ti_var", ctx=_ast.Load()), args=[node], keywords=[])

    def visit_Assign(self, node):
        self.generic_visit(node)
        if all(isinstance(t, _ast.Name) for t in node.targets):
            node.value = self._wrap(node.value)
        return node

    def visit_AnnAssign(self, node):
        self.generic_visit(node)
        if node.value is not None and isinstance(node.target, _ast.Name):
            node.value = self._wrap(node.value)
        return node

    def visit_Return(self, node):
        self.generic_visit(node)
        if node.value is not None:
            node.value = self._wrap(node.value)
        return node

    def visit_Call(self, node):
        self.generic_visit(node)
        if isinstance(node.func, _ast.Name):
            if node.func.id == "range":
                node.func = _ast.Name(id="__ti_range", ctx=_ast.Load())
            elif node.func.id in self.assigned:
                # name is also a local variable (e.g. `land_normal = land_normal(...)`):
                # Taichi resolves the call to the global function at build time.
                node.func = _ast.Subscript(
                    value=_ast.Call(func=_ast.Name(id="globals", ctx=_ast.Load()), args=[], keywords=[]),
                    slice=_ast.Constant(value=node.func.id), ctx=_ast.Load())
        return node


def _is_template_ann(a):
    if a is None:
        return False
    src = _ast.unparse(a)
    return "template" in src or "texture" in src


def _compile(fn):
    src = _tw.dedent(_inspect.getsource(fn))
    tree = _ast.parse(src)
    fdef = tree.body[0]
    assert isinstance(fdef, _ast.FunctionDef)
    fdef.decorator_list = []
    fdef.returns = None
    assigned = set()
    for n in _ast.walk(fdef):
        if isinstance(n, _ast.Assign):
            for t in n.targets:
                for nn in _ast.walk(t):
                    if isinstance(nn, _ast.Name):
                        assigned.add(nn.id)
    called = {n.func.id for n in _ast.walk(fdef) if isinstance(n, _ast.Call) and isinstance(n.func, _ast.Name)}
    shadowing = {n for n in assigned & called if callable(fn.__globals__.get(n))}
    _Rewriter(shadowing).visit(fdef)
    # pass-by-value prologue
    pro = []
    for a in fdef.args.args:
        if a.arg == "self" or _is_template_ann(a.annotation):
            a.annotation = None
            continue
        a.annotation = None
        pro.append(_ast.Assign(
            targets=[_ast.Name(id=a.arg, ctx=_ast.Store())],
            value=_ast.Call(func=_ast.Name(id="__ti_var", ctx=_ast.Load()), args=[_ast.Name(id=a.arg, ctx=_ast.Load())], keywords=[])))
    fdef.body = pro + fdef.body
    _ast.fix_missing_locations(tree)
    g = fn.__globals__
    g.setdefault("__ti_var", _var)
    g.setdefault("__ti_range", _range)
    loc = {}
    code = compile(tree, filename="<ti_shim:%s>" % fn.__qualname__, mode="exec")
    exec(code, g, loc)
    return loc[fdef.name]


class _LazyFunc:
    def __init__(self, fn):
        self._fn = fn
        self._c = None
        _ft.update_wrapper(self, fn)

    def _get(self):
        if self._c is None:
            self._c = _compile(self._fn)
        return self._c

    def __call__(self, *a, **k):
        return self._get()(*a, **k)

    def __get__(self, obj, objtype=None):
        if obj is None:
            return self
        return _ft.partial(self.__call__, obj)


def func(fn):
    return _LazyFunc(fn)


def kernel(fn):
    return _LazyFunc(fn)


from . import math  # noqa: E402  (ti.math.*)
