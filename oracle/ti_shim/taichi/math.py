"""taichi.math stand-in (see package docstring).  Everything defined here is a
Taichi `@func` in the real package, so arguments are variables (f32), never
folded in f64 -- except the names re-exported from taichi.lang.ops (exp, log,
pow, sqrt, sin, cos, max, min, abs ...) which fold on Python-only operands."""
import numpy as _np

from . import (Vector, Matrix, _VecType, _MatType, _var, _f, I32, f32,  # noqa: F401
               exp, log, sin, cos, tan, asin, acos, tanh, floor, ceil, sqrt, pow, atan2, max, min, abs)

pi = 3.141592653589793
e = 2.718281828459045
inf = float("inf")
nan = float("nan")

vec2 = _VecType(2)
vec3 = _VecType(3)
vec4 = _VecType(4)
mat2 = _MatType(2)
mat3 = _MatType(3)
mat4 = _MatType(4)


def _v(x):
    return _var(x)


def mix(x, y, a):
    x, y, a = _v(x), _v(y), _v(a)
    return x * (1.0 - a) + y * a


def clamp(x, xmin, xmax):
    return min(_v(xmax), max(_v(xmin), _v(x)))


def step(edge, x):
    edge, x = _v(edge), _v(x)
    if isinstance(x, Vector) or isinstance(edge, Vector):
        ev = edge.e if isinstance(edge, Vector) else [edge] * len(x.e)
        xv = x.e if isinstance(x, Vector) else [x] * len(edge.e)
        return Vector([f32(1.0) if b >= a else f32(0.0) for a, b in zip(ev, xv)])
    return f32(1.0) if x >= edge else f32(0.0)


def smoothstep(edge0, edge1, x):
    edge0, edge1, x = _v(edge0), _v(edge1), _v(x)
    t = clamp((x - edge0) / (edge1 - edge0), 0.0, 1.0)
    return t * t * (3.0 - 2.0 * t)


def fract(x):
    x = _v(x)
    return x - floor(x)


def log2(x):
    return log(_v(x)) / 0.6931471805599453


def length(x):
    return _v(x).norm()


def normalize(x):
    return _v(x).normalized()


def dot(a, b):
    return _v(a).dot(_v(b))


def cross(a, b):
    return _v(a).cross(_v(b))


def distance(a, b):
    return (_v(a) - _v(b)).norm()


def isnan(x):
    return bool(_np.isnan(f32(x)))


def isinf(x):
    return bool(_np.isinf(f32(x)))


def sign(x):
    x = _v(x)
    return f32(1.0) if x > 0 else (f32(-1.0) if x < 0 else f32(0.0))


def mod(x, y):
    x, y = _v(x), _v(y)
    return x - y * floor(x / y)


__all__ = [n for n in dir() if not n.startswith("_") and n not in ("Vector", "Matrix", "I32", "f32")]
