"""ORACLE (test infrastructure only -- never imported by the product path).

CPU restatement of the reference's forward-mode dual numbers,
`/root/reference/src/adiff.jl` (AD4SM.jl v0.0.7).  Every operator follows the
reference's formula *and association order* so that rounding matches wherever
the arithmetic is IEEE-deterministic.

parity status: PARITY UNPINNED BY REFERENCE TESTS -- the reference ships no
tests/golden vectors and Julia is not available here.  The restatement is
pinned instead by analytic known-answer tests, sympy symbolic Hessians and
mpmath finite differences (tests/test_oracle_*.py).

Layout (adiff.jl:29-42):
  Grad{N}   -> 1-D float64 ndarray of length N
  D1{N}     -> D1(v, g)
  D2{N,M}   -> D2(v, g, h) with h the packed upper triangle, column-major:
               (i<=j), 1-based  ->  (j-1)*j/2 + i     (adiff.jl:26,84)
"""
from __future__ import annotations

import math
from functools import lru_cache

import numpy as np


# ----------------------------------------------------------------------------
# packed upper-triangular helpers (adiff.jl:26 `tupfy2`, :84 getindex, :94 outer)
# ----------------------------------------------------------------------------
@lru_cache(maxsize=None)
def _tri_idx(n: int):
    """(I, J) 0-based index vectors enumerating `for j in 1:N for i in 1:j`."""
    ii, jj = [], []
    for j in range(n):
        for i in range(j + 1):
            ii.append(i)
            jj.append(j)
    return np.asarray(ii, dtype=np.intp), np.asarray(jj, dtype=np.intp)


def packed_index(i: int, j: int) -> int:
    """0-based packed slot of the 0-based pair (i, j); adiff.jl:84."""
    if i > j:
        i, j = j, i
    return j * (j + 1) // 2 + i


def outer(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """`*(x::Grad{N}, y::Grad{N})` -> Grad{N(N+1)/2}: x[i]*y[j], i<=j  (adiff.jl:94)."""
    ii, jj = _tri_idx(len(x))
    return x[ii] * y[jj]


def _nm(n: int) -> int:
    return (n + 1) * n // 2


def _pow(v: float, n):
    """Scalar `v^n` as Julia evaluates it for Float64 base (libm pow semantics)."""
    if isinstance(n, (int, np.integer)):
        n = int(n)
        if n >= 0:
            # Julia: power-by-squaring with compensation == correctly rounded for small n
            return float(v) ** n
        return (1.0 / float(v)) ** (-n) if v != 0 else math.inf
    return math.pow(v, n) if v >= 0 or float(n).is_integer() else math.nan


# ----------------------------------------------------------------------------
# D1  (adiff.jl:33-36, 110-129)
# ----------------------------------------------------------------------------
class D1:
    __slots__ = ("v", "g")
    __array_priority__ = 1000

    def __init__(self, v, g):
        self.v = float(v)
        self.g = np.asarray(g, dtype=np.float64)

    @property
    def N(self):
        return len(self.g)

    # promotion of reals (adiff.jl:55,80)
    def _lift(self, o):
        if isinstance(o, D1):
            return o
        if isinstance(o, D2):
            raise TypeError("D1 (op) D2 is promoted to D2 in the reference; not needed on the hot path")
        return D1(float(o), np.zeros(self.N))

    def __add__(self, o):
        o = self._lift(o)
        return D1(self.v + o.v, self.g + o.g)

    def __radd__(self, o):
        o = self._lift(o)
        return D1(o.v + self.v, o.g + self.g)

    def __sub__(self, o):
        o = self._lift(o)
        return D1(self.v - o.v, self.g - o.g)

    def __rsub__(self, o):
        o = self._lift(o)
        return D1(o.v - self.v, o.g - self.g)

    def __neg__(self):
        return D1(-self.v, -self.g)

    def __mul__(self, o):  # adiff.jl:114
        o = self._lift(o)
        return D1(self.v * o.v, self.v * o.g + o.v * self.g)

    def __rmul__(self, o):
        o = self._lift(o)
        return D1(o.v * self.v, o.v * self.g + self.v * o.g)

    def inv(self):  # adiff.jl:115
        return D1(1 / self.v, (-1 / _pow(self.v, 2)) * self.g)

    def __truediv__(self, o):  # adiff.jl:116
        return self * self._lift(o).inv()

    def __rtruediv__(self, o):
        return self._lift(o) * self.inv()

    def __pow__(self, n):  # adiff.jl:117-118
        if n == 0:
            return D1(1.0, np.zeros(self.N))
        if n == 1:
            return self
        return D1(_pow(self.v, n), (n * _pow(self.v, n - 1)) * self.g)

    def log(self):  # adiff.jl:119
        return D1(math.log(self.v), self.g / self.v)

    def sqrt(self):  # adiff.jl:126
        return self ** 0.5

    # comparisons on the value (adiff.jl:97-108)
    def __lt__(self, o):
        return self.v < val(o)

    def __le__(self, o):
        return self.v <= val(o)

    def __gt__(self, o):
        return self.v > val(o)

    def __ge__(self, o):
        return self.v >= val(o)

    def __repr__(self):
        return f"D1({self.v}, N={self.N})"


# ----------------------------------------------------------------------------
# D2  (adiff.jl:38-42, 131-150)
# ----------------------------------------------------------------------------
class D2:
    __slots__ = ("v", "g", "h")
    __array_priority__ = 1000

    def __init__(self, v, g, h=None):
        self.v = float(v)
        self.g = np.asarray(g, dtype=np.float64)
        if h is None:  # adiff.jl:56
            h = np.zeros(_nm(len(self.g)))
        self.h = np.asarray(h, dtype=np.float64)

    @property
    def N(self):
        return len(self.g)

    def _lift(self, o):  # adiff.jl:54,78-79
        if isinstance(o, D2):
            return o
        if isinstance(o, D1):
            return D2(o.v, o.g, np.zeros(_nm(o.N)))
        return D2(float(o), np.zeros(self.N), np.zeros(len(self.h)))

    def __add__(self, o):  # adiff.jl:132
        o = self._lift(o)
        return D2(self.v + o.v, self.g + o.g, self.h + o.h)

    def __radd__(self, o):
        o = self._lift(o)
        return D2(o.v + self.v, o.g + self.g, o.h + self.h)

    def __sub__(self, o):  # adiff.jl:133
        o = self._lift(o)
        return D2(self.v - o.v, self.g - o.g, self.h - o.h)

    def __rsub__(self, o):
        o = self._lift(o)
        return D2(o.v - self.v, o.g - self.g, o.h - self.h)

    def __neg__(self):  # adiff.jl:134
        return D2(-self.v, -self.g, -self.h)

    @staticmethod
    def _mul(x, y):  # adiff.jl:135  (association order kept)
        return D2(
            x.v * y.v,
            x.v * y.g + y.v * x.g,
            x.v * y.h + y.v * x.h + outer(x.g, y.g) + outer(y.g, x.g),
        )

    def __mul__(self, o):
        return D2._mul(self, self._lift(o))

    def __rmul__(self, o):
        return D2._mul(self._lift(o), self)

    def inv(self):  # adiff.jl:136
        x = self
        return D2(
            1 / x.v,
            (-1 / _pow(x.v, 2)) * x.g,
            (2 / _pow(x.v, 3)) * outer(x.g, x.g) - (1 / _pow(x.v, 2)) * x.h,
        )

    def __truediv__(self, o):  # adiff.jl:137
        return D2._mul(self, self._lift(o).inv())

    def __rtruediv__(self, o):
        return D2._mul(self._lift(o), self.inv())

    def __pow__(self, n):  # adiff.jl:138-139
        x = self
        if n == 0:
            return D2(1.0, np.zeros(x.N), np.zeros(len(x.h)))
        if n == 1:
            return x
        return D2(
            _pow(x.v, n),
            (n * _pow(x.v, n - 1)) * x.g,
            (n * (n - 1) * _pow(x.v, n - 2)) * outer(x.g, x.g) + (n * _pow(x.v, n - 1)) * x.h,
        )

    def log(self):  # adiff.jl:140
        x = self
        return D2(math.log(x.v), x.g / x.v, -outer(x.g, x.g) / _pow(x.v, 2) + x.h / x.v)

    def exp(self):  # adiff.jl:141
        x = self
        e = math.exp(x.v)
        return D2(e, e * x.g, e * outer(x.g, x.g) + e * x.h)

    def sqrt(self):  # adiff.jl:147
        return self ** 0.5

    def __abs__(self):  # adiff.jl:148
        return self if self.v >= 0 else -self

    def __lt__(self, o):
        return self.v < val(o)

    def __le__(self, o):
        return self.v <= val(o)

    def __gt__(self, o):
        return self.v > val(o)

    def __ge__(self, o):
        return self.v >= val(o)

    def __repr__(self):
        return f"D2({self.v}, N={self.N})"


# ----------------------------------------------------------------------------
# seeding and extraction  (adiff.jl:58-68, 71-74, 153-160)
# ----------------------------------------------------------------------------
def seed_D2(x) -> np.ndarray:
    """`D2(x::AbstractArray)`: variable k = column-major linear index (adiff.jl:58-62).
    Returns an object array of the same shape."""
    x = np.asarray(x, dtype=np.float64)
    flat = x.reshape(-1, order="F")
    n = flat.size
    eye = np.eye(n)
    out = np.empty(n, dtype=object)
    for k in range(n):
        out[k] = D2(flat[k], eye[k].copy(), np.zeros(_nm(n)))
    return out.reshape(x.shape, order="F")


def seed_D1(x) -> np.ndarray:
    """`D1(x::AbstractArray)` (adiff.jl:64-68)."""
    x = np.asarray(x, dtype=np.float64)
    flat = x.reshape(-1, order="F")
    n = flat.size
    eye = np.eye(n)
    out = np.empty(n, dtype=object)
    for k in range(n):
        out[k] = D1(flat[k], eye[k].copy())
    return out.reshape(x.shape, order="F")


def to_D1(x):
    """`D1(x::D2)` drops the second order (adiff.jl:71)."""
    if isinstance(x, np.ndarray):
        out = np.empty(x.shape, dtype=object)
        for idx in np.ndindex(x.shape):
            out[idx] = to_D1(x[idx])
        return out
    return D1(x.v, x.g.copy())


def val(x):
    """`val` (adiff.jl:155) and `convert(::Type{<:Real}, x)` (adiff.jl:73-74)."""
    if isinstance(x, (D1, D2)):
        return x.v
    if isinstance(x, np.ndarray) and x.dtype == object:
        out = np.empty(x.shape)
        for idx in np.ndindex(x.shape):
            out[idx] = val(x[idx])
        return out
    return x


def grad(x) -> np.ndarray:
    """adiff.jl:157-158"""
    return np.array(x.g, dtype=np.float64)


def hess(x) -> np.ndarray:
    """Full symmetric N x N matrix from the packed triangle (adiff.jl:159-160)."""
    n = x.N
    if isinstance(x, D1):
        return np.zeros((n, n))
    ii, jj = _tri_idx(n)
    H = np.empty((n, n))
    H[ii, jj] = x.h
    H[jj, ii] = x.h
    return H


def dlog(x):
    return x.log() if isinstance(x, (D1, D2)) else math.log(x)


def dsqrt(x):
    return x.sqrt() if isinstance(x, (D1, D2)) else math.sqrt(x)
