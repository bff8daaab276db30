"""Run-time support for the translated reference routines (TEST INFRASTRUCTURE):
Fortran arrays with arbitrary lower bounds, inclusive DO ranges, Fortran integer
division and the intrinsics the MOLOCH path uses.  See fortran_subset.py."""
from __future__ import annotations

import math

import numpy as np


class FArr:
    """A Fortran array a(l1:u1, l2:u2, ...) -- first index fastest -- backed by a
    C-ordered NumPy array of reversed shape, i.e. a(j,i,k) is self.a[k-lk, i-li, j-lj]
    (the layout of the oracle's global arrays).  Every access is bounds checked."""
    __slots__ = ("a", "lb", "nd")

    def __init__(self, a: np.ndarray, lb):
        self.a, self.lb, self.nd = a, tuple(int(x) for x in lb), a.ndim
        assert len(self.lb) == a.ndim

    @classmethod
    def alloc(cls, bounds, kind="float"):
        shape = tuple(int(hi) - int(lo) + 1 for lo, hi in reversed(bounds))
        return cls(np.zeros(shape, dtype=np.int64 if kind == "int" else np.float64), [lo for lo, _ in bounds])

    def bounds(self):
        return [(self.lb[d], self.lb[d] + self.a.shape[self.nd - 1 - d] - 1) for d in range(self.nd)]

    def _off(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        if len(idx) != self.nd:
            raise IndexError(f"rank mismatch: {len(idx)} subscripts for a rank-{self.nd} array")
        out = []
        for d in range(self.nd - 1, -1, -1):
            x, n = idx[d], self.a.shape[self.nd - 1 - d]
            if isinstance(x, slice):
                lo = 0 if x.start is None else x.start - self.lb[d]
                hi = n if x.stop is None else x.stop - self.lb[d] + 1       # Fortran sections are inclusive
                if lo < 0 or hi > n:
                    raise IndexError(f"section {x} outside bounds {self.bounds()[d]}")
                out.append(slice(lo, hi))
            else:
                o = int(x) - self.lb[d]
                if o < 0 or o >= n:
                    raise IndexError(f"subscript {d + 1} = {x} outside bounds {self.bounds()[d]}")
                out.append(o)
        return tuple(out)

    def __getitem__(self, idx):
        if idx is Ellipsis:
            return self.a
        o = self._off(idx)
        v = self.a[o]
        if not isinstance(v, np.ndarray):
            return v.item()
        return FArr(v, [1] * v.ndim)       # an array section: a view, lower bounds 1 (assumed-shape dummy semantics)

    # whole-array arithmetic (sections in expressions): element-wise on the storage
    def __array__(self, dtype=None, copy=None):
        return self.a if dtype is None else self.a.astype(dtype)

    def __mul__(self, o): return self.a * (o.a if isinstance(o, FArr) else o)
    def __rmul__(self, o): return (o.a if isinstance(o, FArr) else o) * self.a
    def __add__(self, o): return self.a + (o.a if isinstance(o, FArr) else o)
    def __radd__(self, o): return (o.a if isinstance(o, FArr) else o) + self.a
    def __sub__(self, o): return self.a - (o.a if isinstance(o, FArr) else o)
    def __rsub__(self, o): return (o.a if isinstance(o, FArr) else o) - self.a
    def __truediv__(self, o): return self.a / (o.a if isinstance(o, FArr) else o)
    def __neg__(self): return -self.a

    def __setitem__(self, idx, v):
        if isinstance(v, FArr):
            v = v.a
        if idx is Ellipsis:
            self.a[...] = v
            return
        self.a[self._off(idx)] = v

    def __call__(self, *idx):      # an array reference the translator took for a function call
        return self[idx if len(idx) > 1 else idx[0]]

    def view_last(self, n):
        """a(:,:,:,n) as a rank-3 array sharing memory (assignpnt(a,ptr,n))."""
        return FArr(self.a[n - self.lb[-1]], self.lb[:-1])


def _frange(a, b, s=1):
    a, b, s = int(a), int(b), int(s)
    return range(a, b + (1 if s > 0 else -1), s)


def _div(a, b):
    if isinstance(a, int) and isinstance(b, int) and not isinstance(a, bool):
        q = abs(a) // abs(b)
        return q if (a >= 0) == (b >= 0) else -q
    return a / b


def _pow(a, b):
    """x**n with an integer n: the multiplication chain compilers emit (x*x, (x*x)*x, (x*x)*(x*x));
    otherwise the C library's pow."""
    if isinstance(b, int) and not isinstance(b, bool):
        if isinstance(a, int):
            return a ** b if b >= 0 else 0
        n, r = abs(b), 1.0
        if n == 1:
            r = a
        elif n == 2:
            r = a * a
        elif n == 3:
            r = (a * a) * a
        elif n == 4:
            r = (a * a) * (a * a)
        elif n > 4:
            r, base = 1.0, a
            while n:
                if n & 1:
                    r = r * base
                base = base * base
                n >>= 1
        return r if b >= 0 else 1.0 / r
    return math.pow(a, b)


def _elemental(f):
    """An elemental function applied to whole arrays: element by element, in storage order."""
    def g(*args):
        arrs = [x for x in args if isinstance(x, (FArr, np.ndarray))]
        if not arrs:
            return f(*args)
        shape = (arrs[0].a if isinstance(arrs[0], FArr) else arrs[0]).shape
        flat = [(x.a if isinstance(x, FArr) else x).reshape(-1) if isinstance(x, (FArr, np.ndarray)) else None
                for x in args]
        out = np.empty(int(np.prod(shape)))
        for e in range(out.size):
            out[e] = f(*[(fl[e].item() if fl is not None else x) for fl, x in zip(flat, args)])
        return out.reshape(shape)
    return g


def _vec(fn):
    def g(x, *rest):
        if isinstance(x, FArr):
            x = x.a
        if isinstance(x, np.ndarray):
            return np.array([fn(v) for v in x.reshape(-1).tolist()]).reshape(x.shape)
        return fn(x, *rest)
    return g


def _size(a, dim=None):
    if dim is None:
        return int(a.a.size)
    return int(a.a.shape[a.nd - dim])


def _r4(x):
    return float(np.float32(x))


def _alloc(bounds, kind):
    return FArr.alloc(bounds, kind)


def _farr(vals, lo):
    allint = all(isinstance(v, int) for v in vals)
    return FArr(np.array(vals, dtype=np.int64 if allint else np.float64), [lo])


def _getmem(*b):
    """getmem(a, l1,u1, l2,u2, ...): a zero-initialised array with those bounds (Share/mod_memutil.F90)."""
    return FArr.alloc([(b[k], b[k + 1]) for k in range(0, len(b) - 1, 2)])


def _assignpnt(a, n=None):
    return a if n is None else a.view_last(n)


def _mod(a, b):
    if isinstance(a, int) and isinstance(b, int):
        return int(math.fmod(a, b))
    return math.fmod(a, b)


def _sign(a, b):
    return math.copysign(abs(a), b) if isinstance(a, float) or isinstance(b, float) else (abs(a) if b >= 0 else -abs(a))


def _real(x, kind=None):
    if kind == 4:
        return _r4(x)
    return float(x)


def _int(x, kind=None):
    return int(x)        # truncation towards zero, as Fortran int()


def _nint(x, kind=None):
    return int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))


INTRINSICS = {
    "_frange": _frange, "_div": _div, "_r4": _r4, "_alloc": _alloc, "_assignpnt": _assignpnt, "_farr": _farr, "_getmem": _getmem,
    "_pow": _pow, "_elemental": _elemental, "size": _size,
    "max": max, "min": min, "abs": abs, "sqrt": _vec(math.sqrt), "exp": _vec(math.exp), "log": _vec(math.log),
    "sin": _vec(math.sin), "cos": _vec(math.cos), "tan": math.tan, "atan": math.atan, "mod": _mod, "sign": _sign, "real": _real, "int": _int,
    "nint": _nint, "dble": float, "null": lambda: None,
    "rkx": 8, "rk8": 8, "rk4": 4, "rk16": 16, "ik4": 4, "ik8": 8, "wrkp": 8,
}
