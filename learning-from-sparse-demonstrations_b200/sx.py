"""Minimal symbolic layer with the CasADi names the reference's JinEnv / CPDP code uses.

The reference builds every model with ``casadi.SX`` (``/root/reference/JinEnv/JinEnv.py``,
``/root/reference/CPDP/CPDP.py:2``).  CasADi is not available here and, more to the point, the
B200 path does not evaluate expression graphs in a VM: expressions are lowered to C device
functions (``codegen.py``).  This module is therefore only a thin façade over sympy exposing
``SX.sym, vertcat, vcat, horzcat, mtimes, dot, jacobian, pinv, diag, transpose, trace, sin, cos,
fmax, DM, substitute`` with CasADi's conventions (everything is a dense 2-D matrix, a scalar is
1x1, ``*`` is element-wise with 1x1 broadcasting, ``@``/``mtimes`` is the matrix product).
"""
import numbers

import numpy as np
import sympy as sp

__all__ = ["SX", "DM", "vertcat", "vcat", "horzcat", "hcat", "mtimes", "dot", "jacobian", "pinv",
           "inv", "diag", "transpose", "trace", "sin", "cos", "tan", "exp", "sqrt", "fmax", "fmin",
           "substitute", "pi"]

pi = np.pi


def _to_matrix(v):
    """Anything -> sympy Matrix (2-D)."""
    if isinstance(v, SX):
        return v.m
    if isinstance(v, sp.MatrixBase):
        return sp.Matrix(v)
    if isinstance(v, np.ndarray):
        a = np.atleast_1d(v)
        if a.ndim == 1:
            a = a.reshape(-1, 1)
        return sp.Matrix(a.shape[0], a.shape[1], [_num(x) for x in a.flatten()])
    if isinstance(v, (list, tuple)):
        if len(v) and isinstance(v[0], (list, tuple)):
            return sp.Matrix([[_num(x) for x in row] for row in v])
        return sp.Matrix(len(v), 1, [_num(x) for x in v])
    return sp.Matrix(1, 1, [_num(v)])


def _num(x):
    if isinstance(x, SX):
        assert x.m.shape == (1, 1)
        return x.m[0, 0]
    if isinstance(x, sp.Basic):
        return x
    if isinstance(x, (bool, np.bool_)):
        return sp.Integer(int(x))
    if isinstance(x, (numbers.Integral, np.integer)):
        return sp.Integer(int(x))
    if isinstance(x, (numbers.Real, np.floating)):
        f = float(x)
        # keep small integers exact so that 0*expr and 1*expr simplify structurally
        if f == int(f) and abs(f) < 2 ** 31:
            return sp.Integer(int(f))
        return sp.Float(f, 17)
    raise TypeError("cannot convert %r to a symbolic scalar" % (x,))


def _bcast(a, b, op):
    A, B = _to_matrix(a), _to_matrix(b)
    if A.shape == B.shape:
        return SX(sp.Matrix(A.shape[0], A.shape[1], [op(x, y) for x, y in zip(A, B)]))
    if A.shape == (1, 1):
        s = A[0, 0]
        return SX(sp.Matrix(B.shape[0], B.shape[1], [op(s, y) for y in B]))
    if B.shape == (1, 1):
        s = B[0, 0]
        return SX(sp.Matrix(A.shape[0], A.shape[1], [op(x, s) for x in A]))
    raise ValueError("dimension mismatch %s vs %s" % (A.shape, B.shape))


class SX:
    """Dense symbolic matrix (sympy-backed) with CasADi ``SX`` semantics."""
    __array_ufunc__ = None  # make numpy defer to our reflected operators
    __array_priority__ = 1000

    def __init__(self, m=0):
        self.m = _to_matrix(m)

    # -- construction -----------------------------------------------------------------
    @staticmethod
    def sym(name, *dims):
        if len(dims) == 1 and isinstance(dims[0], (tuple, list)):
            dims = tuple(dims[0])
        if len(dims) == 0:
            return SX(sp.Matrix(1, 1, [sp.Symbol(name, real=True)]))
        nr = int(dims[0])
        nc = int(dims[1]) if len(dims) > 1 else 1
        if nr * nc == 1:
            return SX(sp.Matrix(1, 1, [sp.Symbol(name, real=True)]))
        # CasADi names entries name_0 ... in column-major order
        M = sp.zeros(nr, nc)
        k = 0
        for j in range(nc):
            for i in range(nr):
                M[i, j] = sp.Symbol("%s_%d" % (name, k), real=True)
                k += 1
        return SX(M)

    @staticmethod
    def zeros(nr, nc=1):
        return SX(sp.zeros(nr, nc))

    @staticmethod
    def eye(n):
        return SX(sp.eye(n))

    # -- shape ------------------------------------------------------------------------
    def numel(self):
        return self.m.shape[0] * self.m.shape[1]

    def size(self, axis=None):
        return self.m.shape if axis is None else self.m.shape[axis - 1]

    def size1(self):
        return self.m.shape[0]

    def size2(self):
        return self.m.shape[1]

    @property
    def shape(self):
        return self.m.shape

    @property
    def T(self):
        return SX(self.m.T)

    def __len__(self):
        return self.m.shape[0]

    def __iter__(self):
        for i in range(self.numel()):
            yield self[i]

    def __getitem__(self, idx):
        if isinstance(idx, tuple):
            sub = self.m[idx]
            return SX(sub if isinstance(sub, sp.MatrixBase) else sp.Matrix(1, 1, [sub]))
        # linear (column-major) indexing like CasADi; vectors are the only use in the reference
        flat = list(self.m.T) if self.m.shape[1] > 1 else list(self.m)
        if isinstance(idx, slice):
            sel = flat[idx]
            return SX(sp.Matrix(len(sel), 1, sel))
        return SX(sp.Matrix(1, 1, [flat[idx]]))

    # -- arithmetic -------------------------------------------------------------------
    def __add__(self, o): return _bcast(self, o, lambda x, y: x + y)
    def __radd__(self, o): return _bcast(o, self, lambda x, y: x + y)
    def __sub__(self, o): return _bcast(self, o, lambda x, y: x - y)
    def __rsub__(self, o): return _bcast(o, self, lambda x, y: x - y)
    def __mul__(self, o): return _bcast(self, o, lambda x, y: x * y)
    def __rmul__(self, o): return _bcast(o, self, lambda x, y: x * y)
    def __truediv__(self, o): return _bcast(self, o, lambda x, y: x / y)
    def __rtruediv__(self, o): return _bcast(o, self, lambda x, y: x / y)
    def __pow__(self, o): return _bcast(self, o, lambda x, y: x ** y)
    def __rpow__(self, o): return _bcast(o, self, lambda x, y: x ** y)
    def __neg__(self): return SX(-self.m)
    def __pos__(self): return self
    def __matmul__(self, o): return mtimes(self, o)
    def __rmatmul__(self, o): return mtimes(o, self)

    def __repr__(self):
        return "SX(%s)" % (self.m.tolist() if self.numel() > 1 else self.m[0, 0],)

    # scalar access for host-side numeric use
    def __float__(self):
        assert self.m.shape == (1, 1)
        return float(self.m[0, 0])

    def free_symbols(self):
        return self.m.free_symbols


def DM(v):
    """Numeric matrix; represented as SX with numeric entries (the reference only uses DM as the
    substitution value in ``CPDP.py:107-108``)."""
    return SX(v)


def vertcat(*args):
    mats = [_to_matrix(a) for a in args]
    mats = [m for m in mats if m.shape[0] * m.shape[1] > 0]
    if not mats:
        return SX(sp.zeros(0, 1))
    return SX(sp.Matrix.vstack(*mats))


def vcat(lst):
    return vertcat(*lst)


def horzcat(*args):
    mats = [_to_matrix(a) for a in args]
    if not mats:
        return SX(sp.zeros(1, 0))
    return SX(sp.Matrix.hstack(*mats))


def hcat(lst):
    return horzcat(*lst)


def mtimes(*args):
    if len(args) == 1 and isinstance(args[0], (list, tuple)):
        args = tuple(args[0])
    out = _to_matrix(args[0])
    for a in args[1:]:
        b = _to_matrix(a)
        if out.shape == (1, 1) and b.shape != (1, 1) and out.shape[1] != b.shape[0]:
            out = out[0, 0] * b
        elif b.shape == (1, 1) and out.shape[1] != 1:
            out = out * b[0, 0]
        else:
            out = out * b
    return SX(out)


def dot(a, b):
    A, B = _to_matrix(a), _to_matrix(b)
    assert A.shape == B.shape
    return SX(sp.Matrix(1, 1, [sum(x * y for x, y in zip(A, B))]))


def jacobian(expr, wrt):
    E, Wm = _to_matrix(expr), _to_matrix(wrt)
    ev = list(E) if E.shape[1] == 1 else list(E.T)   # column-major vec, as CasADi does
    wv = list(Wm) if Wm.shape[1] == 1 else list(Wm.T)
    J = sp.zeros(len(ev), len(wv))
    for i, e in enumerate(ev):
        fs = e.free_symbols if isinstance(e, sp.Basic) else set()
        for j, w in enumerate(wv):
            if w in fs:
                J[i, j] = sp.diff(e, w)
    return SX(J)


def inv(a):
    A = _to_matrix(a)
    n = A.shape[0]
    assert A.shape == (n, n)
    if n == 1:
        return SX(sp.Matrix(1, 1, [1 / A[0, 0]]))
    if A.is_diagonal():
        return SX(sp.diag(*[1 / A[i, i] for i in range(n)]))
    if n == 2:
        det = A[0, 0] * A[1, 1] - A[0, 1] * A[1, 0]
        return SX(sp.Matrix([[A[1, 1] / det, -A[0, 1] / det], [-A[1, 0] / det, A[0, 0] / det]]))
    return SX(A.adjugate() / A.det())


def pinv(a):
    """The reference only takes ``pinv`` of square, generically non-singular matrices (the arm's
    2x2 mass matrix ``JinEnv.py:236``, diagonal inertias ``:749,1321`` and ``Huu`` ``CPDP.py:262``),
    where the pseudo-inverse is the inverse."""
    A = _to_matrix(a)
    assert A.shape[0] == A.shape[1], "pinv: only the square case occurs on the CPDP path"
    return inv(a)


def diag(a):
    A = _to_matrix(a)
    if A.shape[1] == 1 or A.shape[0] == 1:
        return SX(sp.diag(*list(A)))
    return SX(sp.Matrix([A[i, i] for i in range(min(A.shape))]))


def transpose(a):
    return SX(_to_matrix(a).T)


def trace(a):
    A = _to_matrix(a)
    return SX(sp.Matrix(1, 1, [sum(A[i, i] for i in range(A.shape[0]))]))


def _ew(fn):
    def g(a):
        if isinstance(a, (numbers.Real, np.floating)):
            return float(getattr(np, fn.__name__)(a))
        A = _to_matrix(a)
        return SX(sp.Matrix(A.shape[0], A.shape[1], [fn(x) for x in A]))
    return g


sin = _ew(sp.sin)
cos = _ew(sp.cos)
tan = _ew(sp.tan)
exp = _ew(sp.exp)
sqrt = _ew(sp.sqrt)


def fmax(a, b):
    if not isinstance(a, SX) and not isinstance(b, SX):
        return max(float(a), float(b))
    return _bcast(a, b, lambda x, y: sp.Max(x, y))


def fmin(a, b):
    if not isinstance(a, SX) and not isinstance(b, SX):
        return min(float(a), float(b))
    return _bcast(a, b, lambda x, y: sp.Min(x, y))


def substitute(expr, var, val):
    E, V, W = _to_matrix(expr), _to_matrix(var), _to_matrix(val)
    assert V.shape == W.shape or V.shape[0] * V.shape[1] == W.shape[0] * W.shape[1]
    return SX(E.subs(dict(zip(list(V), list(W)))))
