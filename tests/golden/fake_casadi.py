"""A sympy-backed stand-in for the tiny part of the `casadi` API that the reference uses.

Purpose (golden-vector generation only, never imported by the product): `casadi` is not installable here
(no network), so the reference's IPOPT solve cannot run.  But everything *around* the solve is plain
Python: quad_model.Quadrotor builds the dynamics/cost expressions, quad_OC.OCSys.ocSolver assembles the
NLP (objective J, constraints g, bounds, initial guess) and unpacks the solution, quad_policy.run_quad
computes the reward.  With this module installed as `sys.modules['casadi']` that code runs UNMODIFIED
from /root/reference; the expressions it builds are sympy expressions, so the NLP the reference defines
can be evaluated exactly (f, g, grad f, jac g) and a candidate solution can be certified as a KKT point
of the reference's own NLP.  `nlpsol` delegates the actual solve to a backend installed by the caller
(tests/golden/make_golden.py installs the CPU oracle and records the certificate).

Names leaked by `from casadi import *` that the reference relies on (numpy, np, atan, pi) are provided.
"""
from __future__ import annotations

import math

import numpy
import numpy as np
import sympy as sp
from math import atan, pi  # noqa: F401  (leaked names used by quad_model.py:820 / quad_moving.py:40 / main.py:46)


def _to_matrix(v):
    if isinstance(v, SX):
        return v.M
    if isinstance(v, sp.MatrixBase):
        return sp.Matrix(v)
    if isinstance(v, (list, tuple)):
        a = numpy.array(v, dtype=object)
        if a.ndim == 1:
            return sp.Matrix(len(a), 1, [sp.Float(x, 17) if isinstance(x, float) else sp.sympify(x) for x in a])
        return sp.Matrix(a.shape[0], a.shape[1],
                         [sp.Float(x, 17) if isinstance(x, float) else sp.sympify(x) for x in a.reshape(-1)])
    if isinstance(v, numpy.ndarray):
        if v.ndim == 0:
            return sp.Matrix([[_num(v.item())]])
        if v.ndim == 1:
            return sp.Matrix(v.shape[0], 1, [_num(x) for x in v])
        return sp.Matrix(v.shape[0], v.shape[1], [_num(x) for x in v.reshape(-1)])
    return sp.Matrix([[_num(v)]])


def _num(x):
    if isinstance(x, (float, numpy.floating)):
        return sp.Float(float(x), 17)
    if isinstance(x, (int, numpy.integer)):
        return sp.Integer(int(x))
    if isinstance(x, SX):
        assert x.M.shape == (1, 1)
        return x.M[0, 0]
    return sp.sympify(x)


class SX:
    """Dense symbolic matrix with casadi-like semantics (elementwise * and /, scalar broadcast)."""

    __array_ufunc__ = None       # make numpy defer to our reflected operators
    __array_priority__ = 1000

    def __init__(self, M):
        self.M = M if isinstance(M, sp.MatrixBase) else _to_matrix(M)

    # -- constructors ------------------------------------------------------------------------------
    @staticmethod
    def sym(name, n=1, m=1):
        if n == 1 and m == 1:
            return SX(sp.Matrix([[sp.Symbol(name, real=True)]]))
        return SX(sp.Matrix(n, m, [sp.Symbol(f"{name}_{i}", real=True) for i in range(n * m)]))

    # -- shape -------------------------------------------------------------------------------------
    def numel(self):
        return self.M.shape[0] * self.M.shape[1]

    @property
    def shape(self):
        return self.M.shape

    def __getitem__(self, idx):
        if isinstance(idx, tuple):
            r = self.M[idx]
        else:
            flat = list(self.M) if self.M.shape[1] == 1 or self.M.shape[0] == 1 else None
            if flat is None:
                raise IndexError("single index on a matrix")
            r = flat[idx]
        if isinstance(r, sp.MatrixBase):
            return SX(r)
        if isinstance(r, list):
            return SX(sp.Matrix(len(r), 1, r))
        return SX(sp.Matrix([[r]]))

    # -- arithmetic --------------------------------------------------------------------------------
    @staticmethod
    def _bin(a, b, op):
        A, B = _to_matrix(a), _to_matrix(b)
        if A.shape == (1, 1) and B.shape != (1, 1):
            return SX(B.applyfunc(lambda e: op(A[0, 0], e)))
        if B.shape == (1, 1) and A.shape != (1, 1):
            return SX(A.applyfunc(lambda e: op(e, B[0, 0])))
        if A.shape != B.shape:
            if A.shape == B.T.shape and 1 in A.shape:      # row/column vector mix (numpy 1-D vs column)
                B = B.T
            else:
                raise ValueError(f"shape mismatch {A.shape} vs {B.shape}")
        return SX(sp.Matrix(A.shape[0], A.shape[1], [op(x, y) for x, y in zip(A, B)]))

    def __add__(self, o): return SX._bin(self, o, lambda x, y: x + y)
    def __radd__(self, o): return SX._bin(o, self, lambda x, y: x + y)
    def __sub__(self, o): return SX._bin(self, o, lambda x, y: x - y)
    def __rsub__(self, o): return SX._bin(o, self, lambda x, y: x - y)
    def __mul__(self, o): return SX._bin(self, o, lambda x, y: x * y)
    def __rmul__(self, o): return SX._bin(o, self, lambda x, y: x * y)
    def __truediv__(self, o): return SX._bin(self, o, lambda x, y: x / y)
    def __rtruediv__(self, o): return SX._bin(o, self, lambda x, y: x / y)
    def __pow__(self, e): return SX(self.M.applyfunc(lambda x: x ** e))
    def __neg__(self): return SX(-self.M)

    # -- numeric access ----------------------------------------------------------------------------
    def full(self):
        return numpy.array(self.M.tolist(), dtype=float)

    def __float__(self):
        assert self.M.shape == (1, 1)
        return float(self.M[0, 0])

    def __repr__(self):
        return f"SX({self.M!r})"


MX = SX
DM = SX


def vertcat(*args):
    if len(args) == 0:
        return SX(sp.zeros(0, 1))
    mats = [_to_matrix(a) for a in args]
    return SX(sp.Matrix.vstack(*mats))


def horzcat(*args):
    mats = [_to_matrix(a) for a in args]
    return SX(sp.Matrix.hstack(*mats))


def vcat(lst):
    return vertcat(*lst)


def mtimes(a, b):
    return SX(_to_matrix(a) * _to_matrix(b))


def transpose(a):
    return SX(_to_matrix(a).T)


def inv(a):
    return SX(_to_matrix(a).inv())


def diag(a):
    M = _to_matrix(a)
    if M.shape[1] == 1:
        return SX(sp.diag(*list(M)))
    return SX(sp.Matrix([M[i, i] for i in range(M.shape[0])]))


def dot(a, b):
    A, B = _to_matrix(a), _to_matrix(b)
    if A.shape != B.shape:
        B = B.T
    return SX(sp.Matrix([[sum(x * y for x, y in zip(A, B))]]))


def trace(a):
    return SX(sp.Matrix([[_to_matrix(a).trace()]]))


def exp(x):
    if isinstance(x, SX):
        return SX(x.M.applyfunc(sp.exp))
    return math.exp(x)


def jacobian(expr, var):
    return SX(_to_matrix(expr).jacobian(list(_to_matrix(var))))


class Function:
    """casadi.Function(name, [inputs], [outputs]) with a single output; call = symbolic substitution."""

    def __init__(self, name, ins, outs, *_, **__):
        self.name = name
        self.ins = [_to_matrix(i) for i in ins]
        assert len(outs) == 1
        self.out = _to_matrix(outs[0])

    def __call__(self, *args):
        assert len(args) == len(self.ins), (self.name, len(args), len(self.ins))
        sub = {}
        for sym_m, a in zip(self.ins, args):
            A = _to_matrix(a)
            if A.shape != sym_m.shape:
                A = A.reshape(*sym_m.shape)
            for s, v in zip(sym_m, A):
                sub[s] = v
        return SX(self.out.xreplace(sub))


# ---------------------------------------------------------------------------------------------------
# nlpsol: the solve itself is delegated to a caller-installed backend
# ---------------------------------------------------------------------------------------------------
NLP_BACKEND = None      # callable(nlp: "RecordedNLP", x0, lbx, ubx, lbg, ubg) -> dict(x=..., f=..., lam_g=...)
LAST_NLP = None


class RecordedNLP:
    """The NLP exactly as quad_OC.py:171 hands it to casadi.nlpsol, with exact evaluators."""

    def __init__(self, prob):
        self.x = list(_to_matrix(prob["x"]))
        self.f = _to_matrix(prob["f"])[0, 0]
        self.g = list(_to_matrix(prob["g"]))
        self._fn = None
        self._grad = None
        self._jac = None

    def _lam(self, exprs):
        return sp.lambdify(self.x, exprs, modules="math", cse=True)

    def eval_fg(self, w):
        if self._fn is None:
            self._fn = self._lam([self.f] + self.g)
        out = self._fn(*[float(v) for v in w])
        return float(out[0]), numpy.array(out[1:], dtype=float)

    def eval_grad(self, w):
        if self._grad is None:
            # f is a sum of stage terms: differentiate term by term (each touches ~20 symbols)
            terms = sp.Add.make_args(self.f)
            g = {s: [] for s in self.x}
            for t in terms:
                for s in t.free_symbols:
                    if s in g:
                        g[s].append(sp.diff(t, s))
            self._grad = self._lam([sp.Add(*g[s]) for s in self.x])
        return numpy.array(self._grad(*[float(v) for v in w]), dtype=float)

    def eval_jac(self, w):
        if self._jac is None:
            idx = {s: i for i, s in enumerate(self.x)}
            rows, cols, ex = [], [], []
            for r, gi in enumerate(self.g):
                for s in gi.free_symbols:
                    if s in idx:
                        rows.append(r); cols.append(idx[s]); ex.append(sp.diff(gi, s))
            self._jac = (numpy.array(rows), numpy.array(cols), self._lam(ex))
        rows, cols, fn = self._jac
        J = numpy.zeros((len(self.g), len(self.x)))
        J[rows, cols] = fn(*[float(v) for v in w])
        return J


class _DM:
    def __init__(self, a):
        self.a = numpy.atleast_2d(numpy.asarray(a, dtype=float))
        if self.a.shape[0] == 1 and self.a.size > 1:
            self.a = self.a.T

    def full(self):
        return self.a


class _Solver:
    def __init__(self, nlp):
        self.nlp = nlp

    def __call__(self, x0=None, lbx=None, ubx=None, lbg=None, ubg=None, **_):
        if NLP_BACKEND is None:
            raise RuntimeError("fake_casadi.NLP_BACKEND is not installed")
        res = NLP_BACKEND(self.nlp, x0, lbx, ubx, lbg, ubg)
        return {"x": _DM(res["x"]), "f": _DM([[res["f"]]]), "lam_g": _DM(res["lam_g"])}


def nlpsol(name, plugin, prob, opts=None):
    global LAST_NLP
    assert plugin == "ipopt"
    LAST_NLP = RecordedNLP(prob)
    LAST_NLP.opts = dict(opts or {})
    return _Solver(LAST_NLP)
