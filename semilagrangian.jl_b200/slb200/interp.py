"""Interpolation types -- host mirror of the reference's AbstractInterpolation subtypes.

Reference (paths relative to the reference root): src/interpolation.jl:18-46 (abstract
type, `tabfct`, get_order, sol), src/lagrange.jl:25-72, src/spline.jl:54-97 +
src/bspline.jl:9-16 + src/bsplinelu.jl:253-270 + src/bsplinefft.jl:25-45, src/hermite.jl:4-133.

The weight polynomials `tabfct` are built in exact rational arithmetic and rounded once to
Float64, exactly as the reference does (Rational{BigInt} -> Float64); the device object
(slb_interp) receives the rounded table.  The constructions below are closed forms, not the
reference's recursive polynomial algebra:
  * Lagrange:  l_j(x) = W(x) / ((x - x_j) W'(x_j)),  W(x) = prod_l (x - x_l)   (synthetic division)
  * B-spline:  B_p(x) = 1/p! sum_k (-1)^k C(p+1,k) (x-k)_+^p  (truncated powers, binomial expansion)
  * Hermite:   two-point-family Hermite basis H_i, K_i with finite-difference derivative
               weights b+/- (same definition as src/hermite.jl, evaluated directly).
"""
from fractions import Fraction
from math import comb, factorial

import numpy as np

from . import _lib

LAGRANGE, BSPLINE_LU, BSPLINE_FFT, HERMITE = 0, 1, 2, 3


# ---- small exact polynomial helpers (ascending Fraction coefficient lists) -------------
def _pmul(a, b):
    r = [Fraction(0)] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                r[i + j] += x * y
    return r


def _padd(a, b):
    n = max(len(a), len(b))
    return [(a[i] if i < len(a) else 0) + (b[i] if i < len(b) else 0) for i in range(n)]


def _pscale(a, s):
    return [x * s for x in a]


def _peval(a, x):
    r = Fraction(0)
    for c in reversed(a):
        r = r * x + c
    return r


def _from_roots(roots):
    p = [Fraction(1)]
    for r in roots:
        p = _pmul(p, [Fraction(-r), Fraction(1)])
    return p


def _deflate(p, r):
    """p(x) / (x - r) for a root r (synthetic division), ascending coefficients."""
    n = len(p) - 1
    q = [Fraction(0)] * n
    carry = Fraction(0)
    for k in range(n, 0, -1):
        carry = p[k] + carry * r
        q[k - 1] = carry
    return q


def lagrange_table(order):
    """Exact weight polynomials of Lagrange(order): nodes origin..origin+order,
    origin = -div(order, 2)  (src/lagrange.jl:25-35, :62-70)."""
    origin = -(order // 2)
    nodes = [Fraction(origin + l) for l in range(order + 1)]
    W = _from_roots(nodes)
    tab = []
    for j, xj in enumerate(nodes):
        q = _deflate(W, xj)
        tab.append(_pscale(q, 1 / _peval(q, xj)))
    return tab


def bspline_piece(p, i):
    """Polynomial of the cardinal B-spline of degree p on [i, i+1) (src/spline.jl:67-88),
    from the truncated-power form."""
    acc = [Fraction(0)] * (p + 1)
    for k in range(0, i + 1):
        sgn = -1 if k % 2 else 1
        ck = Fraction(sgn * comb(p + 1, k), factorial(p))
        # (x - k)^p = sum_m C(p,m) x^m (-k)^(p-m)
        for m in range(p + 1):
            acc[m] += ck * comb(p, m) * Fraction(-k) ** (p - m)
    return acc


def bspline_value(p, x):
    x = Fraction(x)
    i = x.numerator // x.denominator
    if 0 <= i <= p:
        return _peval(bspline_piece(p, i), x)
    return Fraction(0)


def bspline_table(order):
    """tabfct[j](t) = B(order - j + t), j = 0..order (src/bsplinelu.jl:262-263)."""
    tab = []
    for j in range(order + 1):
        piece = bspline_piece(order, order - j)
        shift = Fraction(order - j)
        # compose piece(t + shift)
        out = [Fraction(0)] * (order + 1)
        for m, c in enumerate(piece):
            if c:
                for q in range(m + 1):
                    out[q] += c * comb(m, q) * shift ** (m - q)
        tab.append(out)
    return tab


def bspline_nodes(order):
    """B(1..order): the circulant's non-zero entries (src/bsplinelu.jl:264, src/bsplinefft.jl:35)."""
    return [bspline_value(order, i) for i in range(1, order + 1)]


def get_kl_ku(order):
    """src/bspline.jl:12-16"""
    ku = order // 2
    return order - 1 - ku, ku


def hermite_table(order, flbis=False):
    """Weight polynomials of Hermite(order) (src/hermite.jl:99-132): Hermite interpolation on
    the nodes -d..d+1 (d = div(ord,2)) whose derivatives are replaced by one-sided
    finite-difference formulas b+ (nodes <= 0) / b- (nodes >= 1)."""
    if flbis:
        if order % 4 != 3:
            raise ValueError(f"order={order} modulo 4 must equal to 3")
        ord_ = order // 2
    else:
        if order % 4 != 1:
            raise ValueError(f"order={order} modulo 4 must equal to 1")
        ord_ = order // 2 + 1
    d = ord_ // 2
    nodes = list(range(-d, d + 2))
    W = _from_roots(nodes)

    def ell(i):
        q = _deflate(W, Fraction(i))
        return _pscale(q, 1 / _peval(q, Fraction(i)))

    def ellprime_at_node(i):
        return sum(Fraction(1, i - j) for j in nodes if j != i)

    rplus, splus = (-d - 1, d) if flbis else (-d, d + 1)
    rminus, sminus = -splus, -rplus

    def fd_weights(lo, hi):
        """weights of the first-derivative finite-difference formula at 0 on nodes lo..hi"""
        w = {}
        for i in range(lo, hi + 1):
            if i == 0:
                continue
            num = Fraction(1)
            for j in range(lo, hi + 1):
                if j != 0 and j != i:
                    num *= -j
            den = Fraction(1)
            for j in range(lo, hi + 1):
                if j != i:
                    den *= i - j
            w[i] = num / den
        w[0] = -sum(w.values())
        return w

    bplus = fd_weights(rplus, splus)
    bminus = {-k: -v for k, v in fd_weights(-sminus, -rminus).items()}
    bminus[0] = -sum(v for k, v in bminus.items() if k != 0)

    centre = (order + 1) // 2  # 1-based slot of node 0 in the reference's tabfct
    tab = [[Fraction(0)] for _ in range(order + 1)]
    for i in nodes:
        li = ell(i)
        l2 = _pmul(li, li)
        xm = [Fraction(-i), Fraction(1)]
        Hi = _pmul(l2, _padd([Fraction(1)], _pscale(xm, -2 * ellprime_at_node(i))))
        Ki = _pmul(l2, xm)
        tab[centre + i - 1] = _padd(tab[centre + i - 1], Hi)
        bw = bplus if i <= 0 else bminus
        for k, wk in bw.items():
            tab[centre + i + k - 1] = _padd(tab[centre + i + k - 1], _pscale(Ki, wk))
    return tab


def _to_float_table(tab):
    nc = max(len(p) for p in tab)
    while nc > 1 and all((p[nc - 1] if nc - 1 < len(p) else 0) == 0 for p in tab):
        nc -= 1
    return np.array([[float(p[k]) if k < len(p) else 0.0 for k in range(nc)] for p in tab], dtype=np.float64)


# ---- the user-facing types -------------------------------------------------------------
class AbstractInterpolation:
    """AbstractInterpolation{T,edge,order} (src/interpolation.jl:18): CircEdge, T = Float64.
    `tabfct` is the (order+1) x ncoef Float64 coefficient table (row j = tabfct[j+1])."""

    kind = None

    def __init__(self, order, n=0):
        self.order = int(order)
        self.n = int(n)
        self.nodes = None
        self._handles = {}

    def get_order(self):
        return self.order

    def handle(self, ctx, n):
        """device object for lines of length n (created lazily, cached per context)."""
        key = (ctx.serial, int(n) if self.kind in (BSPLINE_LU, BSPLINE_FFT) else 0)
        h = self._handles.get(key)
        if h is None:
            if self.kind in (BSPLINE_LU, BSPLINE_FFT) and int(n) != self.n:
                raise ValueError(f"{type(self).__name__} built for n={self.n}, line length is {n}")
            import ctypes as C

            h = C.c_void_p()
            _lib.check(
                _lib.lib().slb_interp_create(
                    ctx.h, self.kind, self.order, int(n), _lib.dptr(self.tabfct), self.tabfct.shape[1],
                    _lib.dptr(self.nodes) if self.nodes is not None else None, C.byref(h),
                )
            )
            self._handles[key] = h

            def release(key=key, h=h):  # the context is closing: its device objects go with it
                if self._handles.pop(key, None) is not None:
                    _lib.lib().slb_interp_destroy(h)

            ctx.on_close(release)
        return h

    def getprecal(self, decf):
        """src/interpolation.jl:96-98 (host-side convenience; Horner in Float64)."""
        t = float(decf)
        out = np.empty(self.order + 1)
        for j in range(self.order + 1):
            ex = self.tabfct[j, -1]
            for k in range(self.tabfct.shape[1] - 2, -1, -1):
                ex = t * ex + self.tabfct[j, k]
            out[j] = ex
        return out

    def __repr__(self):
        return f"{type(self).__name__}{{Float64,CircEdge,{self.order}}}"


# @enum EdgeType, src/interpolation.jl:3
CircEdge, InsideEdge = 1, 2


class Lagrange(AbstractInterpolation):
    """Lagrange(order; edge = CircEdge) -- src/lagrange.jl:58-72.  edge = InsideEdge selects the
    non-periodic variant of the kernel seam (src/interpolation.jl:123-132, :250-286): one-sided stencils
    near the ends of the line instead of the periodic wrap."""

    kind = LAGRANGE

    def __init__(self, order, edge=CircEdge):
        super().__init__(order)
        if order < 1:
            raise ValueError("order must be >= 1")
        if edge not in (CircEdge, InsideEdge):
            raise ValueError("edge must be CircEdge or InsideEdge")
        self.edge = edge
        self.tabfct = _to_float_table(lagrange_table(order))


class _BSpline(AbstractInterpolation):
    def __init__(self, order, n):
        super().__init__(order, n)
        self.tabfct = _to_float_table(bspline_table(order))
        self.nodes = np.array([float(v) for v in bspline_nodes(order)], dtype=np.float64)


class BSplineLU(_BSpline):
    """BSplineLU(order, n) -- src/bsplinelu.jl:253-270 (odd orders only)"""

    kind = BSPLINE_LU

    def __init__(self, order, n):
        if order % 2 == 0:
            raise ValueError(f"order={order} BSplineLU for even  order is not implemented n={n}")
        super().__init__(order, n)


class BSplineFFT(_BSpline):
    """BSplineFFT(order, n) -- src/bsplinefft.jl:25-45; n must be a power of two
    (src/fftbig.jl:57).  Even orders are constructible in the reference but singular at the
    Nyquist mode for even n, so they are rejected here."""

    kind = BSPLINE_FFT

    def __init__(self, order, n):
        if n < 1 or (n & (n - 1)) != 0:
            raise ValueError(f"BSplineFFT: n={n} must be a power of two")
        if order % 2 == 0:
            raise ValueError(f"order={order}: even-order periodic B-splines are singular for even n")
        super().__init__(order, n)


class Hermite(AbstractInterpolation):
    """Hermite(order; flbis=false) -- src/hermite.jl:99-132"""

    kind = HERMITE

    def __init__(self, order, flbis=False):
        super().__init__(order)
        self.tabfct = _to_float_table(hermite_table(order, flbis=flbis))


def get_order(interp):
    return interp.order
