"""ORACLE (test infrastructure, not product code) -- exact-rational interpolation tables.

CPU restatement of the reference's host-side table constructors.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import anything under `oracle/`.

Every function cites the reference file:line (relative to /root/reference) it follows.
All arithmetic is `fractions.Fraction` (== Julia `Rational{BigInt}`); the final
`to_float64_table` rounding is the `convert.(Polynomial{Float64}, tabfct_rat)` step
(correctly rounded Rational -> Float64, same as Julia).

Parity status: these tables are pinned by the reference's own exact tests
(test/test_lagrange.jl:18-31, test/test_bspline.jl:70-121, test/test_hermite.jl:48-68),
re-expressed in tests/test_oracle_tables.py.
"""
from fractions import Fraction
from math import floor

# --------------------------------------------------------------------------------------
# minimal exact polynomial algebra (ascending coefficient lists of Fraction);
# stands in for Polynomials.jl `Polynomial{Rational{BigInt}}`
# --------------------------------------------------------------------------------------


def p_trim(a):
    a = list(a)
    while len(a) > 1 and a[-1] == 0:
        a.pop()
    return a


def p_add(a, b):
    n = max(len(a), len(b))
    return p_trim([(a[i] if i < len(a) else 0) + (b[i] if i < len(b) else 0) for i in range(n)])


def p_sub(a, b):
    return p_add(a, [-x for x in b])


def p_mul(a, b):
    r = [Fraction(0)] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        if x == 0:
            continue
        for j, y in enumerate(b):
            r[i + j] += x * y
    return p_trim(r)


def p_scale(a, s):
    return p_trim([x * s for x in a])


def p_eval(a, x):
    r = Fraction(0)
    for c in reversed(a):
        r = r * x + c
    return r


def p_compose(a, b):
    """a(b(x))"""
    r = [Fraction(0)]
    for c in reversed(a):
        r = p_add(p_mul(r, b), [Fraction(c)])
    return r


ONE = [Fraction(1)]
ZERO = [Fraction(0)]

# --------------------------------------------------------------------------------------
# Lagrange -- src/lagrange.jl:25-35 (_getpolylagrange), :58-72 (Lagrange ctor)
# --------------------------------------------------------------------------------------


def getpolylagrange(k, order, origin):
    """src/lagrange.jl:25-35: prod_{l != k} (x - l - origin) / (k - l)."""
    if not (0 <= k <= order):
        raise ValueError("the constant 0 <= k <= order is false")
    result = ONE
    for l in range(order + 1):
        if l != k:
            d = Fraction(k - l)
            result = p_mul(result, [Fraction(-(l + origin)) / d, Fraction(1) / d])
    return result


def lagrange_tabfct_rat(order):
    """src/lagrange.jl:62-70: origin = -div(order,2); tabfct[i+1] = _getpolylagrange(i, order, origin)."""
    origin = -(order // 2)
    return [getpolylagrange(i, order, origin) for i in range(order + 1)]



def p_integrate(a):
    """Polynomials.integrate: antiderivative with zero constant term"""
    return p_trim([Fraction(0)] + [Fraction(c) / (i + 1) for i, c in enumerate(a)])


def abcoef_rat(ordermax):
    """src/lagrange.jl:74-88  ABcoef(ordermax): tab[i, j] = _c(i-1, j-1) for i <= j, with
    _c(k, n) = P(0) - P(-1), P = integrate(_getpolylagrange(k, n, 0)): the Adams-Bashforth
    weights (integral over [-1, 0] of the Lagrange basis on the nodes 0..n).  Returned 0-based:
    tab[i][j] (Fraction), zero above the diagonal."""
    tab = [[Fraction(0)] * ordermax for _ in range(ordermax)]
    for j in range(1, ordermax + 1):
        for i in range(1, j + 1):
            P = p_integrate(getpolylagrange(i - 1, j - 1, 0))
            tab[i - 1][j - 1] = p_eval(P, Fraction(0)) - p_eval(P, Fraction(-1))
    return tab

# --------------------------------------------------------------------------------------
# cardinal B-spline -- src/spline.jl:6-97
# A Spline is a list of polynomial pieces; piece i is valid on [i, i+1).
# --------------------------------------------------------------------------------------


def _sp_get(sp, i):
    """src/spline.jl:14-21 getindex (0-based), zero polynomial outside."""
    return sp[i] if 0 <= i < len(sp) else ZERO


def _sp_add(a, b):
    """src/spline.jl:25-32"""
    n = max(len(a), len(b))
    return [p_add(_sp_get(a, i), _sp_get(b, i)) for i in range(n)]


def _sp_mulpoly(a, pol):
    """src/spline.jl:47-53"""
    return [p_mul(p, pol) for p in a]


def _sp_decal(a, n):
    """src/spline.jl:54-65: shift the spline by n to the right (compose with x-n, prepend n zero pieces)."""
    if n == 0:
        return a
    poldec = [Fraction(-n), Fraction(1)]
    return [ZERO] * n + [p_compose(p, poldec) for p in a]


def _w(p, j):
    """src/spline.jl:66: w(p,j) = (x - j)/p"""
    return [Fraction(-j, p), Fraction(1, p)]


def _getbspline(n, j):
    """src/spline.jl:67-76: Cox-de Boor recursion on piecewise polynomials."""
    if n == 0:
        return _sp_decal([ONE], j)
    n1 = _getbspline(n - 1, j)
    n2 = _sp_decal(n1, 1)
    return _sp_add(_sp_mulpoly(n1, _w(n, j)), _sp_mulpoly(n2, p_sub(ONE, _w(n, j + 1))))


def getbspline(n, j=0):
    """src/spline.jl:82-88 (BigInt switch is moot with Python ints)."""
    return _getbspline(n, j)


def bspline_eval(sp, x):
    """src/spline.jl:90-97: i = floor(x); piece_i(x) if 0 <= i < size else 0."""
    i = floor(x)
    if 0 <= i < len(sp):
        return p_eval(sp[i], Fraction(x))
    return Fraction(0)


def get_kl_ku(order):
    """src/bspline.jl:12-16"""
    ku = order // 2
    kl = order - 1 - ku
    return kl, ku


def bspline_tabfct_rat(order):
    """src/bsplinelu.jl:262-263 / src/bsplinefft.jl:30-31:
    tabfct_rat[x+1] = bspline[order-x](Polynomial([order-x, 1])), x = 0..order."""
    bsp = getbspline(order, 0)
    return [p_compose(_sp_get(bsp, order - x), [Fraction(order - x), Fraction(1)]) for x in range(order + 1)]


def bspline_node_values_rat(order):
    """src/bsplinelu.jl:264 / src/bsplinefft.jl:35: bspline.(1:order)."""
    bsp = getbspline(order, 0)
    return [bspline_eval(bsp, Fraction(i)) for i in range(1, order + 1)]


# --------------------------------------------------------------------------------------
# Hermite -- src/hermite.jl:4-133
# --------------------------------------------------------------------------------------


def _L(i, ord_):
    """src/hermite.jl:4-14"""
    if ord_ % 2 != 1:
        raise ValueError("ord must be odd")
    d = ord_ // 2
    result = ONE
    for j in range(-d, d + 2):
        if j != i:
            den = Fraction(i - j)
            result = p_mul(result, [Fraction(-j) / den, Fraction(1) / den])
    return result


def _Lprim(i, ord_):
    """src/hermite.jl:16-25"""
    d = ord_ // 2
    r = Fraction(0)
    for j in range(-d, d + 2):
        if i != j:
            r += Fraction(1, i - j)
    return r


def _K(i, ord_):
    """src/hermite.jl:26"""
    l = _L(i, ord_)
    return p_mul(p_mul(l, l), [Fraction(-i), Fraction(1)])


def _H(i, ord_):
    """src/hermite.jl:27"""
    l = _L(i, ord_)
    return p_mul(p_mul(l, l), p_sub(ONE, p_scale([Fraction(-i), Fraction(1)], 2 * _Lprim(i, ord_))))


def _bplus(i, rplus, splus):
    """src/hermite.jl:29-32"""
    res = Fraction(1)
    for j in range(rplus, splus + 1):
        if j != 0 and j != i:
            res *= Fraction(-j)
    for j in range(rplus, splus + 1):
        if j != i:
            res *= Fraction(1, i - j)
    return res


def _bminus(i, rminus, sminus):
    """src/hermite.jl:34-36"""
    return -_bplus(-i, -sminus, -rminus)


class PrecalHermite:
    """src/hermite.jl:44-80"""

    def __init__(self, ord_, flbis=False):
        if ord_ % 2 != 1:
            raise ValueError("ord must be odd")
        d = ord_ // 2
        self.ord, self.d = ord_, d
        rng = range(-d, d + 2)
        self.L = [_L(i, ord_) for i in rng]
        self.Lprim = [_Lprim(i, ord_) for i in rng]
        self.K = [_K(i, ord_) for i in rng]
        self.H = [_H(i, ord_) for i in rng]
        if flbis:
            rplus, splus = -d - 1, d
        else:
            rplus, splus = -d, d + 1
        rminus, sminus = -splus, -rplus
        self.rplus, self.splus, self.rminus, self.sminus = rplus, splus, rminus, sminus
        self.bplus = [(_bplus(i, rplus, splus) if i != 0 else Fraction(0)) for i in range(rplus, splus + 1)]
        self.bminus = [(_bminus(i, rminus, sminus) if i != 0 else Fraction(0)) for i in range(rminus, sminus + 1)]
        self.bplus[-rplus] = -sum(self.bplus)
        self.bminus[-rminus] = -sum(self.bminus)

    def Kf(self, i):
        return self.K[self.d + i]

    def Hf(self, i):
        return self.H[self.d + i]

    def bp(self, i):
        return self.bplus[-self.rplus + i]

    def bm(self, i):
        return self.bminus[-self.rminus + i]


def hermite_tabfct_rat(order, flbis=False):
    """src/hermite.jl:99-132"""
    if flbis:
        if order % 4 != 3:
            raise ValueError("order modulo 4 must equal to 3")
        ord_ = order // 2
    else:
        if order % 4 != 1:
            raise ValueError("order modulo 4 must equal to 1")
        ord_ = order // 2 + 1
    decal = (order + 1) // 2  # 1-based slot of stencil node 0
    d = ord_ // 2
    ph = PrecalHermite(ord_, flbis=flbis)
    tab = [ZERO for _ in range(order + 1)]  # tab[s-1] <-> Julia tabfct[s]
    for i in range(-d, d + 2):
        tab[decal + i - 1] = p_add(tab[decal + i - 1], ph.Hf(i))
    for i in range(-d, 1):
        for k in range(ph.rplus, ph.splus + 1):
            tab[decal + i + k - 1] = p_add(tab[decal + i + k - 1], p_scale(ph.Kf(i), ph.bp(k)))
    for i in range(1, d + 2):
        for k in range(ph.rminus, ph.sminus + 1):
            tab[decal + i + k - 1] = p_add(tab[decal + i + k - 1], p_scale(ph.Kf(i), ph.bm(k)))
    return tab


# --------------------------------------------------------------------------------------
# Rational -> Float64 (convert.(Polynomial{T}, tabfct_rat))
# --------------------------------------------------------------------------------------


def to_float64_table(tab_rat):
    """Row j = ascending Float64 coefficients of weight polynomial j, zero-padded to a
    common length.  float(Fraction) is correctly rounded, like Julia's Rational{BigInt}->Float64."""
    ncoef = max(len(p) for p in tab_rat)
    return [[float(p[k]) if k < len(p) else 0.0 for k in range(ncoef)] for p in tab_rat]
