"""The oracle's Float64 line algorithms against EXACT rational arithmetic of the same mathematical
definition (the reference's formulas with Rational{BigInt}, which is how the reference itself tests
them: test/test_interpolation.jl:22-81).  This bounds what the unpinned last-ulp choices of the oracle
(FMA contraction in the Horner evaluation, summation order, LU elimination order) can amount to: a few
ulp of the largest value for the stencils, cond(A) times that for the B-spline solve.  CPU only."""
from fractions import Fraction as F
import math

import numpy as np
import pytest

from oracle import refmodel as R, tables as T


def _exact_shift(res, tab_rat, order, alpha):
    """out[i] = sum_j res[(i + decint - order/2 + j) mod n] * tabfct[j](decfloat), exactly"""
    n = len(res)
    decint = math.floor(alpha)
    t = F(alpha) - decint
    w = [T.p_eval(p, t) for p in tab_rat]
    out = []
    for i in range(n):
        s = F(0)
        for j in range(order + 1):
            s += res[(i + decint - order // 2 + j) % n] * w[j]
        out.append(s)
    return out


def _exact_circulant_solve(order, b):
    """A c = b with A[i][(i+m) mod n] = B((order+1)/2 + m) (test/test_splinelu.jl:10-29), Gauss in Fractions"""
    n = len(b)
    h = (order - 1) // 2
    nodes = T.bspline_node_values_rat(order)
    A = [[F(0)] * n for _ in range(n)]
    for i in range(n):
        for m in range(-h, h + 1):
            A[i][(i + m) % n] += nodes[h + m]
    M = [row[:] + [bi] for row, bi in zip(A, b)]
    for c in range(n):
        p = next(r for r in range(c, n) if M[r][c] != 0)
        M[c], M[p] = M[p], M[c]
        inv = 1 / M[c][c]
        M[c] = [v * inv for v in M[c]]
        for r in range(n):
            if r != c and M[r][c] != 0:
                f = M[r][c]
                M[r] = [vr - f * vc for vr, vc in zip(M[r], M[c])]
    return [M[i][n] for i in range(n)]


@pytest.mark.parametrize("kind,order", [("lagrange", 3), ("lagrange", 7), ("lagrange", 9), ("lagrange", 12), ("hermite", 5), ("hermite", 9)])
def test_stencil_kinds_against_exact_arithmetic(kind, order):
    rng = np.random.default_rng(order)
    n = 40
    f = rng.random(n)
    it = R.Lagrange(order) if kind == "lagrange" else R.Hermite(order)
    tab = T.lagrange_tabfct_rat(order) if kind == "lagrange" else T.hermite_tabfct_rat(order)
    for alpha in (0.345141526199181716, -1.2856139011444161, 4.98766514566778, -5.6785132567900988, 3.0, 0.999999999999):
        fp = np.empty(n)
        R.interpolate(fp, f, alpha, it)
        ex = _exact_shift([F(v) for v in f], tab, order, alpha)
        err = max(abs(F(a) - b) for a, b in zip(fp, ex))
        scale = max(abs(b) for b in ex)
        assert float(err / scale) <= 64 * np.finfo(np.float64).eps, (alpha, float(err / scale))


@pytest.mark.parametrize("kind,order,n", [("bspline_lu", 3, 16), ("bspline_lu", 5, 24), ("bspline_lu", 11, 32), ("bspline_fft", 5, 32), ("bspline_fft", 11, 32)])
def test_bspline_kinds_against_exact_arithmetic(kind, order, n):
    rng = np.random.default_rng(100 + order)
    f = rng.random(n)
    it = R.BSplineLU(order, n) if kind == "bspline_lu" else R.BSplineFFT(order, n)
    c_exact = _exact_circulant_solve(order, [F(v) for v in f])
    c = it.sol(f)
    scale = max(abs(v) for v in c_exact)
    # condition number of the collocation matrix: ~ (pi/2)^(order+1)
    tol = 32 * np.finfo(np.float64).eps * (math.pi / 2) ** (order + 1)
    assert float(max(abs(F(a) - b) for a, b in zip(c, c_exact)) / scale) <= tol
    tab = T.bspline_tabfct_rat(order)
    for alpha in (0.345141526199181716, -2.2856139011444161, 7.0):
        fp = np.empty(n)
        R.interpolate(fp, f, alpha, it)
        ex = _exact_shift(c_exact, tab, order, alpha)
        s2 = max(abs(b) for b in ex)
        assert float(max(abs(F(a) - b) for a, b in zip(fp, ex)) / s2) <= tol


def test_per_point_2d_against_exact_arithmetic():
    """the N-D per-point interpolate! (src/interpolation.jl:561-621) restated in oracle.c vs exact rationals"""
    from oracle import unsplit2d as U

    rng = np.random.default_rng(42)
    n1, n2 = 11, 9
    f = np.asfortranarray(rng.random((n1, n2)))
    dec = np.asfortranarray(rng.uniform(-6, 6, (n1, n2, 2)))
    oa, ob = 3, 5
    got = U.interpolate_points(f, dec, [R.Lagrange(oa), R.Lagrange(ob)])
    ta, tb = T.lagrange_tabfct_rat(oa), T.lagrange_tabfct_rat(ob)
    worst = F(0)
    for i in range(n1):
        for j in range(n2):
            da, db = math.floor(dec[i, j, 0]), math.floor(dec[i, j, 1])
            wa = [T.p_eval(p, F(dec[i, j, 0]) - da) for p in ta]
            wb = [T.p_eval(p, F(dec[i, j, 1]) - db) for p in tb]
            s = F(0)
            for b in range(ob + 1):
                for a in range(oa + 1):
                    s += F(f[(i + da - oa // 2 + a) % n1, (j + db - ob // 2 + b) % n2]) * wa[a] * wb[b]
            worst = max(worst, abs(F(got[i, j]) - s))
    assert float(worst) <= 64 * np.finfo(np.float64).eps * float(np.max(np.abs(got)))
