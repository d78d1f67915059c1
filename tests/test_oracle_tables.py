"""Pins oracle/tables.py against the reference's own exact-arithmetic tests.

test/test_lagrange.jl:18-31, test/test_bspline.jl:7-121, test/test_hermite.jl:19-68
(paths relative to /root/reference), re-expressed with fractions.Fraction.
"""
import random
from fractions import Fraction as F

import pytest

from oracle import tables as T


def rand_poly(rng, deg):
    return [F(rng.randint(-1000, 1000), rng.randint(1, 1000)) for _ in range(deg + 1)]


@pytest.mark.parametrize("order", [3, 4, 5, 7, 9, 11, 12, 19, 27])
def test_lagrange_reproduces_polynomials(order):
    # test/test_lagrange.jl:18-31: sum_j tabfct[j+1](x) * p(j - dec) == p(x)
    rng = random.Random(1000 + order)
    tab = T.lagrange_tabfct_rat(order)
    dec = order // 2
    for _ in range(3):
        p = rand_poly(rng, order)
        x = F(rng.randint(0, 10**6), 10**6)
        res = sum(T.p_eval(tab[j], x) * T.p_eval(p, F(j - dec)) for j in range(order + 1))
        assert res == T.p_eval(p, x)


def test_lagrange_order3_known_weights():
    # order 3 weights at x = 1/4 are (-7, 105, 35, -5)/128
    tab = T.lagrange_tabfct_rat(3)
    assert [T.p_eval(p, F(1, 4)) for p in tab] == [F(-7, 128), F(105, 128), F(35, 128), F(-5, 128)]


def test_get_kl_ku():
    # test/test_bspline.jl:114-121
    assert T.get_kl_ku(5) == (2, 2)
    assert T.get_kl_ku(6) == (2, 3)


@pytest.mark.parametrize("order", [1, 2, 3, 5, 9, 11])
def test_bspline_partition_of_unity_and_continuity(order):
    # test/test_bspline.jl:7-68
    sp = T.getbspline(order, 0)
    assert len(sp) == order + 1
    rng = random.Random(order)
    for _ in range(5):
        x = F(rng.randint(0, 10**6), 10**6)
        assert sum(T.bspline_eval(sp, x + i) for i in range(order + 1)) == 1
    # C^{order-1} at the knots: all derivatives up to order-1 agree
    def deriv(p, k):
        p = list(p)
        for _ in range(k):
            p = [i * p[i] for i in range(1, len(p))] or [F(0)]
        return p
    for knot in range(1, order + 1):
        for k in range(order):
            assert T.p_eval(deriv(sp[knot - 1], k), F(knot)) == T.p_eval(deriv(sp[knot], k), F(knot))


def test_bspline_matches_de_boor():
    # test/test_bspline.jl:70-112: equality with the Cox-de Boor recursion, evaluated pointwise
    def deboor(p, j, x):
        if p == 0:
            return F(1) if j <= x < j + 1 else F(0)
        return (x - j) / p * deboor(p - 1, j, x) + (j + p + 1 - x) / p * deboor(p - 1, j + 1, x)
    rng = random.Random(7)
    for order in (3, 5, 7):
        sp = T.getbspline(order, 0)
        for _ in range(10):
            x = F(rng.randint(0, (order + 1) * 1000), 1000)
            assert T.bspline_eval(sp, x) == deboor(order, 0, x)


def test_bspline_node_values():
    # SURVEY.md 3.3 check: order 5 -> (1, 26, 66, 26, 1)/120
    assert T.bspline_node_values_rat(5) == [F(1, 120), F(26, 120), F(66, 120), F(26, 120), F(1, 120)]
    assert T.bspline_node_values_rat(3) == [F(1, 6), F(4, 6), F(1, 6)]


@pytest.mark.parametrize("order", [3, 5, 11])
def test_bspline_tabfct_is_shifted_pieces(order):
    # src/bsplinelu.jl:263: tabfct[j+1](t) = B(order - j + t)
    sp = T.getbspline(order, 0)
    tab = T.bspline_tabfct_rat(order)
    t = F(3, 11)
    for j in range(order + 1):
        assert T.p_eval(tab[j], t) == T.bspline_eval(sp, order - j + t)
    assert sum(T.p_eval(p, t) for p in tab) == 1


def test_hermite_bplus_known_answers():
    # test/test_hermite.jl:48-52
    assert T.PrecalHermite(3).bplus == [F(-1, 3), F(-1, 2), F(1), F(-1, 6)]
    assert T.PrecalHermite(5).bplus == [F(1, 20), F(-1, 2), F(-1, 3), F(1), F(-1, 4), F(1, 30)]
    for o in (3, 5, 7):
        ph = T.PrecalHermite(o)
        assert sum(ph.bplus) == 0 and sum(ph.bminus) == 0


@pytest.mark.parametrize("order", [5, 9, 13, 17])
def test_hermite_reproduces_polynomials(order):
    # test/test_hermite.jl:54-68: polynomials of degree ord = div(order,2)+1 are reproduced
    herm = T.hermite_tabfct_rat(order)
    ord_ = order // 2 + 1
    assert len(herm) == order + 1
    assert max(len(p) - 1 for p in herm) == 2 * ord_ + 1
    rng = random.Random(order)
    dec = order // 2
    for _ in range(3):
        p = rand_poly(rng, ord_)
        x = F(rng.randint(0, 10**6), 10**6)
        res = sum(T.p_eval(herm[j], x) * T.p_eval(p, F(j - dec)) for j in range(order + 1))
        assert res == T.p_eval(p, x)


def test_hermite_domain_errors():
    with pytest.raises(ValueError):
        T.hermite_tabfct_rat(7)
    with pytest.raises(ValueError):
        T.hermite_tabfct_rat(5, flbis=True)
    assert len(T.hermite_tabfct_rat(7, flbis=True)) == 8
