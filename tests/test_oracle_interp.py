"""Pins oracle.c (Float64 line algorithms) against the reference's known-answer tests.

test/test_interpolation.jl:22-81,488-492 (exact cubic-B-spline shift),
test/test_interpolation.jl:424-519 (analytic Float64 shifts, 100 repeated sweeps),
test/test_splinelu.jl:10-29,111-146,181-183,214-216 (LU layout and residuals),
test/testfftbig.jl:30-57 (DFT convention).   Paths relative to /root/reference.
"""
from fractions import Fraction as F
import math

import numpy as np
import pytest

from oracle import clib, refmodel as R, tables as T


# ---- tiny exact restatement of interpolate! for the rational KAT (pure Python, small n) ----
def interp_exact(fi, dec, tab_rat, order, solve=None):
    n = len(fi)
    decint = math.floor(dec)
    t = dec - decint
    w = [T.p_eval(p, t) for p in tab_rat]
    res = solve(fi) if solve else fi
    origin = -(order // 2)
    return [sum(res[(i + decint + origin + j) % n] * w[j] for j in range(order + 1)) for i in range(n)]


def solve_circulant_exact(nodes, order, n):
    kl, ku = T.get_kl_ku(order)
    A = [[F(0)] * n for _ in range(n)]
    for i in range(n):  # topl(): test/test_splinelu.jl:10-29
        for j, v in enumerate(nodes, start=1):
            A[i][(i + j - kl - 1) % n] = v
    def solve(b):
        M = [row[:] + [bi] for row, bi in zip(A, b)]
        for c in range(n):
            piv = next(r for r in range(c, n) if M[r][c] != 0)
            M[c], M[piv] = M[piv], M[c]
            inv = 1 / M[c][c]
            M[c] = [x * inv for x in M[c]]
            for r in range(n):
                if r != c and M[r][c] != 0:
                    f = M[r][c]
                    M[r] = [x - f * y for x, y in zip(M[r], M[c])]
        return [M[r][n] for r in range(n)]
    return solve


def test_exact_cubic_bspline_shift_kat():
    # test/test_interpolation.jl:22-81 with sz=32 (the reference uses 128; the property --
    # B-spline interpolation of a cubic B-spline is exact, Lagrange-3 is exact except at the
    # 2*(order+1) points whose stencil straddles a knot... -- is size independent up to the count)
    order, sz = 3, 32
    sp = T.getbspline(order, 0)
    mesh = [F((order + 1) * i, sz) for i in range(sz)]
    deb = [T.bspline_eval(sp, x) for x in mesh]
    nodes = T.bspline_node_values_rat(order)
    solve = solve_circulant_exact(nodes, order, sz)
    btab, ltab = T.bspline_tabfct_rat(order), T.lagrange_tabfct_rat(order)
    for i in range(3):
        dec = F(1235, 10240) + i
        decint, decf = math.floor(dec), dec - math.floor(dec)
        ref = [T.bspline_eval(sp, x + (order + 1) * decf / sz) for x in mesh]
        ref = ref[decint % sz:] + ref[: decint % sz]  # circshift(ref, -decint)
        resb = interp_exact(deb, dec, btab, order, solve)
        assert resb == ref
        resl = interp_exact(deb, dec, ltab, order)
        # Lagrange: exact wherever the 4-point stencil sits inside one cubic piece
        assert sum(1 for a, b in zip(resl, ref) if a == b) == sz - 8
        # and the Float64 C oracle agrees with the exact results
        for kind, exact in ((R.Lagrange(3), resl), (R.BSplineLU(3, sz), resb)):
            fp = np.empty(sz)
            R.interpolate(fp, np.array([float(x) for x in deb]), float(dec), kind)
            assert np.max(np.abs(fp - np.array([float(x) for x in exact]))) < 2e-16 * 8


def test_exact_kat_full_size_lagrange_count():
    # test/test_interpolation.jl:488: sz = 128 -> exactly 120 exact points for Lagrange 3
    order, sz = 3, 128
    sp = T.getbspline(order, 0)
    mesh = [F((order + 1) * i, sz) for i in range(sz)]
    deb = [T.bspline_eval(sp, x) for x in mesh]
    dec = F(1235, 10240) + 2
    decint, decf = 2, F(1235, 10240)
    ref = [T.bspline_eval(sp, x + (order + 1) * decf / sz) for x in mesh]
    ref = ref[decint:] + ref[:decint]
    res = interp_exact(deb, dec, T.lagrange_tabfct_rat(order), order)
    assert sum(1 for a, b in zip(res, ref) if a == b) == 120


TABDEC = [
    0.345141526199181716726626262655544,
    -0.3859416191876155241320011187619,
    -1.28561390114441619187615524132001118762519,
    -0.885901390114441619187615524132001118762519,
    -5.678513256790098898776656565545454544544545,
    4.9876651456677809099887665655556565565656565,
    0.186666659416191876155241320011187619,
    0.590999232323232323232365566787878898898,
    1.231098015934444444444444788888888878878,
]


def _interpfloat(interp, sz, tol, nb=100):
    # test/test_interpolation.jl:424-483
    mesh = np.arange(sz) / sz
    fcts = [lambda x: np.cos(2 * np.pi * x + 0.25), lambda x: np.exp(-((np.cos(2 * np.pi * x + 0.25) - 1) ** 2))]
    nmax = 0.0
    for fct in fcts:
        for dec in TABDEC:
            fp = fct(mesh)
            decint = math.floor(dec)
            value = dec - decint
            if interp.order % 2 == 0 and value > 0.5:
                value -= 1
                decint += 1
            precal = interp.getprecal(value)
            for i in range(1, nb + 1):
                fi = fp.copy()
                ref = fct(mesh + i * dec / sz)
                fp = np.empty(sz)
                R.interpolate_precal(fp, fi, decint, precal, interp)
                nmax = max(nmax, float(np.max(np.abs(fp - ref))))
    assert nmax < tol, nmax
    return nmax


def test_interpfloat_lagrange():
    _interpfloat(R.Lagrange(3), 128, 1e-3)      # :495
    _interpfloat(R.Lagrange(9), 256, 1e-10)     # :501
    _interpfloat(R.Lagrange(4), 256, 1e-5)      # :504
    _interpfloat(R.Lagrange(12), 256, 1e-10)    # :507


def test_interpfloat_bspline_lu():
    _interpfloat(R.BSplineLU(3, 256), 256, 1e-5)     # :510
    _interpfloat(R.BSplineLU(11, 256), 256, 1e-12)   # :513


def test_interpfloat_bspline_fft():
    _interpfloat(R.BSplineFFT(3, 256), 256, 1e-5)    # :516
    _interpfloat(R.BSplineFFT(11, 256), 256, 1e-12)  # :519


def test_interpfloat_hermite():
    # Hermite has no test_interpfloat call in the reference; same harness, loose tolerance
    _interpfloat(R.Hermite(9), 256, 1e-6)


def topl(n, t):
    # test/test_splinelu.jl:10-29 (circular)
    kl, ku = T.get_kl_ku(len(t))
    A = np.zeros((n, n))
    for i in range(n):
        for j, v in enumerate(t, start=1):
            A[i, (i + j - kl - 1) % n] = v
    return A


@pytest.mark.parametrize("n,order", [(30, 9), (30, 3), (31, 5), (128, 9), (1000, 9)])
def test_luspline_layout_and_residual(n, order):
    L = clib.lib()
    nodes = np.array([float(x) for x in T.bspline_node_values_rat(order)])
    A = topl(n, nodes)
    # (a) un-factorised storage reproduces the circulant matrix (band / lastrows / lastcols)
    h = L.orc_lu_create(n, clib.dp(nodes), order, 0)
    dims = (clib.C.c_long * 5)()
    L.orc_lu_dims(h, dims)
    _, szb, wd, kl, ku = list(dims)
    band = np.ctypeslib.as_array(L.orc_lu_band(h), shape=(szb, wd)).T          # (wd, szb) col-major
    lastrows = np.ctypeslib.as_array(L.orc_lu_lastrows(h), shape=(n, ku)).T    # (ku, n)
    lastcols = np.ctypeslib.as_array(L.orc_lu_lastcols(h), shape=(kl, n - ku)).T  # (n-ku, kl)
    assert np.array_equal(lastrows, A[n - ku:, :])
    assert np.array_equal(lastcols, A[: n - ku, n - kl:])
    for j in range(szb):  # src/bsplinelu.jl:150-156
        for i in range(j - ku, j + kl + 1):
            if 0 <= i < n - ku:
                assert band[ku + i - j, j] == A[i, j]
    L.orc_lu_destroy(h)
    # (b) test/test_splinelu.jl:181-183, :214-216: residual of the factorised solve
    it = R.BSplineLU(order, n)
    rng = np.random.default_rng(5431221)
    b = rng.random(n)
    x = it.sol(b)
    tol = 1e-12 if n == 1000 else 1e-10
    assert np.max(np.abs(A @ x - b)) < tol
    # (c) agrees with a dense LAPACK solve
    assert np.max(np.abs(x - np.linalg.solve(A, b))) < 1e-12 * max(1.0, np.max(np.abs(x)))


def test_fft_convention():
    # test/testfftbig.jl:30-57: forward exp(-2 pi i jk/n) unnormalised, inverse /n (== FFTW == numpy)
    L = clib.lib()
    rng = np.random.default_rng(1)
    for n in (8, 128, 1024):
        z = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        re, im = z.real.copy(), z.imag.copy()
        assert L.orc_fft(clib.dp(re), clib.dp(im), n, 0) == 0
        ref = np.fft.fft(z)
        assert np.max(np.abs(re + 1j * im - ref)) < 1e-15 * n * np.max(np.abs(ref))
        assert L.orc_fft(clib.dp(re), clib.dp(im), n, 1) == 0
        assert np.max(np.abs(re + 1j * im - z)) < 1e-14
    re = np.zeros(12); im = np.zeros(12)
    assert L.orc_fft(clib.dp(re), clib.dp(im), 12, 0) != 0  # src/fftbig.jl:57 power-of-two assert


@pytest.mark.parametrize("order,n", [(3, 64), (5, 128), (11, 256)])
def test_bspline_fft_equals_lu_and_numpy(order, n):
    # src/bsplinefft.jl:36-39,49-51; both types solve the same circulant system for odd orders
    rng = np.random.default_rng(20240611)
    b = rng.random(n)
    xf = R.BSplineFFT(order, n).sol(b)
    xl = R.BSplineLU(order, n).sol(b)
    nodes = np.array([float(x) for x in T.bspline_node_values_rat(order)])
    kl, ku = T.get_kl_ku(order)
    c = np.zeros(n)
    for i in range(1, order + 1):
        c[(n - kl - 1 + i) % n] = nodes[i - 1]
    xn = np.real(np.fft.ifft(np.fft.fft(b) / np.fft.fft(c)))
    scale = np.max(np.abs(xl))
    assert np.max(np.abs(xf - xn)) < 1e-13 * scale
    assert np.max(np.abs(xf - xl)) < 1e-12 * scale
    with pytest.raises(ValueError):
        R.BSplineFFT(order, 100)
    with pytest.raises(ValueError):
        R.BSplineLU(4, 64)


def test_split_alpha_edge_cases():
    # src/interpolation.jl:384-385: decfloat may be exactly 1.0 for a tiny negative alpha
    L = clib.lib()
    t = clib.C.c_double()
    assert L.orc_split_alpha(-1e-20, clib.C.byref(t)) == -1 and t.value == 1.0
    assert L.orc_split_alpha(2.75, clib.C.byref(t)) == 2 and t.value == 0.75
    assert L.orc_split_alpha(-5.25, clib.C.byref(t)) == -6 and t.value == 0.75
    # weights at t = 0 select node 0, at t = 1 node +1 (continuity across the integer boundary)
    it = R.Lagrange(7)
    w0, w1 = it.getprecal(0.0), it.getprecal(1.0)
    assert w0[3] == 1.0 and np.count_nonzero(w0) == 1
    assert abs(w1[4] - 1.0) < 1e-13 and np.max(np.abs(np.delete(w1, 4))) < 1e-13


def test_sweep_equals_line_loop_and_threads():
    # orc_sweep == loop of orc_interpolate_alpha over lines, for every dim, 1 or many threads
    rng = np.random.default_rng(20240611)
    ext = (12, 10, 9)
    f = np.asfortranarray(rng.random(ext))
    it = R.Lagrange(5)
    L = clib.lib()
    for dim in range(3):
        other = [d for d in range(3) if d != dim]
        tshape = [ext[d] for d in other]
        tab = np.asfortranarray(rng.uniform(-8, 8, tshape))
        astride = [0, 0, 0]
        astride[other[0]] = 1
        astride[other[1]] = tshape[0]
        ref = f.copy(order="F")
        for i0 in range(tshape[0]):
            for i1 in range(tshape[1]):
                idx = [slice(None)] * 3
                idx[other[0]], idx[other[1]] = i0, i1
                line = np.ascontiguousarray(f[tuple(idx)])
                out = np.empty(ext[dim])
                R.interpolate(out, line, tab[i0, i1], it)
                ref[tuple(idx)] = out
        for nth in (1, 4):
            g = f.copy(order="F")
            scratch = np.empty(g.size)
            rc = L.orc_sweep(g.ctypes.data_as(clib.c_double_p), clib.dp(scratch), 3, clib.lp(ext), dim, it._h,
                             clib.dp(tab.reshape(-1, order="F").copy()), clib.lp(astride), nth)
            assert rc == 0
            assert np.array_equal(g, ref)


def test_oracle_nd_constant_shift_state_equals_split_1d_states():
    """src/interpolation.jl:212-231: the N-D constant-shift tensor stencil of a state with ndims = 2
    (test/test_poisson2d.jl:276) is the product of the 1-D stencils, so a 2D2V Vlasov-Poisson run
    with states [(v1 v2), (x1 x2)] must agree with the run whose states are the four 1-D sweeps
    (examples/vlasov-poisson-2d2v.jl:126-131) to rounding."""
    import math

    from oracle import refmodel as R

    def build(states, n=10):
        ms = (R.UniformMesh(0.0, 4 * math.pi, n), R.UniformMesh(0.0, 4 * math.pi, n + 2),
              R.UniformMesh(-6.0, 6.0, n + 4), R.UniformMesh(-6.0, 6.0, n))
        adv = R.Advection(ms, [R.Lagrange(5)] * 4, 0.1, states)
        fsp = lambda x: 0.5 * np.cos(x / 2) + 1
        fv = lambda v: np.exp(-v**2 / 2) / math.sqrt(2 * math.pi)
        f = R.dotprod((fsp(ms[0].points), fsp(ms[1].points), fv(ms[2].points), fv(ms[3].points)))
        return R.AdvectionData(adv, f, R.getpoissonvar(adv))

    a = build([([3, 4, 1, 2], 2, 1, True), ([1, 2, 3, 4], 2, 2, True)])
    b = build([([3, 4, 1, 2], 1, 1, True), ([4, 3, 1, 2], 1, 1, True), ([1, 2, 4, 3], 1, 2, True), ([2, 1, 3, 4], 1, 2, True)])
    assert a.adv.nbstates == 3 and b.adv.nbstates == 6
    for _ in range(2):
        while R.advection(a):
            pass
        while R.advection(b):
            pass
    assert np.max(np.abs(a.data - b.data)) <= 1e-13 * np.max(np.abs(b.data))
    assert abs(R.compute_ee(a) - R.compute_ee(b)) <= 1e-12 * abs(R.compute_ee(b))


TABDEC_INSIDE = [0.345141526199181716726626262655544, -0.3859416191876155241320011187619, -1.28561390114441619187615524132001118762519,
                 -0.885901390114441619187615524132001118762519, 0.186666659416191876155241320011187619,
                 0.590999232323232323232365566787878898898, 1.231098015934444444444444788888888878878]


def test_inside_edge_kats():
    """InsideEdge Lagrange (src/interpolation.jl:123-132, :250-286), pinned by the reference's own tests:
    a cubic is reproduced at EVERY point, ends included (test/test_interpolation.jl:22-75, :488: exact in
    rational arithmetic there, to rounding here), and the analytic shifts of test_interpfloat
    (:424-483, :498: Lagrange 7, n = 128, 3 repeated shifts, tolerance 1e-3, |decint| <= 2)."""
    sz = 128
    mesh = np.arange(sz) / sz
    cubic = lambda x: x**3 - x**2 - x / 6 + 0.25
    fp = np.empty(sz)
    R.interpolate(fp, cubic(mesh), 3 / 1024, R.Lagrange(3, edge=R.InsideEdge))
    assert np.max(np.abs(fp - cubic(mesh + (3 / 1024) / sz))) < 1e-15
    it = R.Lagrange(7, edge=R.InsideEdge)
    for fct in (lambda x: np.cos(2 * np.pi * x + 0.25), lambda x: np.exp(-((np.cos(2 * np.pi * x + 0.25) - 1) ** 2))):
        for dec in TABDEC_INSIDE:
            fp = fct(mesh)
            for i in range(1, 4):
                fi = fp.copy()
                fp = np.empty(sz)
                R.interpolate(fp, fi, dec, it)
                assert np.max(np.abs(fp - fct(mesh + i * dec / sz))) < 1e-3
    # a shift whose window leaves the array is an error (the reference indexes out of bounds)
    with pytest.raises(ValueError):
        R.interpolate(np.empty(sz), cubic(mesh), 5.7, it)
    # away from the ends InsideEdge and CircEdge agree
    a, b = np.empty(sz), np.empty(sz)
    f = np.random.default_rng(2).random(sz)
    R.interpolate(a, f, 1.3, it)
    R.interpolate(b, f, 1.3, R.Lagrange(7))
    assert np.array_equal(a[8:-8], b[8:-8]) and not np.array_equal(a[:3], b[:3])
