"""GPU parity of the sweep kernels (K1, K2) against the oracle, through the C ABI.

Tolerance (BASELINE.json north_star): per sweep, max|gpu - oracle| / max|oracle| <= 1e-12.
With SLB_SWEEP_EXACT the Lagrange/Hermite stencil is evaluated in the reference's own
operation order and must agree with the oracle BIT FOR BIT.
"""
import math

import numpy as np
import pytest

from helpers import DeviceGrid, make_pair, oracle_sweep, relerr
from oracle import refmodel as R

pytestmark = pytest.mark.gpu
TOL = 1e-12
SEED = 20240611


def _alpha_case(rng, shape, dim, mode):
    """alpha table + strides over the non-swept dims: 'full' = one value per line,
    'lastdim' = depends on one other dim only (Vlasov space sweep), 'const' = broadcast."""
    nd = len(shape)
    other = [d for d in range(nd) if d != dim]
    astride = [0] * nd
    if mode == "const" or not other:
        return np.array([rng.uniform(-8, 8)]), astride
    if mode == "lastdim":
        d = other[-1]
        astride[d] = 1
        return rng.uniform(-8, 8, shape[d]), astride
    stride = 1
    for d in other:
        astride[d] = stride
        stride *= shape[d]
    return rng.uniform(-8, 8, stride), astride


@pytest.mark.parametrize("kind,order", [("lagrange", o) for o in (3, 4, 5, 7, 9, 11, 12, 13)] + [("hermite", 5), ("hermite", 9)])
@pytest.mark.parametrize("shape", [(128, 40), (40, 128), (64, 33), (33, 64), (50, 7, 9), (16, 12, 10, 14), (256, 8), (8, 256), (1000, 3)])
def test_stencil_sweeps_match_oracle(kind, order, shape):
    rng = np.random.default_rng(SEED + order)
    f = np.asfortranarray(rng.random(shape))
    for dim in range(len(shape)):
        interp, ointerp = make_pair(kind, order, shape[dim])
        for mode in ("full", "lastdim", "const"):
            tab, astride = _alpha_case(rng, shape, dim, mode)
            ref = oracle_sweep(f, dim, ointerp, tab, astride)
            for flags, exact in ((0, False), (1, True)):
                g = DeviceGrid(f)
                g.sweep(dim, interp, tab, astride, flags=flags)
                out = g.get()
                g.close()
                if exact:
                    assert np.array_equal(out, ref), (kind, order, shape, dim, mode, relerr(out, ref))
                else:
                    assert relerr(out, ref) <= TOL, (kind, order, shape, dim, mode, relerr(out, ref))


@pytest.mark.parametrize("order", [15, 19, 27])
def test_high_order_generic_kernel(order):
    rng = np.random.default_rng(SEED)
    shape = (64, 20, 6)
    f = np.asfortranarray(rng.random(shape))
    for dim in range(3):
        interp, ointerp = make_pair("lagrange", order, shape[dim])
        tab, astride = _alpha_case(rng, shape, dim, "full")
        ref = oracle_sweep(f, dim, ointerp, tab, astride)
        g = DeviceGrid(f)
        g.sweep(dim, interp, tab, astride)
        assert relerr(g.get(), ref) <= TOL


@pytest.mark.parametrize("kind,order,n", [("bspline_lu", 3, 64), ("bspline_lu", 5, 128), ("bspline_lu", 11, 128), ("bspline_lu", 5, 100),
                                          ("bspline_fft", 5, 128), ("bspline_fft", 11, 128), ("bspline_fft", 11, 256), ("bspline_lu", 9, 1000)])
def test_bspline_sweeps_match_oracle(kind, order, n):
    rng = np.random.default_rng(SEED)
    for shape, dim in (((n, 70), 0), ((70, n), 1), ((6, n, 5), 1)):
        f = np.asfortranarray(rng.random(shape))
        interp, ointerp = make_pair(kind, order, n)
        tab, astride = _alpha_case(rng, shape, dim, "full")
        ref = oracle_sweep(f, dim, ointerp, tab, astride)
        g = DeviceGrid(f)
        g.sweep(dim, interp, tab, astride)
        out = g.get()
        assert relerr(out, ref) <= TOL, (kind, order, shape, dim, relerr(out, ref))


def test_presolve_matches_oracle_sol():
    import slb200 as S
    from oracle import refmodel as R

    rng = np.random.default_rng(SEED)
    for order, n in ((3, 32), (5, 128), (11, 256), (9, 1000)):
        b = np.asfortranarray(rng.random((n, 37)))
        x = S.sol(S.BSplineLU(order, n), b)
        oi = R.BSplineLU(order, n)
        ref = np.stack([oi.sol(b[:, j]) for j in range(b.shape[1])], axis=1)
        assert relerr(x, ref) <= TOL
        if n & (n - 1) == 0:
            xf = S.sol(S.BSplineFFT(order, n), b)
            of = R.BSplineFFT(order, n)
            reff = np.stack([of.sol(b[:, j]) for j in range(b.shape[1])], axis=1)
            assert relerr(xf, reff) <= TOL
    # Lagrange: sol is the identity (src/interpolation.jl:40)
    b = rng.random(50)
    assert np.array_equal(S.sol(S.Lagrange(5), b), b)


def test_edge_cases():
    import slb200 as S
    from oracle import refmodel as R

    rng = np.random.default_rng(SEED)
    # alpha = tiny negative -> decint = -1, decfloat == 1.0 (src/interpolation.jl:384-385); huge shifts;
    # exact integers; line shorter than the stencil (periodic wrap more than once)
    for n, order in ((128, 7), (5, 7), (3, 9), (31, 5), (2, 3)):
        f = np.asfortranarray(rng.random((n, 9)))
        tab = np.array([-1e-20, 0.0, 1.0, -3.0, 1e6 + 0.25, -1e6 - 0.75, 127.5, -0.5, 1e-300])
        interp, ointerp = make_pair("lagrange", order, n)
        ref = oracle_sweep(f, 0, ointerp, tab, [0, 1])
        g = DeviceGrid(f)
        g.sweep(0, interp, tab, [0, 1], flags=1)
        assert np.array_equal(g.get(), ref), (n, order)
        fT = np.asfortranarray(f.T)
        refT = oracle_sweep(fT, 1, ointerp, tab, [1, 0])
        g = DeviceGrid(fT)
        g.sweep(1, interp, tab, [1, 0], flags=1)
        assert np.array_equal(g.get(), refT), (n, order)
    # 1-D grid, single line
    f = rng.random(77)
    fp = np.empty(77)
    S.interpolate(fp, f, 2.625, S.Lagrange(9))
    ref = np.empty(77)
    R.interpolate(ref, f, 2.625, R.Lagrange(9))
    assert relerr(fp, ref) <= TOL
    # integer shift is an exact circular shift
    S.interpolate(fp, f, 5.0, S.Lagrange(7))
    assert np.array_equal(fp, np.roll(f, -5))


def test_error_behaviour():
    """Argument errors mirror the reference's exceptions (ValueError here)."""
    import slb200 as S

    with pytest.raises(ValueError):
        S.BSplineLU(4, 64)                      # src/bsplinelu.jl:257-261
    with pytest.raises(ValueError):
        S.BSplineFFT(5, 100)                    # src/fftbig.jl:57
    with pytest.raises(ValueError):
        S.Hermite(7)                            # src/hermite.jl:110
    m = S.UniformMesh(0.0, 1.0, 16)
    with pytest.raises(ValueError):
        S.Advection((m, m), [S.Lagrange(3)], 0.1, [([1, 2], 1, 1, True)])   # src/advection.jl:101-102
    adv = S.Advection((m, m), [S.Lagrange(3)] * 2, 0.1, [([1, 2], 1, 1, True), ([2, 1], 1, 2, True)])
    with pytest.raises(ValueError):
        S.AdvectionData(adv, np.zeros((16, 8)), S.gettranslationvar((1.0, 1.0)))  # src/advection.jl:251-253
    # B-spline object bound to another line length
    g = DeviceGrid(np.zeros((32, 4)))
    with pytest.raises(ValueError):
        g.sweep(0, S.BSplineLU(5, 64), np.zeros(1), [0, 0])
    # alpha table too short
    with pytest.raises(ValueError):
        g.sweep(0, S.Lagrange(5), np.zeros(2), [0, 1])


@pytest.mark.gpu
def test_inside_edge_matches_oracle():
    """InsideEdge Lagrange at the kernel seam (src/interpolation.jl:123-132, :250-286): SLB_SWEEP_EXACT is
    bitwise the oracle, the FMA form within 1e-12; the reference's cubic KAT holds on the device; lines
    along either axis; shifts that leave the array raise."""
    import slb200 as S
    from test_oracle_interp import TABDEC_INSIDE

    rng = np.random.default_rng(8)
    for order, n in ((3, 128), (7, 128), (5, 40), (9, 64)):
        its, ito = S.Lagrange(order, edge=S.InsideEdge), R.Lagrange(order, edge=R.InsideEdge)
        decs = [d for d in TABDEC_INSIDE if abs(math.floor(d)) <= order // 2 - 1] + [0.0, float(order // 2) - 0.25, -float(order // 2) + 0.5]
        f = np.asfortranarray(rng.random((n, len(decs))))
        ref = np.empty_like(f)
        for k, d in enumerate(decs):
            col = np.empty(n)
            R.interpolate(col, np.ascontiguousarray(f[:, k]), d, ito)
            ref[:, k] = col
        out = S.interpolate_lines(f, decs, its, flags=S.SLB_SWEEP_EXACT)
        assert np.array_equal(out, ref)
        out = S.interpolate_lines(f, decs, its)
        assert relerr(out, ref) <= TOL
        out_t = S.interpolate_lines(np.asfortranarray(f.T), decs, its, axis=1)  # lines along the strided axis
        assert relerr(out_t.T, ref) <= TOL
    mesh = np.arange(128) / 128
    cubic = lambda x: x**3 - x**2 - x / 6 + 0.25
    fp = np.empty(128)
    S.interpolate(fp, cubic(mesh), 3 / 1024, S.Lagrange(3, edge=S.InsideEdge))
    assert np.max(np.abs(fp - cubic(mesh + (3 / 1024) / 128))) < 1e-15   # test/test_interpolation.jl:488
    with pytest.raises(ValueError):
        S.interpolate(fp, cubic(mesh), 5.7, S.Lagrange(7, edge=S.InsideEdge))
    with pytest.raises(S.SlbError):
        S.interpolate_lines(np.zeros((64, 2), order="F"), [0.1, 0.2], S.BSplineLU(5, 64), flags=S.SLB_SWEEP_INSIDE_EDGE)


def test_inside_edge_is_a_kernel_seam_feature_only():
    import slb200 as S

    m = S.UniformMesh(0.0, 1.0, 16)
    with pytest.raises(ValueError):
        S.Advection((m, m), [S.Lagrange(3, edge=S.InsideEdge)] * 2, 0.1, [([1, 2], 1, 1, True), ([2, 1], 1, 2, True)])
