"""GPU parity of the unsplit 2-D path through the C ABI: the per-point N-D
interpolation kernel, the device array operations of the Adams-Bashforth time algorithms and the
single-state advection! driver, against the oracle on identical inputs.  Tolerance: 1e-12
relative max-abs per interpolation (north_star), 1e-10 on histories; SLB_SWEEP_EXACT is bitwise."""
import ctypes as C
import math

import numpy as np
import pytest

from helpers import make_pair, relerr
from oracle import refmodel as R, unsplit2d as U
from test_oracle_unsplit2d import interp2d_kat_inputs, poisson2d_run, swirling_setup

pytestmark = pytest.mark.gpu


def _pairs(spec, n1, n2):
    (ka, oa), (kb, ob) = spec
    a, ra = make_pair(ka, oa, n1)
    b, rb = make_pair(kb, ob, n2)
    return [a, b], [ra, rb]


SPECS = [
    ((("lagrange", 5), ("lagrange", 5)), 40, 33, 1),
    ((("lagrange", 7), ("lagrange", 7)), 64, 48, 2),
    ((("lagrange", 9), ("lagrange", 9)), 100, 100, 1),
    ((("lagrange", 3), ("lagrange", 9)), 21, 20, 1),       # different orders: run-time-order kernel
    ((("lagrange", 17), ("lagrange", 4)), 30, 19, 2),      # order + 1 > 14
    ((("hermite", 5), ("hermite", 5)), 32, 40, 1),
    ((("hermite", 9), ("lagrange", 9)), 50, 24, 2),
    ((("bspline_lu", 5), ("bspline_lu", 5)), 128, 256, 1),
    ((("bspline_lu", 9), ("bspline_lu", 9)), 100, 100, 2),
    ((("bspline_fft", 5), ("bspline_fft", 5)), 128, 64, 1),
    ((("bspline_lu", 3), ("lagrange", 3)), 36, 20, 2),     # pre-solve along dim 1 only
    ((("lagrange", 5), ("bspline_lu", 5)), 20, 36, 1),     # pre-solve along dim 2 only
]


@pytest.mark.parametrize("spec,n1,n2,ncomp", SPECS)
def test_interp2d_points_matches_oracle(spec, n1, n2, ncomp):
    import slb200 as S

    its, rits = _pairs(spec, n1, n2)
    rng = np.random.default_rng(20240611)
    f = np.asfortranarray(rng.random((n1, n2, ncomp)))
    dec = np.asfortranarray(rng.uniform(-8, 8, (n1, n2, 2)))
    dec[0, 0, :] = (3.0, -2.0)
    dec[1, 0, :] = (1e6 + 0.25, -1e6 - 0.75)
    fin = f if ncomp > 1 else np.asfortranarray(f[:, :, 0])
    ref = U.interpolate_points(fin, dec, rits, nthreads=4)
    plain = all(k in ("lagrange", "hermite") for k, _ in spec)
    for flags in (0, S.SLB_SWEEP_EXACT):
        out = np.empty_like(fin)
        S.interpolate_nd(out, fin, dec, its, flags=flags)
        if flags and plain:
            assert np.array_equal(out, ref)
        else:
            assert relerr(out, ref) <= 1e-12


def test_interp2d_points_function_form_and_errors():
    import slb200 as S

    rng = np.random.default_rng(3)
    n1, n2 = 12, 9
    f = np.asfortranarray(rng.random((n1, n2)))
    dec = np.asfortranarray(rng.uniform(-4, 4, (n1, n2, 2)))
    its, rits = _pairs((("lagrange", 3), ("lagrange", 5)), n1, n2)
    a, b = np.empty_like(f), np.empty_like(f)
    S.interpolate_nd(a, f, dec, its)
    S.interpolate_nd(b, f, lambda ind: (dec[ind[0], ind[1], 0], dec[ind[0], ind[1], 1]), its)
    assert np.array_equal(a, b)
    assert relerr(a, U.interpolate_points(f, dec, rits)) <= 1e-12
    with pytest.raises(ValueError):
        S.interpolate_nd(f, f, dec, its)  # fp and fi must not alias
    with pytest.raises(ValueError):
        S.interpolate_nd(a, f, dec, its[:1])
    with pytest.raises(ValueError):  # a B-spline object bound to another line length
        S.interpolate_nd(a, f, dec, [S.BSplineLU(3, 16), S.Lagrange(3)])


def test_interp2d_reference_kat():
    """test/test_interpolation.jl:120-178, :486 (test_interp2d, Lagrange 11, (128, 100)) on the device: scalar
    and two-component fields, array and function form, within 1000 eps of each other and 1e-12 of the
    oracle"""
    import slb200 as S

    prec = 1000 * np.finfo(np.float64).eps
    dec, ref, ref2, op = interp2d_kat_inputs()
    its, rits = _pairs((("lagrange", 11), ("lagrange", 11)), 128, 100)
    res1, res3, res4 = np.empty_like(ref), np.empty_like(ref), np.empty_like(ref)
    S.interpolate_nd(res1, ref, dec, its)
    S.interpolate_nd(res3, ref, lambda ind: (dec[ind[0], ind[1], 0], dec[ind[0], ind[1], 1]), its)
    S.interpolate_nd(res4, ref2, dec, its)
    assert np.linalg.norm(res1 - res3) < prec
    both = np.empty(ref.shape + (2,), order="F")
    both[:, :, 0], both[:, :, 1] = ref, ref2
    r = np.empty_like(both)
    S.interpolate_nd(r, both, dec, its)
    assert np.linalg.norm(r[:, :, 0] - res1) < prec and np.linalg.norm(r[:, :, 1] - res4) < prec
    opres = np.empty_like(op)
    S.interpolate_nd(opres, op, dec, its)
    assert relerr(opres, U.interpolate_points(op, dec, rits)) <= 1e-12
    assert relerr(res1, U.interpolate_points(ref, dec, rits)) <= 1e-12


def test_device_array_operations_are_bitwise():
    """slb_lincomb (rounded products summed left to right), slb_fill_dec2d, slb_memcpy_d2d"""
    import slb200 as S
    from slb200 import _lib, unsplit2d as D

    ctx = _lib.default_context()
    rng = np.random.default_rng(9)
    n1, n2 = 37, 23
    hs = [np.asfortranarray(rng.standard_normal((n1, n2, 2))) for _ in range(5)]
    fs = [S.DeviceField.from_host(ctx, h) for h in hs]
    coefs = [55 / 24, -59 / 24, 37 / 24, -3 / 8, 0.1]
    for n in (1, 2, 4, 5):
        out = D.lincomb(fs[0].like(), coefs[:n], fs[:n])
        acc = coefs[0] * hs[0]
        for k in range(1, n):
            acc = acc + coefs[k] * hs[k]
        assert np.array_equal(out.to_host(), acc)
        out.free()
    cp = fs[1].copy()
    assert np.array_equal(cp.to_host(), hs[1])
    tj, ti = rng.standard_normal(n2), rng.standard_normal(n1)
    dj, di = ctx.to_device(tj), ctx.to_device(ti)
    dec = S.DeviceField(ctx, n1, n2, 2)
    _lib.check(_lib.lib().slb_fill_dec2d(ctx.h, dec.ptr, n1, n2, dj, -0.37, di, 1.9))
    got = dec.to_host()
    assert np.array_equal(got[:, :, 0], np.broadcast_to((-0.37 * tj)[None, :], (n1, n2)))
    assert np.array_equal(got[:, :, 1], np.broadcast_to((1.9 * ti)[:, None], (n1, n2)))
    with pytest.raises(ValueError):
        _lib.check(_lib.lib().slb_lincomb(ctx.h, dec.ptr, 9, (C.c_double * 9)(), (C.c_void_p * 9)(), 10))
    ctx.free(dj)
    ctx.free(di)


class SwirlingOracle:
    """test/test_swirling.jl:153-167"""

    def __init__(self, dec):
        self.ref = dec.copy(order="F")

    def initcoef(self, advd):
        coef = advd.adv.dt_base * math.cos(math.pi * advd.time_cur / 1.5)
        if advd.bufcur is None:
            advd.bufcur = np.zeros(advd.adv.sizeall + (2,), order="F")
        advd.bufcur[...] = coef * self.ref


def _swirling_device_provider(dec):
    import slb200 as S
    from slb200 import unsplit2d as D

    class SwirlingDev(S.AbstractExtDataAdv):
        """the same user-defined provider on the product side: the reference field stays on the device"""

        def __init__(self):
            self.ref = None

        def initcoef(self, advd):
            coef = advd.adv.dt_base * math.cos(math.pi * advd.time_cur / 1.5)
            if self.ref is None:
                self.ref = S.DeviceField.from_host(advd.ctx, dec)
            if advd.bufcur is None:
                advd.bufcur = self.ref.like()
            D.lincomb(advd.bufcur, [coef], [self.ref])

    return SwirlingDev()


def _swirling_both(spec, nbdt, nsteps, timealg_name="NoTimeAlg", ordalg=0, sz=(100, 100)):
    import slb200 as S

    its, rits = _pairs(spec, *sz)
    dt = 1.5 / nbdt
    out = []
    for M, interps, prov in ((S, its, None), (R, rits, None)):
        mesh_sp, mesh_v, dec, tabref = swirling_setup(M, sz)
        alg = getattr(M, timealg_name)
        kw = {} if M is S else {"nthreads": 4}
        adv = M.Advection((mesh_sp, mesh_v), interps, dt, [([1, 2], 2, 1, False)], tab_coef=[dt], timealg=alg, ordalg=ordalg, **kw)
        initdatas = [tabref.copy(order="F") for _ in range(3 * ordalg - 1)] if timealg_name == "ABTimeAlg_init" else None
        prov = _swirling_device_provider(dec) if M is S else SwirlingOracle(dec)
        advd = M.AdvectionData(adv, tabref, prov, initdatas=initdatas)
        if initdatas is not None:
            advd.time_cur -= len(initdatas) * dt
        out.append((M, advd, tabref))
    (_, g, tabref), (_, o, _) = out
    worst = 0.0
    for _ in range(nsteps):
        while S.advection(g):
            pass
        while R.advection(o):
            pass
        assert abs(g.time_cur - o.time_cur) <= 1e-14
        worst = max(worst, relerr(g.getdata(), o.data))
    return worst, g, tabref


@pytest.mark.parametrize("spec", [(("lagrange", 9), ("lagrange", 9)), (("bspline_lu", 9), ("bspline_lu", 9)), (("hermite", 9), ("hermite", 9))])
def test_swirling_advection_matches_oracle_and_reference_kat(spec):
    """test/test_swirling.jl:246-261: full 50-step deformation flow; every step within 1e-12 of the
    oracle (errors of earlier steps are carried along: 1e-11 over the run) and the reference's own
    criterion (back to the start within 5 in the 2-norm)"""
    worst, g, tabref = _swirling_both(spec, 50, 50)
    assert worst <= 1e-11
    assert float(np.linalg.norm(g.getdata() - tabref)) < 5


def test_swirling_abtimealg_init_matches_oracle_and_reference_kat():
    """test/test_swirling.jl:263-270: ABTimeAlg_init, ordalg = 4: < 2"""
    worst, g, tabref = _swirling_both((("lagrange", 9), ("lagrange", 9)), 50, 50, "ABTimeAlg_init", 4)
    assert worst <= 1e-10
    assert float(np.linalg.norm(g.getdata() - tabref)) < 2


@pytest.mark.parametrize("alg,ordalg,kind,order", [("ABTimeAlg_ip", 2, "lagrange", 7), ("ABTimeAlg_ip", 3, "lagrange", 7),
                                                 ("ABTimeAlg_ip", 4, "lagrange", 7), ("ABTimeAlg_new", 2, "lagrange", 7),
                                                 ("ABTimeAlg_new", 2, "bspline_lu", 11), ("NoTimeAlg", 0, "lagrange", 7)])
def test_poisson2d_unsplit_matches_oracle(alg, ordalg, kind, order):
    """test/test_poisson2d.jl:178-258 with StdPoisson2d: data, energies and the energy drift of the
    unsplit Vlasov-Poisson solver against the oracle"""
    import slb200 as S

    sz = (128, 100)
    its, rits = _pairs(((kind, order), (kind, order)), *sz)
    dg, g = poisson2d_run(S, lambda adv: S.getpoissonvar(adv, type=S.StdPoisson2d), sz, its, 0.1, 5, getattr(S, alg), ordalg)
    do, o = poisson2d_run(R, U.getpoissonvar2d, sz, rits, 0.1, 5, getattr(R, alg), ordalg, nthreads=4)
    assert relerr(g.getdata(), o.data) <= 1e-10
    eg, eo = S.getenergy(g), R.getenergy(o)
    for a, b in zip(eg, eo):
        assert abs(a - b) <= 1e-10 * abs(b)
    assert abs(dg - do) <= 1e-10 * max(abs(eo[2]), 1.0)


def test_poisson2d_time_algorithm_order_on_device():
    """test/test_poisson2d.jl:353-379 (test_timealg, ABTimeAlg_ip order 3): halving dt divides the
    energy drift by 2^ordalg"""
    import slb200 as S

    rets = []
    for nbdt in (5, 10):
        its = [S.Lagrange(7), S.Lagrange(7)]
        r, _ = poisson2d_run(S, lambda adv: S.getpoissonvar(adv, type=S.StdPoisson2d), (128, 100), its, 0.1, nbdt, S.ABTimeAlg_ip, 3)
        rets.append(r)
    assert 1.25 * rets[0] / rets[1] > 2**3, rets


def test_rotation2d_abtimealg_matches_oracle():
    """test/test_rotation.jl:101-146: unsplit rotation with ABTimeAlg_ip"""
    import slb200 as S

    nbdt, ordalg = 40, 3
    dt = 2 * math.pi / nbdt
    res = []
    for M in (S, R):
        mesh_sp, mesh_v = M.UniformMesh(-5.0, 5.0, 200), M.UniformMesh(-5.0, 5.0, 102)
        interps = [M.Lagrange(9), M.Lagrange(9)]
        kw = {} if M is S else {"nthreads": 4}
        adv = M.Advection((mesh_sp, mesh_v), interps, dt, [([1, 2], 2, 1, False)], tab_coef=M.nosplit(dt), timealg=M.ABTimeAlg_ip,
                          ordalg=ordalg, **kw)
        x = mesh_sp.points[:, None]
        y = mesh_v.points[None, :]
        f0 = np.asfortranarray(np.exp(-2 * (x**2 + (y + 6 / 5) ** 2)))
        pv = S.getrotationvar(adv) if M is S else U.getrotationvar2d(adv)
        advd = M.AdvectionData(adv, f0, pv)
        if M is S:
            # the rotation field is discontinuous across the periodic boundary, where a last-bit change
            # of a displacement can move its floor(): compare in the reference's operation order
            advd.flags = S.SLB_SWEEP_EXACT
        for _ in range(10):
            while M.advection(advd):
                pass
        res.append(advd)
    assert np.array_equal(res[0].getdata(), res[1].data)
    assert np.array_equal(res[0].bufcur.to_host(), res[1].bufcur)


def test_split_states_reject_time_algorithms():
    import slb200 as S

    ms = (S.UniformMesh(0.0, 1.0, 16), S.UniformMesh(0.0, 1.0, 16))
    adv = S.Advection(ms, [S.Lagrange(3), S.Lagrange(3)], 0.1, [([1, 2], 1, 1, True), ([2, 1], 1, 2, True)], timealg=S.ABTimeAlg_ip, ordalg=2)
    advd = S.AdvectionData(adv, np.zeros((16, 16)), S.gettranslationvar((1.0, 1.0)))
    with pytest.raises(NotImplementedError):
        S.advection(advd)


@pytest.mark.parametrize("alg,ordalg", [("NoTimeAlg", 0), ("ABTimeAlg_ip", 2), ("ABTimeAlg_ip", 3)])
def test_quasigeostrophic_matches_oracle(alg, ordalg):
    """test/test_quasigeostrophic.jl:25-96: the SQG provider (velocity = spectral multiplier of the advected
    field, on the library's DFT kernels) with the unsplit driver, 10 steps on 128 x 128, against the oracle"""
    import slb200 as S
    from test_oracle_unsplit2d import sqg_run

    g = sqg_run(S, S.getgeovar, 10, getattr(S, alg), ordalg)
    o = sqg_run(R, U.getgeovar, 10, getattr(R, alg), ordalg, nthreads=4)
    assert abs(g.time_cur - o.time_cur) <= 1e-9
    assert relerr(g.getdata(), o.data) <= 1e-10


def test_quasigeostrophic_order_on_device():
    """test/test_quasigeostrophic.jl:146-161, :189 (test_orderno, ABTimeAlg_ip 2)"""
    import slb200 as S
    from test_oracle_unsplit2d import sqg_run

    d4, d1, d2 = (sqg_run(S, S.getgeovar, n, S.ABTimeAlg_ip, 2).getdata() for n in (40, 10, 20))
    ret1, ret2 = np.linalg.norm(d4 - d1), np.linalg.norm(d4 - d2)
    assert ret1 * 1.2 / ret2 > 2**2, (ret1, ret2)


@pytest.mark.parametrize("split,exact", [("standardsplit", False), ("strangsplit", False), ("strangsplit", True)])
def test_quasigeostrophic_split_form_matches_oracle(split, exact):
    """split states with per-point shifts, [([1, 2], 1, 1, false), ([2, 1], 1, 2, false)] (the split form of the SQG
    driver: src/advection.jl:633-645, src/quasigeostrophic.jl:126-135, test/test_quasigeostrophic.jl:45-58), stage by
    stage against the oracle's line-by-line restatement"""
    import slb200 as S
    from oracle import refmodel as R
    from oracle import unsplit2d as U

    def build(M, getgeovar, splitf):
        mx, my = M.UniformMesh(0.0, 1e6, 48), M.UniformMesh(0.0, 1e6, 40)
        dt = 10000.0 / 4
        adv = M.Advection((mx, my), [M.Lagrange(9), M.Lagrange(9)], dt, [([1, 2], 1, 1, False), ([2, 1], 1, 2, False)], tab_coef=splitf(dt))
        pv = getgeovar(adv)
        advd = M.AdvectionData(adv, np.zeros((48, 40)), pv)
        pv.initdata(advd)
        return advd

    g = build(S, S.getgeovar, getattr(S, split))
    o = build(R, U.getgeovar, getattr(R, split))
    if exact:
        g.flags = S.SLB_SWEEP_EXACT
    for _ in range(2):
        more = True
        while more:
            more = S.advection(g)
            assert more == R.advection(o)
            a, b = g.getdata(), o.data
            assert float(np.max(np.abs(a - b)) / np.max(np.abs(b))) <= (1e-13 if exact else 1e-12)
