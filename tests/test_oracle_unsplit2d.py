"""Pins the ORACLE's restatement of the unsplit 2-D path (per-point N-D interpolation, the
Adams-Bashforth time algorithms, StdPoisson2d / rotation / user-defined providers) to the
reference's own known-answer tests.  CPU only."""
import math
import os
import ctypes as C

import numpy as np
import pytest

from oracle import refmodel as R, tables, unsplit2d as U

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abcoef_known_answers():
    """ABcoef (src/lagrange.jl:74-88) = the classical Adams-Bashforth weights"""
    from fractions import Fraction as F

    t = tables.abcoef_rat(5)
    cols = [[F(1)], [F(3, 2), F(-1, 2)], [F(23, 12), F(-4, 3), F(5, 12)], [F(55, 24), F(-59, 24), F(37, 24), F(-3, 8)],
            [F(1901, 720), F(-1387, 360), F(109, 30), F(-637, 360), F(251, 720)]]
    for j, col in enumerate(cols):
        assert [t[i][j] for i in range(j + 1)] == col
        assert sum(col) == 1


def test_constant_shift_field_equals_the_tensor_stencil_and_two_sweeps():
    """a per-point field that happens to be constant reproduces the const-shift N-D interpolate!
    (src/interpolation.jl:212-231) and the product of two 1-D shifts"""
    rng = np.random.default_rng(5)
    n1, n2 = 24, 20
    f = np.asfortranarray(rng.random((n1, n2)))
    for interps in ([R.Lagrange(5), R.Lagrange(5)], [R.Lagrange(3), R.Lagrange(7)], [R.BSplineLU(5, n1), R.BSplineLU(3, n2)],
                    [R.Hermite(5), R.Lagrange(4)]):
        a = (2.37, -5.81)
        dec = np.empty((n1, n2, 2), order="F")
        dec[:, :, 0], dec[:, :, 1] = a
        got = U.interpolate_points(f, dec, interps)
        ref = R.interpolate_nd_const(f, a, interps)
        assert np.max(np.abs(got - ref)) <= 1e-14 * np.max(np.abs(ref))
        # two 1-D shifts
        tmp = np.empty_like(f)
        for j in range(n2):
            col = np.empty(n1)
            R.interpolate(col, np.ascontiguousarray(f[:, j]), a[0], interps[0])
            tmp[:, j] = col
        two = np.empty_like(f)
        for i in range(n1):
            row = np.empty(n2)
            R.interpolate(row, np.ascontiguousarray(tmp[i, :]), a[1], interps[1])
            two[i, :] = row
        assert np.max(np.abs(got - two)) <= 1e-13 * np.max(np.abs(two))


def test_function_form_equals_array_form():
    rng = np.random.default_rng(6)
    n1, n2 = 12, 9
    f = np.asfortranarray(rng.random((n1, n2)))
    dec = np.asfortranarray(rng.uniform(-4, 4, (n1, n2, 2)))
    interps = [R.Lagrange(3), R.Lagrange(5)]
    a = U.interpolate_points(f, dec, interps)
    b = U.interpolate_fct(f, lambda ind: (dec[ind[0], ind[1], 0], dec[ind[0], ind[1], 1]), interps)
    assert np.array_equal(a, b)


def interp2d_kat_inputs(sz=(128, 100), coeff=1.25):
    """the fields of test/test_interpolation.jl:120-150 (test_interp2d): a smooth displacement field
    of amplitude 1.25 cells, two scalar fields and a two-component field"""
    i = np.arange(1, sz[0] + 1)[:, None] / sz[0]
    j = np.arange(1, sz[1] + 1)[None, :] / sz[1]
    x = 2 * math.pi * (i + j)
    dec = np.empty(sz + (2,), order="F")
    dec[:, :, 0], dec[:, :, 1] = coeff * np.cos(x + 1), coeff * np.cos(x + 2)
    ref = np.asfortranarray(np.sin(x + 3) + np.cos(x - 1))
    ref2 = np.asfortranarray(np.sin(x + 1) + 3 * np.cos(-x + 2) / 5)
    op = np.empty(sz + (2,), order="F")
    op[:, :, 0], op[:, :, 1] = coeff * np.sin(x + 4), coeff * np.cos(x + 5)
    return dec, ref, ref2, op


def test_interp2d_forms_agree():
    """test/test_interpolation.jl:120-178, :486 (test_interp2d, Lagrange 11, (128, 100)): the array form,
    the function form and the component-wise interpolation of a two-component field agree to 1000 eps
    (they are the same arithmetic here: exactly)"""
    prec = 1000 * np.finfo(np.float64).eps
    dec, ref, ref2, op = interp2d_kat_inputs()
    interps = [R.Lagrange(11), R.Lagrange(11)]
    res1 = U.interpolate_points(ref, dec, interps)
    res3 = U.interpolate_fct(ref, lambda ind: (dec[ind[0], ind[1], 0], dec[ind[0], ind[1], 1]), interps)
    res4 = U.interpolate_points(ref2, dec, interps)
    assert np.linalg.norm(res1 - res3) < prec
    both = np.empty(ref.shape + (2,), order="F")
    both[:, :, 0], both[:, :, 1] = ref, ref2
    r = U.interpolate_points(both, dec, interps)
    assert np.linalg.norm(r[:, :, 0] - res1) < prec and np.linalg.norm(r[:, :, 1] - res4) < prec
    opres = U.interpolate_points(op, dec, interps)
    assert np.linalg.norm(opres[:, :, 0] - U.interpolate_points(np.asfortranarray(op[:, :, 0]), dec, interps)) < prec
    # and the result is the field moved along the displacement: compare with the analytic composition
    i = np.arange(1, 129)[:, None]
    j = np.arange(1, 101)[None, :]
    xs = 2 * math.pi * ((i + dec[:, :, 0]) / 128 + (j + dec[:, :, 1]) / 100)
    assert np.max(np.abs(res1 - (np.sin(xs + 3) + np.cos(xs - 1)))) < 1e-9


def test_host_build_of_the_cuda_point_body_matches_the_oracle():
    """slb_point_eval (the per-thread body of k_interp2d_points) compiled for the host: EXACT mode is
    bit-identical to the oracle, the FMA mode agrees to rounding; templated and run-time-order
    instantiations agree with each other."""
    so = os.path.join(ROOT, "semilagrangian.jl_b200", "lib", "libslb200_hosttest.so")
    assert os.path.exists(so), "build first (__graft_entry__.build())"
    L = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    L.slbt_points_host.argtypes = [dp, dp, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, dp, C.c_int, C.c_int, dp, C.c_int, C.c_int]
    rng = np.random.default_rng(11)
    for (oa, ob, n1, n2, ncomp) in [(5, 5, 40, 33, 1), (7, 7, 16, 18, 2), (3, 9, 21, 20, 1), (9, 9, 10, 10, 2), (17, 4, 30, 19, 1)]:
        ia, ib = R.Lagrange(oa), R.Lagrange(ob)
        f = np.asfortranarray(rng.random((n1, n2, ncomp)))
        dec = np.asfortranarray(rng.uniform(-7.5, 7.5, (n1, n2, 2)))
        dec[0, 0, :] = (3.0, -2.0)  # integer shifts: decfloat == 0
        dec[1, 0, :] = (1e6 + 0.25, -1e6 - 0.75)
        ref = U.interpolate_points(f if ncomp > 1 else f[:, :, 0], dec, [ia, ib]).reshape((n1, n2, ncomp), order="F")
        ca, cb = np.ascontiguousarray(ia.tabfct), np.ascontiguousarray(ib.tabfct)
        outs = {}
        for exact in (1, 0):
            for templ in ((1, 0) if (oa == ob and oa + 1 <= 14) else (0,)):
                out = np.empty_like(f, order="F")
                rc = L.slbt_points_host(f.ctypes.data_as(dp), dec.ctypes.data_as(dp), out.ctypes.data_as(dp), n1, n2, ncomp,
                                        oa + 1, ca.shape[1], ca.ctypes.data_as(dp), ob + 1, cb.shape[1], cb.ctypes.data_as(dp), exact, templ)
                assert rc == 0
                outs[(exact, templ)] = out
                if exact:
                    assert np.array_equal(out, ref), (oa, ob, templ)
                else:
                    assert np.max(np.abs(out - ref)) <= 1e-13 * np.max(np.abs(ref))
        if (1, 1) in outs:
            assert np.array_equal(outs[(1, 1)], outs[(1, 0)]) and np.array_equal(outs[(0, 1)], outs[(0, 0)])


def _rotation_points(sz, interps, nbdt):
    """test/test_rotation.jl:148-206: rotation by per-point interpolation with the exact
    characteristic feet; returns the max error against the analytic solution"""
    mesh_sp, mesh_v = R.UniformMesh(-4.5, 5.0, sz[0]), R.UniformMesh(-5.2, 5.0, sz[1])
    dt = 2 * math.pi / nbdt
    tgdt = 2 * math.tan(dt / 2)
    coef = 1 / (1 + tgdt**2 / 4)
    x = mesh_sp.points[:, None]
    y = mesh_v.points[None, :]
    dec = np.empty(sz + (2,), order="F")
    dec[:, :, 0] = -tgdt * coef * (y + tgdt * x / 2) / mesh_sp.step
    dec[:, :, 1] = (tgdt * coef * (x - tgdt * y / 2)) / mesh_v.step

    def exact(tf):
        s, c = math.sin(tf), math.cos(tf)
        xn, yn = c * x - s * y, s * x + c * y
        return np.asfortranarray(np.exp(-2 * (xn**2 + (yn + 6 / 5) ** 2)))

    data = exact(0.0)
    diffmax = 0.0
    for ind in range(1, nbdt + 1):
        data = U.interpolate_points(data, dec, interps, nthreads=4)
        diffmax = max(diffmax, float(np.max(np.abs(data - exact(dt * ind)))))
    return diffmax


def test_rotation_by_per_point_interpolation():
    """test/test_rotation.jl:209 (the per-point variant of the rotation test): < 1e-3"""
    assert _rotation_points((100, 122), [R.Lagrange(5), R.Lagrange(5)], 11) < 1e-3
    assert _rotation_points((128, 256), [R.BSplineLU(5, 128), R.BSplineLU(5, 256)], 11) < 1e-3


class Swirling:
    """the user-defined provider of test/test_swirling.jl:153-167"""

    def __init__(self, dec):
        self.ref = dec.copy(order="F")

    def initcoef(self, advd):
        coef = advd.adv.dt_base * math.cos(math.pi * advd.time_cur / 1.5)
        if advd.bufcur is None:
            advd.bufcur = np.zeros(advd.adv.sizeall + (2,), order="F")
        advd.bufcur[...] = coef * self.ref


def swirling_setup(M, sz):
    mesh_sp, mesh_v = M.UniformMesh(0.0, 1.0, sz[0]), M.UniformMesh(0.0, 1.0, sz[1])
    x = mesh_sp.points[:, None]
    y = mesh_v.points[None, :]
    dec = np.empty(sz + (2,), order="F")
    dec[:, :, 0] = -np.cos(math.pi * x) ** 2 * np.sin(2 * math.pi * y) / mesh_sp.step
    dec[:, :, 1] = np.cos(math.pi * y) ** 2 * np.sin(2 * math.pi * x) / mesh_v.step
    tabref = np.asfortranarray(np.exp(-10 * ((1 / 4 - x) ** 2 + (1 / 2 - y) ** 2)))
    return mesh_sp, mesh_v, dec, tabref


def _swirling_adv(sz, interps, nbdt, timealg=R.NoTimeAlg, ordalg=0):
    """test/test_swirling.jl:169-240"""
    mesh_sp, mesh_v, dec, tabref = swirling_setup(R, sz)
    dt = 1.5 / nbdt
    adv = R.Advection((mesh_sp, mesh_v), interps, dt, [([1, 2], 2, 1, False)], tab_coef=[dt], timealg=timealg, ordalg=ordalg, nthreads=4)
    initdatas = [tabref.copy(order="F") for _ in range(3 * ordalg - 1)] if timealg == R.ABTimeAlg_init else None
    advd = R.AdvectionData(adv, tabref, Swirling(dec), initdatas=initdatas)
    if timealg == R.ABTimeAlg_init:
        advd.time_cur -= len(initdatas) * dt
    for _ in range(nbdt):
        while R.advection(advd):
            pass
    return float(np.linalg.norm(advd.data - tabref))


def test_swirling_interpolate_only():
    """test/test_swirling.jl:60-151, :244: the deformation flow by bare per-point interpolate!: < 15"""
    sz, nbdt = (100, 100), 50
    mesh_sp, mesh_v, dec, _ = swirling_setup(R, sz)
    x = mesh_sp.points[:, None]
    y = mesh_v.points[None, :]
    tabref = np.asfortranarray(np.where((1 - x) ** 2 + (1 - y) ** 2 < 0.8, 1.0, 0.0) + 0 * y)
    dt = 1.5 / nbdt
    data = tabref.copy(order="F")
    interps = [R.Lagrange(9), R.Lagrange(9)]
    for ind in range(1, nbdt + 1):
        coef = dt * math.cos(math.pi * (ind - 1) / nbdt)
        data = U.interpolate_points(data, coef * dec, interps, nthreads=4)
    assert float(np.linalg.norm(data - tabref)) < 15


@pytest.mark.parametrize("kind", ["lagrange", "bsplinelu", "hermite"])
def test_swirling_advection(kind):
    """test/test_swirling.jl:246-261: returns to the start within 5 (2-norm)"""
    interps = {"lagrange": lambda: [R.Lagrange(9), R.Lagrange(9)], "bsplinelu": lambda: [R.BSplineLU(9, 100), R.BSplineLU(9, 100)],
               "hermite": lambda: [R.Hermite(9), R.Hermite(9)]}[kind]()
    assert _swirling_adv((100, 100), interps, 50) < 5


def test_swirling_advection_abtimealg_init():
    """test/test_swirling.jl:263-270: ABTimeAlg_init of order 4: < 2"""
    assert _swirling_adv((100, 100), [R.Lagrange(9), R.Lagrange(9)], 50, timealg=R.ABTimeAlg_init, ordalg=4) < 2


def poisson2d_run(M, getpv, sz, interps, t_max, nbdt, timealg, ordalg, advance=None, **kw):
    """test/test_poisson2d.jl:178-258 (test_poisson2dadv) for the time algorithms that need no
    start-up data: returns (enmax - enmin, advd)"""
    mesh_sp, mesh_v = M.UniformMesh(0.0, 4 * math.pi, sz[0]), M.UniformMesh(-9.0, 9.0, sz[1])
    dt = t_max / nbdt
    adv = M.Advection((mesh_sp, mesh_v), interps, dt, [([1, 2], 2, 1, False)], tab_coef=M.nosplit(dt), timealg=timealg, ordalg=ordalg, **kw)
    x = mesh_sp.points[:, None]
    y = mesh_v.points[None, :]
    data = np.asfortranarray(1 / math.sqrt(2 * math.pi) * np.exp(-0.5 * y**2) * (1 + 0.5 * np.cos(x / 2)))
    advd = M.AdvectionData(adv, data, getpv(adv))
    en = [M.getenergy(advd)[2]]
    borne_t = t_max - dt / 2
    while advd.time_cur < borne_t:
        while M.advection(advd):
            pass
        en.append(M.getenergy(advd)[2])
    return max(en) - min(en), advd


@pytest.mark.parametrize("timealg,ordalg", [(R.ABTimeAlg_ip, 2), (R.ABTimeAlg_ip, 3), (R.ABTimeAlg_new, 2)])
def test_poisson2d_time_algorithm_order(timealg, ordalg):
    """test/test_poisson2d.jl:353-393 (test_timealg): halving dt divides the energy drift by
    2^ordalg (Float64 here, Double64 there)"""
    interps = [R.Lagrange(7), R.Lagrange(7)]
    ret1, _ = poisson2d_run(R, U.getpoissonvar2d, (128, 100), interps, 0.1, 5, timealg, ordalg, nthreads=4)
    interps = [R.Lagrange(7), R.Lagrange(7)]
    ret2, _ = poisson2d_run(R, U.getpoissonvar2d, (128, 100), interps, 0.1, 10, timealg, ordalg, nthreads=4)
    assert 1.25 * ret1 / ret2 > 2**ordalg, (ret1, ret2)


def test_rotation2d_abtimealg_converges():
    """test/test_rotation.jl:101-146 (test_rotation2d, ABTimeAlg_ip): the unsplit rotation tracks the
    analytic solution and halving dt improves it by about 2^ordalg"""
    def run(nbdt, ordalg):
        mesh_sp, mesh_v = R.UniformMesh(-5.0, 5.0, 200), R.UniformMesh(-5.0, 5.0, 102)
        dt = 2 * math.pi / nbdt
        interps = [R.Lagrange(9), R.Lagrange(9)]
        adv = R.Advection((mesh_sp, mesh_v), interps, dt, [([1, 2], 2, 1, False)], tab_coef=R.nosplit(dt), timealg=R.ABTimeAlg_ip,
                          ordalg=ordalg, nthreads=4)
        x = mesh_sp.points[:, None]
        y = mesh_v.points[None, :]

        def exact(tf):
            s, c = math.sin(tf), math.cos(tf)
            xn, yn = c * x - s * y, s * x + c * y
            return np.asfortranarray(np.exp(-2 * (xn**2 + (yn + 6 / 5) ** 2)))

        advd = R.AdvectionData(adv, exact(0.0), U.getrotationvar2d(adv))
        diffmax = 0.0
        for ind in range(1, nbdt + 1):
            while R.advection(advd):
                pass
            diffmax = max(diffmax, float(np.max(np.abs(advd.data - exact(dt * ind)))))
        return diffmax

    # the reference's (disabled) version compares 20 and 40 steps per turn, where the scheme is not yet
    # in its asymptotic regime; 80 / 160 steps show the order cleanly
    r1, r2 = run(80, 2), run(160, 2)
    assert r2 < (r1 * 1.1) / 4 and r1 < 0.1, (r1, r2)


def sqg_run(M, getgv, nbdt, timealg, ordalg, sz=(128, 128), t_max=10000.0, **kw):
    """test/test_quasigeostrophic.jl:25-96 (test_quasigeostrophic, nosplit): returns the final data"""
    mx, my = M.UniformMesh(0.0, 1e6, sz[0]), M.UniformMesh(0.0, 1e6, sz[1])
    dt = t_max / nbdt
    adv = M.Advection((mx, my), [M.Lagrange(9), M.Lagrange(9)], dt, [([1, 2], 2, 1, False)], tab_coef=M.nosplit(dt), timealg=timealg,
                      ordalg=ordalg, **kw)
    pv = getgv(adv)
    advd = M.AdvectionData(adv, np.zeros(sz, order="F"), pv)
    pv.initdata(advd)
    borne_t = t_max - adv.dt_base / 2
    while advd.time_cur < borne_t:
        while M.advection(advd):
            pass
    return advd


@pytest.mark.parametrize("timealg,ordalg", [(R.NoTimeAlg, 0), (R.ABTimeAlg_ip, 2), (R.ABTimeAlg_ip, 3)])
def test_quasigeostrophic_order(timealg, ordalg):
    """test/test_quasigeostrophic.jl:146-161, :188-190 (test_orderno): halving dt divides the distance to the
    4x finer run by 2^ord (Float64 here, Double64 there)"""
    d4, d1, d2 = (np.array(sqg_run(R, U.getgeovar, n, timealg, ordalg, nthreads=4).data) for n in (40, 10, 20))
    ret1, ret2 = np.linalg.norm(d4 - d1), np.linalg.norm(d4 - d2)
    assert ret1 * 1.2 / ret2 > 2 ** (ordalg if ordalg else 1), (ret1, ret2)


def test_poisson_3d3v_consistency():
    """test/test_poisson.jl:38-108 (test_poisson, 3D3V, Lagrange 3, random data): after initcoef! at the first
    velocity state, rho equals the direct charge integral, every E component the direct spectral formula
    (src/util_poisson.jl:84-119) and compute_ee / compute_ke their definitions"""
    ms = [R.UniformMesh(a, b, n) for a, b, n in ((-1, 3, 8), (-10, 6, 4), (-3, 5, 16), (-3, 1, 4), (-9, 7, 8), (1, 5, 4))]
    interp = [R.Lagrange(3) for _ in range(6)]
    adv = R.Advection(tuple(ms), interp, 1 / 80, [([1, 2, 3, 4, 5, 6], 3, 1, False), ([4, 5, 6, 1, 2, 3], 3, 2, False)])
    tab = np.asfortranarray(np.random.default_rng(3).random(adv.sizeall))
    pv = R.getpoissonvar(adv)
    advd = R.AdvectionData(adv, tab, pv)
    assert np.array_equal(pv.v_square, R.dotprod([m.points for m in ms[3:]]) ** 2)
    dv = ms[3].step * ms[4].step * ms[5].step
    dsp = ms[0].step * ms[1].step * ms[2].step
    ke_ref = dsp * dv * float(np.sum(pv.v_square * np.sum(tab, axis=(0, 1, 2))))
    assert abs(R.compute_ke(advd) - ke_ref) <= 1e-13 * abs(ke_ref)
    advd.state_gen = 2
    pv.compute_charge(advd)      # what initcoef! does first at a velocity state that contains dim Nsp+1
    pv.compute_elfield()
    rho = dv * np.sum(tab, axis=(3, 4, 5))
    rho = rho - rho.sum() / rho.size
    assert np.max(np.abs(pv.rho - rho)) <= 1e-12 * np.max(np.abs(rho))
    ks = np.meshgrid(*[R.vec_k_fft(m) for m in ms[:3]], indexing="ij")
    k2 = ks[0] ** 2 + ks[1] ** 2 + ks[2] ** 2
    k2[0, 0, 0] = 1.0
    rk = np.fft.fftn(rho)
    ee = 0.0
    for d in range(3):
        mult = 1j * ks[d] / k2
        mult[0, 0, 0] = 0.0
        E = np.real(np.fft.ifftn(mult * rk))
        assert np.max(np.abs(pv.t_elfield[d] - E)) <= 1e-12 * max(np.max(np.abs(E)), 1e-300)
        ee += float(np.sum(E**2))
    assert abs(R.compute_ee(advd) - dsp * ee) <= 1e-12 * dsp * ee
