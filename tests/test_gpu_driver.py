"""GPU parity of the split advection driver, the charge density, the Poisson solve and the
energy diagnostics against the oracle model (oracle/refmodel.py), on the BASELINE configs'
shapes (reduced where the oracle would take minutes).

Tolerances (BASELINE.json north_star): <= 1e-12 relative max-abs per sweep / per field,
<= 1e-10 on Landau-damping electric-energy histories.
"""
import math

import numpy as np
import pytest

from helpers import relerr

pytestmark = pytest.mark.gpu


def _landau_1d1v(M, nx, nv, interp_f, eps=0.001, kx=0.5, dt=0.1):
    mx = M.UniformMesh(0.0, 2 * math.pi / kx, nx)
    mv = M.UniformMesh(-6.0, 6.0, nv)
    states = [([1, 2], 1, 1, True), ([2, 1], 1, 2, True)]
    adv = M.Advection((mx, mv), [interp_f(nx), interp_f(nv)], dt, states)
    f = M.dotprod((eps * np.cos(kx * mx.points) + 1, np.exp(-mv.points**2 / 2) / math.sqrt(2 * math.pi)))
    pv = M.getpoissonvar(adv)
    return adv, M.AdvectionData(adv, f, pv), pv


def _run(M, advd, nbdt, adv_fn):
    el = []
    for _ in range(nbdt):
        while adv_fn(advd):
            pass
        el.append(M.compute_ee(advd))
    return np.array(el)


def test_landau_1d1v_history_config1():
    """C1: Vlasov-Poisson 1D1V 128x256, Lagrange 9, Strang (examples/vlasov-poisson-1d1v.jl)."""
    import slb200 as S
    from oracle import refmodel as R

    _, advd_g, _ = _landau_1d1v(S, 128, 256, lambda n: S.Lagrange(9))
    _, advd_o, _ = _landau_1d1v(R, 128, 256, lambda n: R.Lagrange(9))
    nb = 100
    el_g = _run(S, advd_g, nb, S.advection)
    el_o = _run(R, advd_o, nb, R.advection)
    assert abs(advd_g.time_cur - advd_o.time_cur) == 0.0
    err = np.max(np.abs(el_g - el_o)) / np.max(np.abs(el_o))
    assert err <= 1e-10, err
    assert relerr(advd_g.getdata(), advd_o.data) <= 1e-11
    # physics: Landau damping rate for k = 0.5 is -0.1533; fit the envelope of log(ee)
    t = 0.1 * np.arange(1, nb + 1)
    peaks = [i for i in range(1, nb - 1) if el_g[i] > el_g[i - 1] and el_g[i] > el_g[i + 1]]
    slope = np.polyfit(t[peaks], np.log(el_g[peaks]), 1)[0]
    assert abs(slope / 2 + 0.1533) < 0.01, slope


def _vp_2d2v(M, sz, interps, dt=0.1, eps=0.5):
    m1 = M.UniformMesh(0.0, 4 * math.pi, sz[0])
    m2 = M.UniformMesh(0.0, 4 * math.pi, sz[1])
    v1 = M.UniformMesh(-6.0, 6.0, sz[2])
    v2 = M.UniformMesh(-6.0, 6.0, sz[3])
    tabst = [([3, 4, 1, 2], 1, 1, True), ([4, 3, 1, 2], 1, 1, True), ([1, 2, 4, 3], 1, 2, True), ([2, 1, 3, 4], 1, 2, True)]
    adv = M.Advection((m1, m2, v1, v2), interps, dt, tabst)
    fsp = lambda x: eps * np.cos(x / 2) + 1
    fv = lambda v: np.exp(-v**2 / 2) / math.sqrt(2 * math.pi)
    f = M.dotprod((fsp(m1.points), fsp(m2.points), fv(v1.points), fv(v2.points)))
    pv = M.getpoissonvar(adv)
    return adv, M.AdvectionData(adv, f, pv), pv


@pytest.mark.parametrize("sz,kind,order,nsteps", [((32, 32, 32, 32), "lagrange", 7, 3), ((16, 20, 24, 12), "lagrange", 9, 2),
                                                  ((32, 32, 32, 32), "bspline_fft", 11, 2), ((32, 16, 32, 16), "bspline_lu", 5, 2)])
def test_vlasov_poisson_2d2v_steps(sz, kind, order, nsteps):
    """C3/C4 shapes reduced (examples/vlasov-poisson-2d2v.jl): every stage of every step is
    compared, plus rho, E, ee, ke."""
    import slb200 as S
    from oracle import refmodel as R

    def mk(M):
        def one(n):
            return {"lagrange": lambda: M.Lagrange(order), "bspline_fft": lambda: M.BSplineFFT(order, n),
                    "bspline_lu": lambda: M.BSplineLU(order, n)}[kind]()
        return [one(n) for n in sz]

    adv_g, advd_g, pv_g = _vp_2d2v(S, sz, mk(S))
    adv_o, advd_o, pv_o = _vp_2d2v(R, sz, mk(R))
    assert adv_g.nbstates == adv_o.nbstates == 6
    for step in range(nsteps):
        more = True
        while more:
            assert advd_g.state_gen == advd_o.state_gen
            more = S.advection(advd_g)
            more_o = R.advection(advd_o)
            assert more == more_o
            assert relerr(advd_g.getdata(), advd_o.data) <= 1e-12 * (1 + 6 * step + advd_g.state_gen + 6)
        assert relerr(pv_g.rho, pv_o.rho) <= 1e-11
        for eg, eo in zip(pv_g.t_elfield, pv_o.t_elfield):
            assert relerr(eg, eo) <= 1e-11
        ee_g, ee_o = S.compute_ee(advd_g), R.compute_ee(advd_o)
        assert abs(ee_g - ee_o) <= 1e-11 * abs(ee_o)
        ke_g, ke_o = S.compute_ke(advd_g), R.compute_ke(advd_o)
        assert abs(ke_g - ke_o) <= 1e-12 * abs(ke_o)
    eg = S.getenergy(advd_g)
    eo = R.getenergy(advd_o)
    assert np.allclose(eg, eo, rtol=1e-11, atol=0)


@pytest.mark.parametrize("n,kind,order", [(64, "lagrange", 7), (128, "bspline_fft", 11)])
def test_vlasov_poisson_2d2v_full_size_step(n, kind, order):
    """C3 (2D2V 64^4, Lagrange 7) and C4 (2D2V 128^4, BSplineFFT 11) at their REAL sizes: one Strang step compared
    with the oracle after every stage (<= 1e-12 relative max-abs per sweep, accumulated), then rho, E and ee.
    The oracle runs on all host cores (OpenMP over lines = SimpleThreadsOpt); a 128^4 B-spline stage takes seconds."""
    import os

    import slb200 as S
    from oracle import refmodel as R

    ncores = len(os.sched_getaffinity(0))
    sz = (n,) * 4

    def build(M, **kw):
        m1 = M.UniformMesh(0.0, 4 * math.pi, n)
        m2 = M.UniformMesh(0.0, 4 * math.pi, n)
        v1 = M.UniformMesh(-6.0, 6.0, n)
        v2 = M.UniformMesh(-6.0, 6.0, n)
        mk = {"lagrange": lambda: M.Lagrange(order), "bspline_fft": lambda: M.BSplineFFT(order, n)}[kind]
        tabst = [([3, 4, 1, 2], 1, 1, True), ([4, 3, 1, 2], 1, 1, True), ([1, 2, 4, 3], 1, 2, True), ([2, 1, 3, 4], 1, 2, True)]
        adv = M.Advection((m1, m2, v1, v2), [mk() for _ in sz], 0.1, tabst, **kw)
        fsp = lambda x: 0.5 * np.cos(x / 2) + 1
        fv = lambda v: np.exp(-v**2 / 2) / math.sqrt(2 * math.pi)
        f = M.dotprod((fsp(m1.points), fsp(m2.points), fv(v1.points), fv(v2.points)))
        pv = M.getpoissonvar(adv)
        return M.AdvectionData(adv, f, pv), pv

    g, pv_g = build(S)
    o, pv_o = build(R, nthreads=ncores)
    buf = np.empty(sz, dtype=np.float64, order="F")
    more, stage = True, 0
    while more:
        more = S.advection(g)
        assert more == R.advection(o)
        stage += 1
        g.getdata(out=buf)
        np.subtract(buf, o.data, out=buf)
        err = float(np.max(np.abs(buf)) / np.max(np.abs(o.data)))
        assert err <= 1e-12 * stage, (stage, err)
    assert stage == 6
    assert relerr(pv_g.rho, pv_o.rho) <= 1e-11
    for eg, eo in zip(pv_g.t_elfield, pv_o.t_elfield):
        assert relerr(eg, eo) <= 1e-11
    ee_g, ee_o = S.compute_ee(g), R.compute_ee(o)
    assert abs(ee_g - ee_o) <= 1e-11 * abs(ee_o)
    g.close()


def test_field_solve_pieces_random():
    """compute_charge! / compute_elfield! on random data (test/test_poisson.jl:38-108 shape)."""
    import slb200 as S
    from oracle import refmodel as R

    rng = np.random.default_rng(20240611)
    sz = (16, 24, 10, 14)
    meshes = lambda M: (M.UniformMesh(-1.0, 3.0, sz[0]), M.UniformMesh(0.0, 5.0, sz[1]), M.UniformMesh(-3.0, 1.0, sz[2]), M.UniformMesh(-9.0, 7.0, sz[3]))
    tabst = [([3, 4, 1, 2], 1, 1, True), ([4, 3, 1, 2], 1, 1, True), ([1, 2, 4, 3], 1, 2, True), ([2, 1, 3, 4], 1, 2, True)]
    f = np.asfortranarray(rng.random(sz))
    out = {}
    for name, M in (("g", S), ("o", R)):
        adv = M.Advection(meshes(M), [M.Lagrange(3)] * 4, 0.1, tabst)
        pv = M.getpoissonvar(adv)
        advd = M.AdvectionData(adv, f, pv)
        pv.compute_charge(advd)
        pv.compute_elfield()
        out[name] = (np.array(pv.rho), [np.array(e) for e in pv.t_elfield], M.compute_ee(advd), M.compute_ke(advd))
    assert relerr(out["g"][0], out["o"][0]) <= 1e-12
    assert abs(np.sum(out["g"][0])) < 1e-10  # zero mean
    for eg, eo in zip(out["g"][1], out["o"][1]):
        assert relerr(eg, eo) <= 1e-12
    assert abs(out["g"][2] - out["o"][2]) <= 1e-12 * abs(out["o"][2])
    assert abs(out["g"][3] - out["o"][3]) <= 1e-12 * abs(out["o"][3])


def _rotation(M, sz, interps, nbdt):
    """test/test_rotation.jl:41-89"""
    mx = M.UniformMesh(-5.0, 5.0, sz[0])
    my = M.UniformMesh(-6.0, 4.5, sz[1])
    dt = 2 * math.pi / nbdt
    states = [([1, 2], 1, 1, True), ([2, 1], 1, 2, True)]
    adv = M.Advection((mx, my), interps, dt, states, tab_coef=M.magicsplit(dt))
    X, Y = np.meshgrid(mx.points, my.points, indexing="ij")
    f = np.asfortranarray(np.exp(-2 * (X**2 + (Y + 1.2) ** 2)))
    return adv, M.AdvectionData(adv, f, M.getrotationvar(adv)), f


@pytest.mark.parametrize("kind,order,sz", [("lagrange", 5, (400, 300)), ("hermite", 5, (400, 300)), ("bspline_lu", 5, (128, 256)),
                                           ("bspline_fft", 5, (128, 256)), ("bspline_lu", 5, (1024, 1024))])
def test_rotation_config2(kind, order, sz):
    """C2 (rotation, B-spline LU 5, 1024x1024) and the reference's rotation tests
    (test/test_rotation.jl:237-252): full turn in 11 steps with magicsplit returns the
    initial Gaussian (error < 1e-3) and matches the oracle stage by stage."""
    import slb200 as S
    from oracle import refmodel as R

    def mk(M):
        return [{"lagrange": lambda n: M.Lagrange(order), "hermite": lambda n: M.Hermite(order),
                 "bspline_lu": lambda n: M.BSplineLU(order, n), "bspline_fft": lambda n: M.BSplineFFT(order, n)}[kind](n) for n in sz]

    nbdt = 11
    _, advd_g, f0 = _rotation(S, sz, mk(S), nbdt)
    _, advd_o, _ = _rotation(R, sz, mk(R), nbdt)
    for _ in range(nbdt):
        while S.advection(advd_g):
            pass
        while R.advection(advd_o):
            pass
    out = advd_g.getdata()
    assert relerr(out, advd_o.data) <= 1e-11
    assert np.max(np.abs(out - f0)) < 1e-3


@pytest.mark.parametrize("kind,order,sz,tol", [("lagrange", 5, (200, 300), 1e-3), ("bspline_lu", 5, (128, 64), 1e-6)])
def test_translation(kind, order, sz, tol):
    """test/test_translation.jl:42-101,139-190: periodic function translated with Strang splitting."""
    import slb200 as S
    from oracle import refmodel as R

    def build(M):
        m1 = M.UniformMesh(0.0, 1.0, sz[0])
        m2 = M.UniformMesh(0.0, 1.0, sz[1])
        dt, v = 0.01, (30.0, -20.0)  # shifts in grid units per unit time
        interps = [M.Lagrange(order) if kind == "lagrange" else M.BSplineLU(order, n) for n in sz]
        adv = M.Advection((m1, m2), interps, dt, [([1, 2], 1, 1, True), ([2, 1], 1, 2, True)], tab_coef=M.strangsplit(dt))
        X, Y = np.meshgrid(m1.points, m2.points, indexing="ij")
        f = np.asfortranarray(np.exp(-(np.sin(2 * np.pi * X) + np.sin(2 * np.pi * Y))))
        return adv, M.AdvectionData(adv, f, M.gettranslationvar(v)), (m1, m2, dt, v)

    _, advd_g, (m1, m2, dt, v) = build(S)
    _, advd_o, _ = build(R)
    nb = 20
    for _ in range(nb):
        while S.advection(advd_g):
            pass
        while R.advection(advd_o):
            pass
    out = advd_g.getdata()
    assert relerr(out, advd_o.data) <= 1e-11
    X, Y = np.meshgrid(m1.points, m2.points, indexing="ij")
    sx, sy = nb * dt * v[0] * m1.step, nb * dt * v[1] * m2.step
    exact = np.exp(-(np.sin(2 * np.pi * (X + sx)) + np.sin(2 * np.pi * (Y + sy))))
    assert np.max(np.abs(out - exact)) < tol


def test_full_size_properties_128_4():
    """Size-independent properties at BASELINE's full size (2D2V 128^4, Lagrange 7), where the
    oracle is too slow for a point-wise check of every sweep: (i) integer shifts are exact
    circular shifts, (ii) a sampled slab of one fractional sweep per dim matches the oracle,
    (iii) the charge density of a product-form f is known in closed form."""
    import slb200 as S
    from oracle import refmodel as R
    from helpers import DeviceGrid, oracle_sweep

    n = 128
    rng = np.random.default_rng(20240611)
    a = [rng.random(n) + 0.5 for _ in range(4)]
    f = S.dotprod(a)  # 2.1 GB product-form array: any slab is cheap to rebuild on the host
    g = DeviceGrid(f)
    it, oit = S.Lagrange(7), R.Lagrange(7)
    sl = (slice(5, 9), slice(17, 19), slice(60, 63), slice(100, 102))
    for dim in range(4):
        # (i) integer shift by 3 along dim == roll
        g.sweep(dim, it, np.array([3.0]), [0, 0, 0, 0])
        out = g.get()
        assert np.array_equal(out, np.roll(f, -3, axis=dim)), dim
        g.sweep(dim, it, np.array([-3.0]), [0, 0, 0, 0])
        # (ii) fractional, index-dependent shift on a slab containing whole lines along dim
        other = (dim + 1) % 4
        tab = rng.uniform(-6, 6, n)
        astride = [0] * 4
        astride[other] = 1
        g.sweep(dim, it, tab, astride)
        out = g.get()
        idx = list(sl)
        idx[dim] = slice(None)
        sub = np.asfortranarray(f[tuple(idx)])
        subtab = tab[idx[other]] if idx[other] != slice(None) else tab
        ref = oracle_sweep(sub, dim, oit, np.ascontiguousarray(subtab), astride)
        assert relerr(out[tuple(idx)], ref) <= 1e-12, dim
        # undo: restore f for the next dim
        g.close()
        g = DeviceGrid(f)
    g.close()


def test_step_graph_replays_whole_steps_bitwise():
    """CUDA-graph replay of whole Strang steps (C1 shape, 1D1V 128 x 256 Lagrange 9): the same kernels in the same
    order, so the data and the electric-energy history equal the step-by-step driver bit for bit."""
    import slb200 as S

    _, a, _ = _landau_1d1v(S, 128, 256, lambda n: S.Lagrange(9))
    _, b, _ = _landau_1d1v(S, 128, 256, lambda n: S.Lagrange(9))
    el_a = _run(S, a, 6, S.advection)
    g = S.StepGraph(b, nsteps=2)
    el_b = []
    for _ in range(3):
        g.launch()
        el_b += g.energies()
    assert b.time_cur == a.time_cur
    assert np.array_equal(np.array(el_b), el_a)
    assert np.array_equal(a.getdata(), b.getdata())
    # the driver can continue step by step after graph replays
    assert S.advection(b) is True
    with pytest.raises(ValueError):   # mid-step
        S.StepGraph(b, nsteps=2)
    while S.advection(b):
        pass
    with pytest.raises(ValueError):   # three passes per step: an odd number of steps leaves the buffers swapped
        S.StepGraph(b, nsteps=1)
    g.close()


@pytest.mark.parametrize("order,nsteps,repeat", [(9, 2, 3), (5, 2, 1), (3, 4, 2)])
def test_step_program_runs_whole_steps_bitwise(order, nsteps, repeat):
    """Step program (one persistent cooperative kernel interpreting the recorded stages, C1 shape 1D1V 128 x 256): every
    op keeps the arithmetic of the kernel it replaces, so the data and the electric-energy history equal the
    step-by-step driver bit for bit."""
    import slb200 as S

    _, a, _ = _landau_1d1v(S, 128, 256, lambda n: S.Lagrange(order))
    _, b, _ = _landau_1d1v(S, 128, 256, lambda n: S.Lagrange(order))
    front0 = S._lib.lib().slb_grid_front(b.grid)
    g = S.StepProgram(b, nsteps=nsteps, repeat=repeat)
    assert S._lib.lib().slb_grid_front(b.grid) == front0 and b.state_gen == 1   # nothing ran, roles restored
    assert g.nops == 6 * nsteps and g.nbarriers == 4 * nsteps   # sweep x | charge | field + sweep v | sweep x (+ ee)
    el_a = _run(S, a, 2 * nsteps * repeat, S.advection)
    el_b = []
    for _ in range(2):
        g.launch()
        el_b += g.energies()
    assert b.time_cur == a.time_cur
    assert np.array_equal(np.array(el_b), el_a)
    assert np.array_equal(a.getdata(), b.getdata())
    # the driver continues step by step after a program, and a second program can be built
    _run(S, a, 1, S.advection)
    _run(S, b, 1, S.advection)
    assert np.array_equal(a.getdata(), b.getdata())
    g.close()


def test_step_program_other_shapes_and_refusals():
    """non-square small grids (64 x 32, Hermite-free Lagrange 7); refusals leave the data untouched: odd step counts,
    B-spline stages, 2D2V (pair fusion / two space dims)"""
    import slb200 as S

    _, a, _ = _landau_1d1v(S, 64, 32, lambda n: S.Lagrange(7))
    _, b, _ = _landau_1d1v(S, 64, 32, lambda n: S.Lagrange(7))
    g = S.StepProgram(b, nsteps=2, repeat=2)
    el_a = _run(S, a, 4, S.advection)
    g.launch()
    assert np.array_equal(np.array(g.energies()), el_a)
    assert np.array_equal(a.getdata(), b.getdata())
    g.close()
    before = b.getdata()
    with pytest.raises(ValueError):   # three sweeps per step: an odd number of steps leaves the buffers swapped
        S.StepProgram(b, nsteps=1)
    _, c, _ = _landau_1d1v(S, 64, 32, lambda n: S.BSplineLU(5, n))
    with pytest.raises(S.SlbError):
        S.StepProgram(c, nsteps=2)
    assert np.array_equal(b.getdata(), before)
    # after a refusal the context records nothing: stepwise calls and graphs work
    _run(S, c, 1, S.advection)
    gg = S.StepGraph(b, nsteps=2)
    gg.launch()
    gg.close()


def test_step_program_with_line_sums_vxv_order():
    """v - x - v Strang order (bench.py's C1): the velocity sweep that ends a step leaves line sums for the next step's
    charge density (slb_grid_set_linesum); the program records that flow too and stays bit-identical"""
    import bench
    import slb200 as S

    a, _ = bench._cfg_c1(S)
    b, _ = bench._cfg_c1(S)
    for g in (a, b):   # one real step first: the line sums of its last sweep feed the first recorded field solve
        while S.advection(g):
            pass
    p = S.StepProgram(b, nsteps=2, repeat=2)
    assert p.nops == 2 * 7   # per step: field (from line sums), v sweep, x sweep, charge, field, v sweep, ee
    el_a = _run(S, a, 4, S.advection)
    p.launch()
    assert np.array_equal(np.array(p.energies()), el_a)
    assert np.array_equal(a.getdata(), b.getdata())
    el_a2 = _run(S, a, 1, S.advection)   # continuing stepwise uses the line sums the program left behind
    el_b2 = _run(S, b, 1, S.advection)
    assert np.array_equal(el_a2, el_b2) and np.array_equal(a.getdata(), b.getdata())
    p.close()
