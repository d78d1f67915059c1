"""GPU parity of the pair-fused sweeps (slb_sweep_pair, K7): two advection! stages in one pass.

The fused pass must equal two separate slb_sweep calls BIT FOR BIT (same arithmetic, the
intermediate is rounded to Float64 either way) and the oracle within 1e-12 per sweep pair.
"""
import numpy as np
import pytest

from helpers import DeviceGrid, make_pair, oracle_sweep, relerr

pytestmark = pytest.mark.gpu
SEED = 20240611


def _alpha(rng, shape, dim, mode, skip=None):
    """alpha table over the dims other than `dim` (and `skip`: the first sweep of a fused pair must
    not depend on the second sweep's dim)"""
    nd = len(shape)
    astr = [0] * nd
    other = [d for d in range(nd) if d != dim and d != skip]
    if not other:
        mode = "const"
    if mode == "const":
        return np.array([rng.uniform(-8, 8)]), astr
    if mode == "space":  # Vlasov velocity sweep: alpha = alpha(x) over the leading dims not swept
        stride = 1
        for d in other[: max(1, len(other) - 1)]:
            astr[d] = stride
            stride *= shape[d]
        return rng.uniform(-3, 3, stride), astr
    stride = 1
    for d in other:
        astr[d] = stride
        stride *= shape[d]
    return rng.uniform(-8, 8, stride), astr


@pytest.mark.parametrize("shape,dimA,dimB", [
    ((16, 12, 20, 24), 2, 3), ((16, 12, 20, 24), 3, 2), ((32, 6, 18, 10), 1, 2), ((8, 10, 12, 14), 1, 3),
    ((24, 5, 16), 1, 2), ((24, 16, 12), 2, 1), ((130, 3, 14, 11), 2, 3), ((48, 3, 70, 5), 2, 3),
    ((16, 12, 20, 24), 0, 1), ((40, 33, 6), 0, 1), ((128, 17, 3, 2), 0, 2), ((20, 31), 0, 1), ((600, 9, 2), 0, 1), ((1100, 12), 0, 1),
])
@pytest.mark.parametrize("kind,order", [("lagrange", 3), ("lagrange", 7), ("lagrange", 11), ("hermite", 5)])
@pytest.mark.parametrize("mode", ["full", "space", "const"])
def test_pair_equals_two_sweeps_bitwise(shape, dimA, dimB, kind, order, mode):
    import slb200 as S

    rng = np.random.default_rng(SEED + order)
    f = np.asfortranarray(rng.random(shape))
    iA, oA = make_pair(kind, order, shape[dimA])
    iB, oB = make_pair(kind, order, shape[dimB])
    tA, sA = _alpha(rng, shape, dimA, mode, skip=dimB)
    tB, sB = _alpha(rng, shape, dimB, mode, skip=dimA)
    for flags in (0, 1):
        g1 = DeviceGrid(f)
        g1.sweep(dimA, iA, tA, sA, flags=flags)
        g1.sweep(dimB, iB, tB, sB, flags=flags)
        two = g1.get()
        g1.close()
        g2 = DeviceGrid(f)
        try:
            g2.sweep_pair(dimA, iA, tA, sA, dimB, iB, tB, sB, flags=flags)
        except S.SlbError as exc:  # documented: callers then issue the two sweeps separately
            assert "-4" in str(exc) and shape[dimA] < order + 1, exc
            g2.close()
            continue
        fused = g2.get()
        g2.close()
        assert np.array_equal(fused, two), (shape, dimA, dimB, kind, order, mode, flags, relerr(fused, two))
    ref = oracle_sweep(oracle_sweep(f, dimA, oA, tA, sA), dimB, oB, tB, sB)
    assert relerr(two, ref) <= 2e-12


def test_pair_rejects_unsupported_combinations():
    import slb200 as S

    rng = np.random.default_rng(SEED)
    f = np.asfortranarray(rng.random((16, 8, 8, 8)))
    g = DeviceGrid(f)
    L7, L5 = S.Lagrange(7), S.Lagrange(5)
    one, z = np.array([0.3]), [0, 0, 0, 0]
    with pytest.raises(S.SlbError):   # the second sweep cannot run along dim 0
        g.sweep_pair(1, L7, one, z, 0, L7, one, z)
    with pytest.raises(S.SlbError):   # alpha_A must not depend on the second sweep's dim
        g.sweep_pair(2, L7, np.full(8, 0.3), [0, 0, 0, 1], 3, L7, one, z)
    with pytest.raises(S.SlbError):   # nor alpha_B on the first sweep's dim
        g.sweep_pair(2, L7, one, z, 3, L7, np.full(8, 0.3), [0, 0, 1, 0])
    with pytest.raises(S.SlbError):   # even orders are not instantiated
        L4 = S.Lagrange(4)
        g.sweep_pair(2, L4, one, z, 3, L4, one, z)
    with pytest.raises(S.SlbError):   # different orders
        g.sweep_pair(2, L7, one, z, 3, L5, one, z)
    with pytest.raises(ValueError):   # same dim twice
        g.sweep_pair(2, L7, one, z, 2, L7, one, z)
    with pytest.raises(S.SlbError):   # B-spline pre-solve
        b = S.BSplineLU(5, 8)
        g.sweep_pair(2, b, one, z, 3, b, one, z)
    g.close()


def test_pair_repeated_launches_and_linesums():
    """ring and counters are reused across launches; phase-B line sums feed the charge density"""
    import ctypes as C

    import slb200 as S
    from slb200 import _lib

    rng = np.random.default_rng(SEED)
    shape = (16, 8, 12, 10)
    f = np.asfortranarray(rng.random(shape))
    L = S.Lagrange(5)
    tA, sA = rng.uniform(-2, 2, 16 * 8), [1, 16, 0, 0]
    tB, sB = rng.uniform(-2, 2, 16 * 8), [1, 16, 0, 0]
    g1, g2 = DeviceGrid(f), DeviceGrid(f)
    ctx = g2.ctx
    nls = 16 * 8 * 12
    ls = ctx.malloc(nls * 8)
    for it in range(5):
        g1.sweep(2, L, tA, sA)
        g1.sweep(3, L, tB, sB)
        if it == 4:
            _lib.check(_lib.lib().slb_grid_set_linesum(g2.h, ls))
        g2.sweep_pair(2, L, tA, sA, 3, L, tB, sB)
    a, b = g1.get(), g2.get()
    assert np.array_equal(a, b)
    sums = ctx.to_host(ls, nls).reshape((16, 8, 12), order="F")
    assert relerr(sums, a.sum(axis=3)) <= 1e-14
    ctx.free(ls)
    g1.close()
    g2.close()


# ---------------------------------------------------------------------------------------------
# driver level: advection() holds a stage back and fuses it with the next one
# ---------------------------------------------------------------------------------------------
def _vp_2d2v(M, sz, order, dt=0.1, eps=0.5):
    import math

    ms = (M.UniformMesh(0.0, 4 * math.pi, sz[0]), M.UniformMesh(0.0, 4 * math.pi, sz[1]),
          M.UniformMesh(-6.0, 6.0, sz[2]), M.UniformMesh(-6.0, 6.0, sz[3]))
    tabst = [([3, 4, 1, 2], 1, 1, True), ([4, 3, 1, 2], 1, 1, True), ([1, 2, 4, 3], 1, 2, True), ([2, 1, 3, 4], 1, 2, True)]
    adv = M.Advection(ms, [M.Lagrange(order)] * 4, dt, tabst)
    fsp = lambda x: eps * np.cos(x / 2) + 1
    fv = lambda v: np.exp(-v**2 / 2) / math.sqrt(2 * math.pi)
    f = M.dotprod((fsp(ms[0].points), fsp(ms[1].points), fv(ms[2].points), fv(ms[3].points)))
    return adv, M.AdvectionData(adv, f, M.getpoissonvar(adv))


@pytest.mark.parametrize("use_linesum", [False, True])
@pytest.mark.parametrize("sz,order", [((32, 32, 32, 32), 7), ((48, 20, 24, 12), 9), ((16, 16, 40, 36), 5)])
def test_driver_fuses_pairs_and_matches_unfused_and_oracle(sz, order, use_linesum):
    """A Strang step of examples/vlasov-poisson-2d2v.jl runs as three fused passes (v1v2, x1x2,
    v1v2); the result equals the stage-by-stage execution bit for bit and the oracle to 1e-12/sweep.
    (With the line-sum shortcut for rho the two drivers add the same numbers in a different
    order -- the fused kernel emits a line's outputs starting at the wrap point -- so they agree
    to rounding only.)"""
    import slb200 as S
    from oracle import refmodel as R

    _, fused = _vp_2d2v(S, sz, order)
    _, plain = _vp_2d2v(S, sz, order)
    _, orc = _vp_2d2v(R, sz, order)
    plain.fuse_pairs = False
    fused.use_linesum = plain.use_linesum = use_linesum
    assert fused.fuse_pairs
    nsteps = 3
    for step in range(nsteps):
        while S.advection(fused):
            pass
        while S.advection(plain):
            pass
        while R.advection(orc):
            pass
        assert fused.n_fused == 3 * (step + 1) and plain.n_fused == 0
        ee_f, ee_p, ee_o = S.compute_ee(fused), S.compute_ee(plain), R.compute_ee(orc)
        assert ee_f == ee_p if not use_linesum else abs(ee_f - ee_p) <= 1e-13 * abs(ee_p)
        assert abs(ee_f - ee_o) <= 1e-11 * abs(ee_o)
    a, b = fused.getdata(), plain.getdata()
    assert np.array_equal(a, b) if not use_linesum else relerr(a, b) <= 1e-13
    assert relerr(a, orc.data) <= 1e-12 * 6 * nsteps
    assert abs(S.compute_ke(fused) - R.compute_ke(orc)) <= 1e-12 * abs(R.compute_ke(orc))


def test_driver_flushes_a_held_back_stage_on_access():
    """getdata() between the two stages of a pair must see the first stage applied (the reference
    mutates advd.data in every advection! call, src/advection.jl:594-657)."""
    import slb200 as S
    from oracle import refmodel as R

    sz = (16, 16, 16, 16)
    _, g = _vp_2d2v(S, sz, 7)
    _, o = _vp_2d2v(R, sz, 7)
    more = True
    while more:
        more = S.advection(g)
        R.advection(o)
        assert relerr(g.getdata(), o.data) <= 1e-12 * 8
    assert g.n_fused == 0  # every stage was flushed by the read that followed it
    while S.advection(g):
        pass
    while R.advection(o):
        pass
    assert g.n_fused == 3
    assert relerr(g.getdata(), o.data) <= 1e-12 * 14


def test_driver_fuses_translation_with_host_tables():
    """2-D translation (src/translation.jl): both stages have constant host-side shifts; the x/y pair
    of a Strang step [x y x] fuses as (x, y) and the trailing x runs alone."""
    import slb200 as S

    def build(fuse):
        m1, m2 = S.UniformMesh(0.0, 1.0, 96), S.UniformMesh(0.0, 1.0, 80)
        adv = S.Advection((m1, m2), [S.Lagrange(5), S.Lagrange(5)], 0.01, [([1, 2], 1, 1, True), ([2, 1], 1, 2, True)])
        X, Y = np.meshgrid(m1.points, m2.points, indexing="ij")
        f = np.asfortranarray(np.exp(-(np.sin(2 * np.pi * X) + np.sin(2 * np.pi * Y))))
        d = S.AdvectionData(adv, f, S.gettranslationvar((30.0, -20.0)))
        d.fuse_pairs = fuse
        return d

    a, b = build(True), build(False)
    for _ in range(4):
        while S.advection(a):
            pass
        while S.advection(b):
            pass
    assert a.n_fused == 4 and b.n_fused == 0
    assert np.array_equal(a.getdata(), b.getdata())


# ---------------------------------------------------------------------------------------------
# states with ndims = 2 and constant shifts: the reference's N-D tensor stencil == one fused pass
# ---------------------------------------------------------------------------------------------
def _vp_2d_states(M, sz, kind, order, states):
    import math

    ms = (M.UniformMesh(0.0, 4 * math.pi, sz[0]), M.UniformMesh(0.0, 4 * math.pi, sz[1]),
          M.UniformMesh(-6.0, 6.0, sz[2]), M.UniformMesh(-6.0, 6.0, sz[3]))
    mk = {"lagrange": lambda n: M.Lagrange(order), "bspline_lu": lambda n: M.BSplineLU(order, n)}[kind]
    adv = M.Advection(ms, [mk(n) for n in sz], 0.1, states)
    fsp = lambda x: 0.5 * np.cos(x / 2) + 1
    fv = lambda v: np.exp(-v**2 / 2) / math.sqrt(2 * math.pi)
    f = M.dotprod((fsp(ms[0].points), fsp(ms[1].points), fv(ms[2].points), fv(ms[3].points)))
    return M.AdvectionData(adv, f, M.getpoissonvar(adv))


@pytest.mark.parametrize("kind,order,sz", [("lagrange", 7, (16, 18, 20, 16)), ("lagrange", 5, (12, 12, 14, 16)),
                                           ("bspline_lu", 5, (16, 12, 16, 12))])
def test_states_with_ndims_2_match_the_nd_tensor_stencil_oracle(kind, order, sz):
    """test/test_poisson2d.jl:276 / examples/vlasov-poisson-2d2v.jl:196: tabst =
    [([1,2,3,4], 2, 1, true), ([3,4,1,2], 2, 2, true)].  The oracle evaluates the N-D tensor
    stencil of src/interpolation.jl:212-231 slice by slice; the product runs one fused pass (or two
    sweeps for B-splines) per state."""
    import slb200 as S
    from oracle import refmodel as R

    states = [([3, 4, 1, 2], 2, 1, True), ([1, 2, 3, 4], 2, 2, True)]
    g = _vp_2d_states(S, sz, kind, order, states)
    o = _vp_2d_states(R, sz, kind, order, states)
    assert g.adv.nbstates == o.adv.nbstates == 3
    nst = 0
    for step in range(2):
        more = True
        while more:
            more = S.advection(g)
            assert more == R.advection(o)
            nst += 1
            assert relerr(g.getdata(), o.data) <= 2e-12 * nst
        assert abs(S.compute_ee(g) - R.compute_ee(o)) <= 1e-11 * abs(R.compute_ee(o))
    if kind == "lagrange":
        assert g.n_fused == 6      # every 2-D state ran as one pass over HBM
    eg, eo = S.getenergy(g), R.getenergy(o)
    assert np.allclose(eg, eo, rtol=1e-11, atol=0)


def test_states_with_ndims_2_space_first_and_translation():
    """the other state order of the reference test (x pair first), and a 2-D translation state"""
    import slb200 as S
    from oracle import refmodel as R

    states = [([1, 2, 3, 4], 2, 1, True), ([3, 4, 1, 2], 2, 2, True)]
    sz = (16, 16, 12, 20)
    # the v-state comes second: the first x half-step runs before any field solve
    g = _vp_2d_states(S, sz, "lagrange", 7, states)
    o = _vp_2d_states(R, sz, "lagrange", 7, states)
    for _ in range(2):
        while S.advection(g):
            pass
        while R.advection(o):
            pass
    assert relerr(g.getdata(), o.data) <= 2e-12 * 6
    assert g.n_fused == 6


def test_fused_pair_full_size_properties_128_4():
    """BASELINE's full size (2D2V 128^4, Lagrange 7), where the oracle is too slow for a point-wise
    check: (i) integer shifts in both sweeps of a fused pass are an exact 2-D circular shift,
    (ii) the fused pass equals two sweeps bit for bit, (iii) a slab that contains whole (v1, v2)
    planes matches the oracle applied to that slab."""
    import slb200 as S
    from oracle import refmodel as R

    n = 128
    rng = np.random.default_rng(20240611)
    a = [rng.random(n) + 0.5 for _ in range(4)]
    f = S.dotprod(a)
    it, oit = S.Lagrange(7), R.Lagrange(7)
    z = [0, 0, 0, 0]
    for dA, dB in ((2, 3), (0, 1)):
        g = DeviceGrid(f)
        g.sweep_pair(dA, it, np.array([3.0]), z, dB, it, np.array([-5.0]), z)
        out = g.get()
        assert np.array_equal(out, np.roll(np.roll(f, -3, axis=dA), 5, axis=dB)), (dA, dB)
        g.close()
    # fractional shifts that depend on the other dims like the Vlasov providers' do
    g1, g2 = DeviceGrid(f), DeviceGrid(f)
    tE1, tE2 = rng.uniform(-0.6, 0.6, n * n), rng.uniform(-0.6, 0.6, n * n)
    sE = [1, n, 0, 0]
    g1.sweep(2, it, tE1, sE)
    g1.sweep(3, it, tE2, sE)
    g2.sweep_pair(2, it, tE1, sE, 3, it, tE2, sE)
    two, fused = g1.get(), g2.get()
    g1.close()
    g2.close()
    assert np.array_equal(two, fused)
    sl = (slice(40, 44), slice(7, 9), slice(None), slice(None))   # whole (v1, v2) planes of a few x
    sub = np.asfortranarray(f[sl])
    t1 = tE1.reshape((n, n), order="F")[sl[0], sl[1]].reshape(-1, order="F")
    t2 = tE2.reshape((n, n), order="F")[sl[0], sl[1]].reshape(-1, order="F")
    ref = oracle_sweep(oracle_sweep(sub, 2, oit, t1, [1, 4, 0, 0]), 3, oit, t2, [1, 4, 0, 0])
    assert relerr(fused[sl], ref) <= 2e-12


def test_const_shift_state_with_ndims_3_and_4():
    """const-shift states with ndims > 2 (src/interpolation.jl:212-231 for N = 3, 4): a fused pair + single sweeps on
    the device against the oracle's N-D tensor stencil"""
    import slb200 as S
    from oracle import refmodel as R

    def build(M, sz, its, perm, vals):
        ms = tuple(M.UniformMesh(0.0, 1.0, n) for n in sz)
        adv = M.Advection(ms, [M.Lagrange(o) for o in its], 0.01, [(perm, len(perm), 1, True)], tab_coef=[0.01])
        rng = np.random.default_rng(9)
        return M.AdvectionData(adv, np.asfortranarray(rng.random(sz)), M.gettranslationvar(vals))

    for sz, its, perm, vals in (((16, 12, 10), (5, 5, 3), [3, 1, 2], (130.0, -270.0, 55.0)),
                                ((12, 10, 8, 8), (3, 3, 3, 3), [2, 1, 4, 3], (130.0, -270.0, 55.0, -20.0))):
        g, o = build(S, sz, its, perm, vals), build(R, sz, its, perm, vals)
        for _ in range(2):
            assert S.advection(g) == R.advection(o)
        assert relerr(g.getdata(), o.data) <= 1e-12
        assert g.n_fused >= 2   # at least one fused pair per call


def test_partial_charge_planes_from_the_space_pass():
    """SLB_FUSED_RHO (off by default: measured slower): the x1x2 pass leaves partial planes of the charge
    density and the field solve reduces those instead of f -- same numbers added in another order"""
    import slb200 as S

    sz = (32, 16, 16, 16)
    _, a = _vp_2d2v(S, sz, 7)
    _, b = _vp_2d2v(S, sz, 7)
    a.use_rhopart = True
    planes = []
    for step in range(2):
        more = True
        while more:
            more = S.advection(a)
            planes.append(a._rhopart_planes)
        while S.advection(b):
            pass
        ee_a, ee_b = S.compute_ee(a), S.compute_ee(b)
        assert abs(ee_a - ee_b) <= 1e-13 * abs(ee_b)
    assert max(planes) == 16 * 16 // 4     # the x1x2 pass ran with four passive points per block
    assert relerr(a.getdata(), b.getdata()) <= 1e-13
