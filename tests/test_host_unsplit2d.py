"""Host sequencing logic of the unsplit 2-D layer (slb200/unsplit2d.py, the StdPoisson2d / rotation
providers, interpolate_nd) on a machine WITHOUT a GPU: libslb200's entry points are replaced by the
test double of tests/fake_device.py (host memory + the CUDA per-point body compiled for the host), the
results are compared with the oracle.  The same comparisons run against the real library in
tests/test_gpu_unsplit2d.py."""
import numpy as np
import pytest

import fake_device
from helpers import relerr
from oracle import refmodel as R, unsplit2d as U
import test_gpu_unsplit2d as G
from test_oracle_unsplit2d import poisson2d_run


@pytest.fixture
def fake(monkeypatch):
    return fake_device.install(monkeypatch)


def test_interpolate_nd_seam(fake):
    import slb200 as S

    rng = np.random.default_rng(1)
    for spec, n1, n2, ncomp in G.SPECS[:1] + G.SPECS[3:4] + G.SPECS[7:8] + G.SPECS[10:]:
        its, rits = G._pairs(spec, n1, n2)
        f = np.asfortranarray(rng.random((n1, n2, ncomp)))
        fin = f if ncomp > 1 else np.asfortranarray(f[:, :, 0])
        dec = np.asfortranarray(rng.uniform(-8, 8, (n1, n2, 2)))
        out = np.empty_like(fin)
        S.interpolate_nd(out, fin, dec, its)
        assert relerr(out, U.interpolate_points(fin, dec, rits)) <= 1e-12
    assert "interp2d_points" in fake.calls


@pytest.mark.parametrize("spec,alg,ordalg", [((("lagrange", 5), ("lagrange", 5)), "NoTimeAlg", 0),
                                            ((("bspline_lu", 5), ("bspline_lu", 5)), "NoTimeAlg", 0),
                                            ((("lagrange", 5), ("lagrange", 5)), "ABTimeAlg_init", 3)])
def test_swirling_driver_sequencing(fake, spec, alg, ordalg):
    worst, g, _ = G._swirling_both(spec, 50, 6, alg, ordalg, sz=(40, 40))
    assert worst <= 1e-11
    if alg != "NoTimeAlg":
        assert len(g.t_bufc) == ordalg - 1


@pytest.mark.parametrize("alg,ordalg", [("ABTimeAlg_ip", 2), ("ABTimeAlg_ip", 4), ("ABTimeAlg_new", 2), ("ABTimeAlg_new", 3), ("NoTimeAlg", 0)])
def test_poisson2d_driver_sequencing(fake, alg, ordalg):
    import slb200 as S

    sz = (32, 40)
    its, rits = G._pairs((("lagrange", 5), ("lagrange", 5)), *sz)
    dg, g = poisson2d_run(S, lambda adv: S.getpoissonvar(adv, type=S.StdPoisson2d), sz, its, 0.1, 4, getattr(S, alg), ordalg)
    do, o = poisson2d_run(R, U.getpoissonvar2d, sz, rits, 0.1, 4, getattr(R, alg), ordalg)
    assert relerr(g.getdata(), o.data) <= 1e-11
    assert abs(dg - do) <= 1e-11 * abs(R.getenergy(o)[2])
    if alg != "NoTimeAlg":
        assert len(g.t_bufc) == len(o.t_bufc) >= ordalg - 1
        hist_g = [f.to_host() for f in g.t_bufc]
        for a, b in zip(hist_g, o.t_bufc):
            assert relerr(a, b) <= 1e-10
    assert relerr(g.bufcur.to_host(), o.bufcur) <= 1e-10


def test_rotation2d_driver_sequencing(fake):
    G.test_rotation2d_abtimealg_matches_oracle()


def test_pool_recycles_storage(fake):
    import slb200 as S
    from slb200 import _lib

    ctx = _lib.default_context()
    a = S.DeviceField(ctx, 8, 8, 2)
    p = a.ptr.value
    a.free()
    b = S.DeviceField(ctx, 8, 8, 2)
    assert b.ptr.value == p and a.ptr is None
    v = S.DeviceField.view(ctx, 8, 8, 2, b.ptr)
    v.free()  # a view never returns storage it does not own
    assert not S.DeviceField._pool.get((id(ctx), 128))


@pytest.mark.parametrize("alg,ordalg", [("NoTimeAlg", 0), ("ABTimeAlg_ip", 2), ("ABTimeAlg_ip", 3)])
def test_quasigeostrophic_driver_sequencing(fake, alg, ordalg):
    """the SQG provider (slb200/quasigeostrophic.py) against the oracle's, through the test double"""
    import slb200 as S
    from test_oracle_unsplit2d import sqg_run

    g = sqg_run(S, S.getgeovar, 6, getattr(S, alg), ordalg, sz=(32, 48))
    o = sqg_run(R, U.getgeovar, 6, getattr(R, alg), ordalg, sz=(32, 48))
    assert relerr(g.getdata(), o.data) <= 1e-11
    assert relerr(g.bufcur.to_host(), o.bufcur) <= 1e-9
    assert "poisson_solve_2d" in fake.calls


def test_const_shift_state_with_ndims_3_matches_the_tensor_stencil_oracle(fake):
    """a const-shift state with ndims = 3 (the N-D tensor stencil of src/interpolation.jl:212-231 for N = 3) runs as
    three 1-D sweeps (a fused pair + one sweep on the GPU); host sequencing checked over the test double"""
    import slb200 as S
    from oracle import refmodel as R

    def build(M):
        ms = (M.UniformMesh(0.0, 1.0, 12), M.UniformMesh(0.0, 1.0, 10), M.UniformMesh(0.0, 1.0, 8))
        adv = M.Advection(ms, [M.Lagrange(5), M.Lagrange(5), M.Lagrange(3)], 0.01, [([3, 1, 2], 3, 1, True)], tab_coef=[0.01])
        rng = np.random.default_rng(9)
        return M.AdvectionData(adv, np.asfortranarray(rng.random((12, 10, 8))), M.gettranslationvar((130.0, -270.0, 55.0)))

    g, o = build(S), build(R)
    for _ in range(2):
        assert S.advection(g) == R.advection(o)
    a, b = g.getdata(), o.data
    assert float(np.max(np.abs(a - b)) / np.max(np.abs(b))) <= 1e-13


def _sqg_split(M, sz, order, split, nbdt, t_max=10000.0):
    """the split form of test/test_quasigeostrophic.jl:45-58: tabst = [([1, 2], 1, 1, false), ([2, 1], 1, 2, false)]"""
    mx, my = M.UniformMesh(0.0, 1e6, sz[0]), M.UniformMesh(0.0, 1e6, sz[1])
    dt = t_max / nbdt
    adv = M.Advection((mx, my), [M.Lagrange(order), M.Lagrange(order)], dt, [([1, 2], 1, 1, False), ([2, 1], 1, 2, False)],
                      tab_coef=split(dt))
    pv = M.getgeovar(adv)
    advd = M.AdvectionData(adv, np.zeros(sz), pv)
    pv.initdata(advd)
    return advd


@pytest.mark.parametrize("split", ["standardsplit", "strangsplit"])
def test_quasigeostrophic_split_form_host_logic(fake, split):
    """split states with per-point shifts (src/advection.jl:633-645, src/quasigeostrophic.jl:126-135): the host
    sequencing of the product (displacement plane of the swept dim + exact identity along the other dim, through the
    per-point 2-D entry point) against the oracle's line-by-line restatement, over the test double"""
    import slb200 as S
    from oracle import refmodel as R
    from oracle import unsplit2d as U

    class MO:   # the oracle's module surface for this driver
        UniformMesh, Lagrange, Advection, AdvectionData = R.UniformMesh, R.Lagrange, R.Advection, R.AdvectionData
        getgeovar = staticmethod(U.getgeovar)

    g = _sqg_split(S, (32, 24), 5, getattr(S, split), 4)
    o = _sqg_split(MO, (32, 24), 5, getattr(R, split), 4)
    assert g.adv.nbstates == o.adv.nbstates
    for _ in range(2):
        more = True
        while more:
            more = S.advection(g)
            assert more == R.advection(o)
            a, b = g.getdata(), o.data
            assert float(np.max(np.abs(a - b)) / np.max(np.abs(b))) <= 1e-12
    assert g.time_cur == o.time_cur
