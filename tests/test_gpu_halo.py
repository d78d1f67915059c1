"""Halo-sharded 2D2V path (slb_sweep_pair_halo, slb_comm_*, slb200/sharded.py) on ONE GPU: P ranks are built inside
this process (each with its own context and stream, wired to each other's buffers directly), so the kernels, the
peer stores, the flag protocol and the host sequencing are the ones a multi-GPU run uses -- only the memory the
"peer" pointers name is local.  Checks:
  * every sharded pass is BIT-IDENTICAL to slb_sweep_pair on the unsharded grid, and the halo planes it pushes are
    the neighbours' boundary planes;
  * whole Strang steps of the sharded driver against the single-grid driver and the oracle (rho is summed in a
    different order across ranks, hence rounding-level differences);
  * a shift beyond the halo is reported, not silently wrong.
"""
import ctypes as C
import math

import numpy as np
import pytest

from helpers import relerr

pytestmark = pytest.mark.gpu


def _adv(S, sz, order, dt=0.1):
    ms = (S.UniformMesh(0.0, 4 * math.pi, sz[0]), S.UniformMesh(0.0, 4 * math.pi, sz[1]), S.UniformMesh(-6.0, 6.0, sz[2]), S.UniformMesh(-6.0, 6.0, sz[3]))
    tabst = [([3, 4, 1, 2], 1, 1, True), ([4, 3, 1, 2], 1, 1, True), ([1, 2, 4, 3], 1, 2, True), ([2, 1, 3, 4], 1, 2, True)]
    adv = S.Advection(ms, [S.Lagrange(order)] * 4, dt, tabst)
    fsp = lambda x: 0.5 * np.cos(x / 2) + 1
    fv = lambda v: np.exp(-v**2 / 2) / math.sqrt(2 * math.pi)
    f = S.dotprod((fsp(ms[0].points), fsp(ms[1].points), fv(ms[2].points), fv(ms[3].points)))
    return adv, f


def _gather(ranks):
    return np.asfortranarray(np.concatenate([s.getdata_local() for s in ranks], axis=3))


@pytest.mark.parametrize("P,order,sz", [(1, 7, (32, 8, 16, 16)), (2, 7, (32, 12, 16, 32)), (4, 7, (32, 8, 16, 32)), (2, 3, (64, 6, 8, 16)),
                                        (4, 11, (32, 6, 12, 48)), (2, 9, (16, 10, 12, 24))])
def test_halo_passes_bitwise_equal_unsharded(P, order, sz):
    """v1 v2 (windowed march + row pushes) and x1 x2 (slab view + plane pushes) passes with given shift tables,
    rank by rank, against slb_sweep_pair on the whole grid."""
    import slb200 as S
    from helpers import DeviceGrid
    from slb200 import _lib
    from slb200.sharded import HaloShardedAdvectionData, local_group

    rng = np.random.default_rng(11 + P + order)
    n1, n2, n3, n4 = sz
    f = np.asfortranarray(rng.random(sz))
    adv, _ = _adv(S, sz, order)
    it = adv.t_interp[0]
    E1 = rng.uniform(-0.999, 0.999, n1 * n2)
    E2 = rng.uniform(-0.999, 0.999, n1 * n2)
    E2[:4] = [-1.0, 0.0, 0.999999, -0.5]   # both ends of the admissible range
    # single grid: v1 v2 pass, then x1 x2 pass, then v1 v2 again
    vx1 = np.linspace(-6.3, 6.1, n3)
    vx2 = np.linspace(-5.7, 6.4, n4)
    ref = DeviceGrid(f)
    stages = []
    ref.sweep_pair(2, it, E1, [1, n1, 0, 0], 3, it, E2, [1, n1, 0, 0])
    stages.append(ref.get())
    ref.sweep_pair(0, it, vx1, [0, 0, 1, 0], 1, it, vx2, [0, 0, 0, 1])
    stages.append(ref.get())
    ref.sweep_pair(2, it, E2, [1, n1, 0, 0], 3, it, E1, [1, n1, 0, 0])
    stages.append(ref.get())
    ref.close()
    # sharded: the driver's own pass routine with the tables swapped in
    ranks = local_group(adv, f, P)
    L = _lib.lib()
    for s in ranks:
        s.has_field = True
        s.ctx_tabs = [s.ctx.to_device(t) for t in (E1, E2, vx1, vx2)]

    def run(kind, k):
        for s in ranks:
            tE1, tE2, tv1, tv2 = s.ctx_tabs
            if kind == "v":
                a, b = (tE1, tE2) if k == 0 else (tE2, tE1)
                s.E_dev_saved = s.E_dev
                s.E_dev = [a, b]
                s._pass(2, adv.t_mesh[2].step, 3, adv.t_mesh[3].step, 2 if k == 2 else 0)   # scale dt/step = 1
                s.E_dev = s.E_dev_saved
            else:
                s.points_saved = s.points
                s.points = [None, None, tv1, tv2]
                s._pass(0, -adv.t_mesh[0].step, 1, -adv.t_mesh[1].step, 2)
                s.points = s.points_saved
        for s in ranks:   # order the pushes of this pass against the next pass (the driver's all-gather does it; the half
            s._consume_copy()   # that travels on the second stream is awaited by the next v pass -- here, explicitly)
            s.sync_ranks()

    run("v", 0)
    assert np.array_equal(_gather(ranks), stages[0])
    run("x", 1)
    got = _gather(ranks)
    assert np.array_equal(got, stages[1])
    # the x pass pushed its boundary planes: every rank's halos now hold the neighbours' planes
    H, c = ranks[0].H, ranks[0].c
    for s in ranks:
        full = s.ctx.to_host(C.c_void_p(s.ptr[s.cur]), s.nhalo).reshape((n1, n2, n3, c + 2 * H), order="F")
        lo = [(s.rank * c - H + j) % n4 for j in range(H)]
        hi = [((s.rank + 1) * c + j) % n4 for j in range(H)]
        assert np.array_equal(full[..., :H], stages[1][..., lo])
        assert np.array_equal(full[..., H + c:], stages[1][..., hi])
    run("v", 2)
    assert np.array_equal(_gather(ranks), stages[2])
    for s in ranks:   # ... and so did the v pass that announced another v pass after it
        full = s.ctx.to_host(C.c_void_p(s.ptr[s.cur]), s.nhalo).reshape((n1, n2, n3, c + 2 * H), order="F")
        lo = [(s.rank * c - H + j) % n4 for j in range(H)]
        hi = [((s.rank + 1) * c + j) % n4 for j in range(H)]
        assert np.array_equal(full[..., :H], stages[2][..., lo])
        assert np.array_equal(full[..., H + c:], stages[2][..., hi])
    for s in ranks:
        s.check()
        for t in s.ctx_tabs:
            s.ctx.free(t)
        s.close()


@pytest.mark.parametrize("split", ["0", "1"])
@pytest.mark.parametrize("P,order,sz,nsteps", [(2, 7, (32, 16, 16, 32), 3), (4, 7, (32, 8, 16, 32), 2), (1, 5, (32, 8, 16, 16), 2),
                                                (4, 9, (16, 16, 20, 40), 2)])
def test_halo_sharded_steps_match_single_grid_and_oracle(P, order, sz, nsteps, split, monkeypatch):
    """split = "1": a pushing pass sends its low-side planes itself and leaves the high side to a peer copy on a second
    stream (signalled to the neighbour, awaited before its next v pass); "0": both sides inside the pass."""
    import slb200 as S

    monkeypatch.setenv("SLB_HALO_SPLIT_PUSH", split)
    from oracle import refmodel as R
    from slb200.sharded import local_group

    adv, f = _adv(S, sz, order)
    plain = S.AdvectionData(adv, f, S.getpoissonvar(adv))
    adv_o, _ = _adv(R, sz, order)
    orc = R.AdvectionData(adv_o, f, R.getpoissonvar(adv_o))
    ranks = local_group(adv, f, P)
    for step in range(nsteps):
        while S.advection(plain):
            pass
        while R.advection(orc):
            pass
        more = True
        while more:
            res = [s.advection() for s in ranks]
            assert len(set(res)) == 1
            more = res[0]
        ee_p, ee_o = S.compute_ee(plain), R.compute_ee(orc)
        ee_s = [s.compute_ee() for s in ranks]
        assert len(set(ee_s)) == 1, "the replicated field solve must be identical on every rank"
        assert abs(ee_s[0] - ee_p) <= 1e-13 * abs(ee_p)
        assert abs(ee_s[0] - ee_o) <= 1e-11 * abs(ee_o)
        g = _gather(ranks)
        assert relerr(g, plain.getdata()) <= 1e-13 * (step + 1)
        assert relerr(g, orc.data) <= 1e-12 * 6 * (step + 1)
    assert all(s.n_fused == 3 * nsteps for s in ranks)
    for s in ranks:
        s.close()


def test_shift_beyond_the_halo_is_reported():
    import slb200 as S
    from slb200 import _lib
    from slb200.sharded import local_group

    sz = (32, 8, 16, 32)
    adv, f = _adv(S, sz, 7)
    ranks = local_group(adv, f, 2)
    s = ranks[0]
    big = s.ctx.to_device(np.full(sz[0] * sz[1], 1.5))   # floor(alpha) = 1: the v2 stencil needs one more halo plane
    s.has_field = True
    saved, s.E_dev = s.E_dev, [big, big]
    s._pass(2, adv.t_mesh[2].step, 3, adv.t_mesh[3].step, 0)
    s.E_dev = saved
    with pytest.raises(_lib.SlbError, match="halo"):
        s.check()
    s.check()   # reading the flag cleared it
    s.ctx.free(big)
    for r in ranks:
        r.close()


def test_halo_driver_refuses_what_it_does_not_cover():
    import slb200 as S
    from slb200.sharded import HaloShardedAdvectionData, HaloUnsupported

    sz = (16, 8, 8, 16)
    ms = (S.UniformMesh(0.0, 4 * math.pi, sz[0]), S.UniformMesh(0.0, 4 * math.pi, sz[1]), S.UniformMesh(-6.0, 6.0, sz[2]), S.UniformMesh(-6.0, 6.0, sz[3]))
    tabst = [([3, 4, 1, 2], 1, 1, True), ([4, 3, 1, 2], 1, 1, True), ([1, 2, 4, 3], 1, 2, True), ([2, 1, 3, 4], 1, 2, True)]
    adv = S.Advection(ms, [S.BSplineLU(5, n) for n in sz], 0.1, tabst)
    with pytest.raises(HaloUnsupported):
        HaloShardedAdvectionData(adv, np.zeros((16, 8, 8, 16)), 0, 1)
    adv = S.Advection(ms, [S.Lagrange(7)] * 4, 0.1, tabst)
    with pytest.raises(HaloUnsupported):   # slab of 4 planes, halo of 4
        HaloShardedAdvectionData(adv, np.zeros((16, 8, 8, 4)), 0, 4, allgather_bytes=lambda b: [b] * 4, max_shift=1.0)


def test_streamed_io_exchange_between_steps():
    """read-back of a step's result overlapped with the upload of the next step's input (the end-to-end leg of
    bench.py): the downloaded slabs are the step's result, and the next step starts from the uploaded data."""
    import slb200 as S
    from slb200 import _lib
    from slb200.sharded import local_group

    sz = (32, 8, 16, 32)
    adv, f = _adv(S, sz, 7)
    plain = S.AdvectionData(adv, f, S.getpoissonvar(adv))
    while S.advection(plain):
        pass
    want = plain.getdata()
    ranks = local_group(adv, f, 2)
    c = ranks[0].c
    outs, ins = [], []
    for s in ranks:
        o, _p = _lib.pinned_empty((s.plane * c,))
        i, _q = _lib.pinned_empty((s.plane * c,))
        i[:] = np.asfortranarray(f[..., s.rank * c:(s.rank + 1) * c]).reshape(-1, order="F")
        outs.append(o)
        ins.append(i)
    for rep in range(2):   # both repetitions start from f: the second one from the uploaded copy
        more = True
        while more:
            more = [s.advection() for s in ranks][0]
        for s, o, i in zip(ranks, outs, ins):
            s.stream_io_exchange(o, i)
        for s in ranks:
            s.ctx.sync()
            s._io["down"].sync()
        got = np.concatenate([o.reshape((sz[0], sz[1], sz[2], c), order="F") for o in outs], axis=3)
        assert relerr(got, want) <= 1e-13, rep
    for s in ranks:
        s.check()
        s.close()
