"""GPU checks of the re-shard-fused sweeps (slb_sweep_ex) on ONE device, plus the sharded
driver with a single rank (P = 1 degenerates to the plain driver).  Multi-rank runs are
exercised by bench.py --gpus N and tools/check_sharded.py under torchrun."""
import ctypes as C
import math

import numpy as np
import pytest

from helpers import DeviceGrid, make_pair, oracle_sweep, relerr

pytestmark = pytest.mark.gpu


def _sweep_ex(f_flat_or_arr, shape, dim, interp, tab, astride, mode, bdim, nblocks):
    from slb200 import _lib

    ctx = _lib.default_context()
    g = DeviceGrid(np.zeros(shape, order="F"))
    flat = np.ascontiguousarray(f_flat_or_arr.reshape(-1, order="F"))
    _lib.check(_lib.lib().slb_grid_upload(g.h, flat.ctypes.data_as(C.c_void_p)))
    tab = np.ascontiguousarray(tab, dtype=np.float64)
    h = interp.handle(ctx, shape[dim])
    _lib.check(_lib.lib().slb_sweep_ex(g.h, dim, h, tab.ctypes.data_as(C.c_void_p), tab.size, _lib.i64(astride), 1.0, 0, 0,
                                       mode, bdim, nblocks))
    out = g.get()
    g.close()
    return out


@pytest.mark.parametrize("kind,order", [("lagrange", 7), ("bspline_lu", 5), ("bspline_fft", 11)])
@pytest.mark.parametrize("nblocks", [2, 4, 8])
def test_out_blocked_equals_block_major_of_plain_sweep(nblocks, kind, order):
    from slb200 import distributed as D, _lib

    rng = np.random.default_rng(1)
    shape = (32, 64, 6, 4)
    f = np.asfortranarray(rng.random(shape))
    interp, ointerp = make_pair(kind, order, shape[1])
    tab = rng.uniform(-6, 6, shape[3])
    ref = oracle_sweep(f, 1, ointerp, tab, [0, 0, 0, 1])
    out = _sweep_ex(f, shape, 1, interp, tab, [0, 0, 0, 1], _lib.SLB_RESHARD_OUT_BLOCKED, 1, nblocks)
    got = D.from_block_major(out.reshape(-1, order="F"), shape, 1, nblocks)
    assert relerr(got, ref) <= 1e-12


@pytest.mark.parametrize("kind,order", [("lagrange", 7), ("bspline_lu", 5), ("bspline_fft", 11)])
@pytest.mark.parametrize("nblocks", [2, 4, 8])
def test_in_blocked_reads_block_major_input(nblocks, kind, order):
    from slb200 import distributed as D, _lib

    rng = np.random.default_rng(2)
    shape = (128, 16, 5, 3)
    f = np.asfortranarray(rng.random(shape))
    interp, ointerp = make_pair(kind, order, shape[0])
    tab = rng.uniform(-6, 6, shape[2])
    ref = oracle_sweep(f, 0, ointerp, tab, [0, 0, 1, 0])
    fb = D.to_block_major(f, 1, nblocks)
    out = _sweep_ex(fb, shape, 0, interp, tab, [0, 0, 1, 0], _lib.SLB_RESHARD_IN_BLOCKED, 1, nblocks)
    assert relerr(out, ref) <= 1e-12
    # blocked along dim 2 (L > 1 path)
    fb2 = D.to_block_major(f, 2, 5)
    out2 = _sweep_ex(fb2, shape, 0, interp, tab, [0, 0, 1, 0], _lib.SLB_RESHARD_IN_BLOCKED, 2, 5)
    assert relerr(out2, ref) <= 1e-12


def _pair_ex(flat_in, shape, dimA, dimB, interp, tabA, strA, tabB, strB, in_nblocks, out_nblocks, flags=0, first_block=0):
    from slb200 import _lib

    ctx = _lib.default_context()
    g = DeviceGrid(np.zeros(shape, order="F"))
    flat = np.ascontiguousarray(flat_in.reshape(-1, order="F"))
    _lib.check(_lib.lib().slb_grid_upload(g.h, flat.ctypes.data_as(C.c_void_p)))
    tabA, tabB = np.ascontiguousarray(tabA, dtype=np.float64), np.ascontiguousarray(tabB, dtype=np.float64)
    hA, hB = interp.handle(ctx, shape[dimA]), interp.handle(ctx, shape[dimB])
    _lib.check(_lib.lib().slb_sweep_pair_ex(
        g.h, dimA, hA, tabA.ctypes.data_as(C.c_void_p), tabA.size, _lib.i64(strA), 1.0,
        dimB, hB, tabB.ctypes.data_as(C.c_void_p), tabB.size, _lib.i64(strB), 1.0, 0, flags, in_nblocks, out_nblocks, None, first_block))
    out = g.get()
    g.close()
    return out


@pytest.mark.parametrize("in_nb,out_nb", [(1, 2), (1, 8), (4, 1), (2, 2), (8, 4)])
@pytest.mark.parametrize("case", ["v1v2", "x1x2"])
def test_pair_ex_block_major_input_and_output(case, in_nb, out_nb):
    """The fused pair with the re-shard maps: block-major input / output along the march dim equal
    the numpy layout algebra applied to the plain fused result (bit for bit)."""
    from slb200 import distributed as D

    rng = np.random.default_rng(3)
    if case == "v1v2":   # layout B slab: sweeps along dims 2, 3; blocks along v2 (B -> A exchange)
        shape, dA, dB = (32, 6, 20, 16), 2, 3
        tA, sA = rng.uniform(-2, 2, 32 * 6), [1, 32, 0, 0]
        tB, sB = rng.uniform(-2, 2, 32 * 6), [1, 32, 0, 0]
    else:                # layout A slab: sweeps along dims 0, 1; blocks along x2 (A -> B exchange)
        shape, dA, dB = (48, 16, 5, 3), 0, 1
        tA, sA = rng.uniform(-6, 6, 5), [0, 0, 1, 0]
        tB, sB = rng.uniform(-6, 6, 3), [0, 0, 0, 1]
    f = np.asfortranarray(rng.random(shape))
    interp, _ = make_pair("lagrange", 7, shape[dA])
    for flags in (0, 1):
        plain = _pair_ex(f, shape, dA, dB, interp, tA, sA, tB, sB, 1, 1, flags)
        fin = D.to_block_major(f, dB, in_nb) if in_nb > 1 else f
        out = _pair_ex(fin, shape, dA, dB, interp, tA, sA, tB, sB, in_nb, out_nb, flags)
        got = D.from_block_major(out.reshape(-1, order="F"), shape, dB, out_nb) if out_nb > 1 else out
        assert np.array_equal(got, plain), (case, in_nb, out_nb, flags)
        if out_nb > 1:
            # rotated march start (what rank r of P uses): same result
            out = _pair_ex(fin, shape, dA, dB, interp, tA, sA, tB, sB, in_nb, out_nb, flags, first_block=out_nb - 1)
            got = D.from_block_major(out.reshape(-1, order="F"), shape, dB, out_nb)
            assert np.array_equal(got, plain), (case, in_nb, out_nb, flags, "rotated")


def test_sharded_driver_single_rank_matches_plain_driver():
    """P = 1: the sharded driver must reproduce the plain AdvectionData history."""
    import os

    import torch
    import torch.distributed as dist

    import slb200 as S
    from slb200.distributed import ShardedAdvectionData

    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29531")
        dist.init_process_group("nccl", rank=0, world_size=1)
    n = 16
    ms = (S.UniformMesh(0.0, 4 * math.pi, n), S.UniformMesh(0.0, 4 * math.pi, n), S.UniformMesh(-6.0, 6.0, n), S.UniformMesh(-6.0, 6.0, n))
    tabst = [([3, 4, 1, 2], 1, 1, True), ([4, 3, 1, 2], 1, 1, True), ([1, 2, 4, 3], 1, 2, True), ([2, 1, 3, 4], 1, 2, True)]
    adv = S.Advection(ms, [S.Lagrange(7)] * 4, 0.1, tabst)
    fsp = lambda x: 0.5 * np.cos(x / 2) + 1
    fv = lambda v: np.exp(-v**2 / 2) / math.sqrt(2 * math.pi)
    f = S.dotprod((fsp(ms[0].points), fsp(ms[1].points), fv(ms[2].points), fv(ms[3].points)))
    plain = S.AdvectionData(adv, f, S.getpoissonvar(adv))
    sh = ShardedAdvectionData(adv, f)
    for _ in range(2):
        while S.advection(plain):
            pass
        while sh.advection():
            pass
        assert abs(sh.compute_ee() - S.compute_ee(plain)) <= 1e-13 * abs(S.compute_ee(plain))
    assert relerr(sh.gather_global(), plain.getdata()) <= 1e-13
    sh.close()
    dist.destroy_process_group()
