"""world_size-2 gloo tests (CPU) of the halo-sharded driver's host conventions (slb200/sharded.py): which planes go to
which neighbour's halo, the periodic wrap at the ends of the ring, halo widths, and the one-off handle exchange the
host language performs.  The passes themselves need a GPU (tests/test_gpu_halo.py runs them with in-process ranks)."""
import os
import socket

import numpy as np
import pytest

from slb200 import sharded as H


def test_halo_width_and_destinations():
    # order/2 + 1 planes for shifts below one cell, one more per extra cell
    assert H.halo_width(7, 1.0) == 4 and H.halo_width(7, 0.3) == 4 and H.halo_width(7, 1.5) == 5
    assert H.halo_width(3, 1.0) == 2 and H.halo_width(11, 2.0) == 7
    assert H.halo_destinations(0, 4, 8, 4) == [((4, 8), 3, (12, 16)), ((8, 12), 1, (0, 4))]
    assert H.halo_destinations(1, 2, 16, 4) == [((4, 8), 0, (20, 24)), ((16, 20), 0, (0, 4))]
    # a single rank is its own neighbour on both sides: the halos are the periodic wrap
    assert H.halo_destinations(0, 1, 16, 4) == [((4, 8), 0, (20, 24)), ((16, 20), 0, (0, 4))]


def test_slab_with_halos_is_periodic():
    g = np.arange(2 * 12, dtype=np.float64).reshape(2, 12)
    s = H.slab_with_halos(g, 0, 3, 2)
    assert s.shape == (2, 8) and list(s[0]) == [10, 11, 0, 1, 2, 3, 4, 5]
    s = H.slab_with_halos(g, 2, 3, 2)
    assert list(s[0]) == [6, 7, 8, 9, 10, 11, 0, 1]


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shape, Hw = (3, 4, 5, 6 * world + (4 if world == 2 else 0)), 3
        g = np.asfortranarray(np.random.default_rng(5).random(shape))
        c = shape[3] // world
        loc = np.zeros(shape[:3] + (c + 2 * Hw,), order="F")
        loc[..., Hw:Hw + c] = g[..., rank * c:(rank + 1) * c]
        H.exchange_halos_reference(loc, rank, world, Hw, dist, torch)
        ok1 = np.array_equal(loc, H.slab_with_halos(g, rank, world, Hw))
        blobs = H.torch_allgather_bytes(dist)(bytes([rank]) * 8)
        ok2 = blobs == [bytes([r]) * 8 for r in range(world)]
        q.put((rank, bool(ok1), bool(ok2)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_and_handle_allgather_gloo(world):
    import torch.multiprocessing as mp

    with socket.socket() as sck:
        sck.bind(("127.0.0.1", 0))
        port = sck.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True, True) for r in range(world)]


def test_estimate_max_shift_sizes_the_halo_from_the_initial_field():
    """halo sizing of the sharded driver (host numpy restatement of compute_elfield!): the 2D2V example's initial
    field gives velocity shifts of 2/3 cell at 128 points per velocity dim (Strang: dt/2 per v stage) -- halo 4 for
    order 7 -- and twice that at 256 points -- halo 6 for order 9"""
    import math

    import slb200 as S

    for n, order, amax, halo in ((128, 7, 2 / 3, 4), (256, 9, 4 / 3, 6)):
        ms = (S.UniformMesh(0.0, 4 * math.pi, n), S.UniformMesh(0.0, 4 * math.pi, n), S.UniformMesh(-6.0, 6.0, n), S.UniformMesh(-6.0, 6.0, n))
        tabst = [([3, 4, 1, 2], 1, 1, True), ([4, 3, 1, 2], 1, 1, True), ([1, 2, 4, 3], 1, 2, True), ([2, 1, 3, 4], 1, 2, True)]
        adv = S.Advection(ms, [S.Lagrange(order)] * 4, 0.1, tabst)
        fsp = lambda x: 0.5 * np.cos(x / 2) + 1
        fv = lambda v: np.exp(-v**2 / 2) / math.sqrt(2 * math.pi)
        dv = ms[2].step * ms[3].step
        rho = np.multiply.outer(fsp(ms[0].points), fsp(ms[1].points)) * (fv(ms[2].points).sum() * fv(ms[3].points).sum() * dv)
        raw = H.estimate_max_shift(adv, rho, margin=1.0)
        assert abs(raw - max(1.0, amax)) < 2e-3
        assert H.halo_width(order, H.estimate_max_shift(adv, rho)) == halo
