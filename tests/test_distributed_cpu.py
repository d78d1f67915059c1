"""world_size-2 gloo tests (CPU) of the sharded driver's host logic: slab ranges, block-major
packing and the two all-to-all re-shards.  The data path itself (sweeps) needs a GPU; here
the sweeps are replaced by identity copies so that only the layout algebra is exercised --
the same conventions slb_sweep_ex's SLB_RESHARD_* modes implement on the device
(tests/test_gpu_distributed.py checks those against to_block_major / from_block_major)."""
import os
import socket

import numpy as np
import pytest

from slb200 import distributed as D


def test_splititr_splitvec_known_answers():
    # test/test_util.jl:14-43 (the reference's work split for its MPI back-end)
    assert D.splititr(3, 15) == [(1, 5), (6, 10), (11, 15)]
    assert D.splititr(3, 11) == [(1, 4), (5, 8), (9, 11)]
    assert D.splititr(5, 24) == [(1, 5), (6, 10), (11, 15), (16, 20), (21, 24)]
    assert D.splitvec(3, list(range(47, 62))) == [list(range(47, 52)), list(range(52, 57)), list(range(57, 62))]
    assert D.splitvec(4, list(range(34, 39))) == [[34, 35], [36], [37], [38]]
    v = list(range(1, 54))
    t = D.splitvec(5, v)
    assert [len(x) for x in t] == [11, 11, 11, 10, 10] and t[4] == list(range(44, 54))


def test_slab_and_block_major_roundtrip():
    assert D.slab(128, 4, 1) == (32, 64)
    with pytest.raises(ValueError):
        D.slab(10, 4, 0)
    assert D.local_shape((8, 12, 6, 4), "A", 2) == (8, 12, 6, 2)
    assert D.local_shape((8, 12, 6, 4), "B", 2) == (8, 6, 6, 4)
    rng = np.random.default_rng(0)
    a = np.asfortranarray(rng.random((5, 12, 3, 4)))
    for bdim, nb in ((1, 3), (1, 4), (3, 2), (2, 3)):
        flat = D.to_block_major(a, bdim, nb)
        assert flat.shape == (a.size,)
        assert np.array_equal(D.from_block_major(flat, a.shape, bdim, nb), a)
    # block-major along the slowest dim is the plain layout
    assert np.array_equal(D.to_block_major(a, 3, 2), a.reshape(-1, order="F"))


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shape = (6, 8, 5, 4)
        g = np.asfortranarray(np.arange(np.prod(shape), dtype=np.float64).reshape(shape, order="F"))
        lo2, hi2 = D.slab(shape[1], world, rank)
        lo4, hi4 = D.slab(shape[3], world, rank)
        local_B = np.asfortranarray(g[:, lo2:hi2, :, :])
        local_A = np.asfortranarray(g[:, :, :, lo4:hi4])
        got_A = D.reshard_B_to_A_reference(local_B, world, dist, torch)
        ok1 = np.array_equal(got_A, local_A)
        got_B = D.reshard_A_to_B_reference(local_A, world, dist, torch)
        ok2 = np.array_equal(got_B, local_B)
        # rho slabs: all-gather of [n1, n2/P] slabs is the full [n1, n2] array
        rho_l = torch.from_numpy(np.ascontiguousarray(g[:, lo2:hi2, 0, 0].reshape(-1, order="F")))
        rho = torch.empty(shape[0] * shape[1], dtype=torch.float64)
        dist.all_gather_into_tensor(rho, rho_l)
        ok3 = np.array_equal(rho.numpy().reshape(shape[:2], order="F"), g[:, :, 0, 0])
        q.put((rank, ok1, ok2, ok3))
    finally:
        dist.destroy_process_group()


def test_reshard_roundtrip_gloo_world2():
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok1, ok2, ok3 in res:
        assert ok1 and ok2 and ok3, (rank, ok1, ok2, ok3)
