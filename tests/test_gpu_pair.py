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
    tB, sB = _alpha(rng, shape, dimB, mode)
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
