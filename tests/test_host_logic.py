"""Host-side logic of the product mirror (no GPU): state schedule, splitting tables, meshes,
weight tables and the bordered-LU B-spline solver, checked against the oracle and against the
reference's own known answers (test/test_advection.jl:58-158, test/test_mesh.jl:57-67,
test/test_util.jl:14-43)."""
import ctypes as C
import math
import os

import numpy as np
import pytest

import slb200 as S
from oracle import refmodel as R, tables as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("M", [S, R], ids=["product", "oracle"])
def test_state_schedule_6d(M):
    """test/test_advection.jl:58-158: 3D3V grid, Strang, space states have stcoef 1."""
    szsp, szv = (2, 4, 8), (4, 8, 4)
    ms = tuple(M.UniformMesh(-1.0, 3.0, n) for n in szsp) + tuple(M.UniformMesh(-3.0, 1.0, n) for n in szv)
    states = [([1, 2, 3, 6, 5, 4], 1, 1, True), ([2, 1, 3, 4, 6, 5], 1, 1, True), ([3, 2, 1, 4, 5, 6], 1, 1, True),
              ([4, 5, 6, 1, 2, 3], 1, 2, True), ([5, 4, 6, 1, 2, 3], 1, 2, True), ([6, 5, 4, 1, 2, 3], 1, 2, True)]
    adv = M.Advection(ms, [M.Lagrange(3) for _ in range(6)], 0.125, states)
    assert adv.sizeall == szsp + szv
    assert adv.nbstates == 9 and adv.maxcoef == 2
    assert adv.getcur_t(1) == adv.tab_coef[0] == adv.tab_coef[2]
    assert adv.getcur_t(4) == adv.tab_coef[1]
    t_coef = [1, 1, 1, 2, 2, 2, 3, 3, 3, 1]
    t_indice = [1, 2, 3, 4, 5, 6, 1, 2, 3, 1]
    t_result = [True] * 8 + [False, True]

    class Dummy:  # the state machine alone: no device data needed
        pass

    advd = M.AdvectionData.__new__(M.AdvectionData)
    advd.adv, advd.state_gen, advd.time_cur = adv, 1, 0.0
    for i in range(10):
        assert advd.getstcoef() == t_coef[i]
        assert advd.state_gen == i % 9 + 1
        assert advd._getcurrentindice() == t_indice[i]
        assert advd.getcur_t() == adv.tab_coef[t_coef[i] - 1]
        assert advd.getinterp()[0] is adv.t_interp[t_indice[i] - 1]
        assert advd.nextstate() == t_result[i]
    assert advd.time_cur == 0.125


def test_schedules_of_the_baseline_configs():
    m = S.UniformMesh(0.0, 1.0, 8)
    it = S.Lagrange(3)
    a1 = S.Advection((m, m), [it, it], 0.1, [([1, 2], 1, 1, True), ([2, 1], 1, 2, True)])
    assert a1.nbstates == 3
    assert [(a1.getst(g).perm[0], a1.getcur_t(g)) for g in (1, 2, 3)] == [(1, 0.05), (2, 0.1), (1, 0.05)]
    tabst = [([3, 4, 1, 2], 1, 1, True), ([4, 3, 1, 2], 1, 1, True), ([1, 2, 4, 3], 1, 2, True), ([2, 1, 3, 4], 1, 2, True)]
    a2 = S.Advection((m,) * 4, [it] * 4, 0.1, tabst)
    assert a2.nbstates == 6
    assert [(a2.getst(g).perm[0], a2.getcur_t(g)) for g in range(1, 7)] == [(3, 0.05), (4, 0.05), (1, 0.1), (2, 0.1), (3, 0.05), (4, 0.05)]
    with pytest.raises(ValueError):
        S.Advection((m, m), [it], 0.1, [([1, 2], 1, 1, True)])


def test_splitting_tables():
    dt = 0.1
    for name in ("nosplit", "standardsplit", "strangsplit", "magicsplit", "triplejumpsplit"):
        assert getattr(S, name)(dt) == getattr(R, name)(dt), name
    assert S.strangsplit(dt) == [0.05, 0.1, 0.05]
    assert S.magicsplit(dt) == [math.tan(0.05), math.sin(0.1), math.tan(0.05)]
    tj = S.triplejumpsplit(dt)
    assert len(tj) == 7 and abs(sum(tj[0::2]) - dt) < 1e-15 and abs(sum(tj[1::2]) - dt) < 1e-15
    o6 = S.order6split(dt)
    assert len(o6) == 23 and abs(sum(o6) - 2 * dt) < 1e-15 and o6 == o6[::-1]
    h = S.hamsplit_3_11(dt)
    assert len(h) == 11 and h == h[::-1]
    assert abs(sum(h[1::2]) - dt) < 1e-15  # the a_i sum to 1


def test_mesh_and_wavenumbers():
    # test/test_mesh.jl:57-67
    for M in (S, R):
        mesh = M.UniformMesh(-1.0, 1.0, 64)
        ref = np.roll(np.arange(-32, 32), 32) * (2 * math.pi / 2.0)
        assert np.array_equal(M.vec_k_fft(mesh), ref)
        assert mesh.step == 2.0 / 64 and mesh.width == 2.0
        assert np.array_equal(mesh.points, -1.0 + np.arange(64) / 32)
    a, b = S.UniformMesh(0.0, 4 * math.pi, 128), R.UniformMesh(0.0, 4 * math.pi, 128)
    assert np.array_equal(a.points, b.points) and a.step == b.step
    assert S.stop(a) == 4 * math.pi or abs(S.stop(a) - 4 * math.pi) < 1e-14
    c = S.UniformMesh(-6.0, 4.5, 1022)  # non-dyadic: exact rational nodes rounded once
    assert abs(c.points[-1] + c.step - 4.5) < 1e-14


def test_weight_tables_match_oracle_tables():
    from slb200 import interp as I

    for o in range(1, 14):
        assert np.array_equal(S.Lagrange(o).tabfct, np.array(T.to_float64_table(T.lagrange_tabfct_rat(o))))
    for o in (3, 5, 7, 9, 11, 13):
        b = S.BSplineLU(o, 64)
        assert np.array_equal(b.tabfct, np.array(T.to_float64_table(T.bspline_tabfct_rat(o))))
        assert list(b.nodes) == [float(x) for x in T.bspline_node_values_rat(o)]
    for o in (5, 9, 13):
        assert np.array_equal(S.Hermite(o).tabfct, np.array(T.to_float64_table(T.hermite_tabfct_rat(o))))
    assert np.array_equal(S.Hermite(7, flbis=True).tabfct, np.array(T.to_float64_table(T.hermite_tabfct_rat(7, True))))
    assert I.get_kl_ku(5) == (2, 2) and I.get_kl_ku(6) == (2, 3)
    assert S.get_order(S.Lagrange(9)) == 9
    # host getprecal agrees with the oracle's FMA Horner to the last bits
    w = S.Lagrange(9).getprecal(0.3)
    assert np.max(np.abs(w - R.Lagrange(9).getprecal(0.3))) < 1e-15 and abs(w.sum() - 1) < 1e-14


@pytest.mark.parametrize("order,n", [(3, 16), (5, 128), (5, 1024), (9, 30), (11, 128), (11, 256), (13, 64)])
def test_bordered_lu_solver_host(order, n):
    """The exact routine the B-spline kernels run (slb_bspline.cuh: bspline_factor +
    bspline_solve_line), executed on the host, equals the reference's cyclic LU solve."""
    so = os.path.join(ROOT, "semilagrangian.jl_b200", "lib", "libslb200_hosttest.so")
    L = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    L.slbt_bspline_solve_host.argtypes = [C.c_int, C.c_longlong, dp, dp, dp]
    rng = np.random.default_rng(5431221)
    nodes = np.array([float(x) for x in T.bspline_node_values_rat(order)])
    b = rng.random(n)
    x = np.empty(n)
    assert L.slbt_bspline_solve_host(order, n, nodes.ctypes.data_as(dp), b.ctypes.data_as(dp), x.ctypes.data_as(dp)) == 0
    ref = R.BSplineLU(order, n).sol(b)
    assert np.max(np.abs(x - ref)) <= 1e-13 * np.max(np.abs(ref))


@pytest.mark.parametrize("order,n", [(3, 16), (3, 128), (5, 64), (5, 128), (7, 128), (9, 30), (9, 100), (11, 128), (11, 256), (13, 64)])
def test_split_line_solver_host(order, n):
    """The two-halves + two-separators formulation of the fused B-spline sweep (slb_bspsplit.cuh:
    bspsplit_factor, the kernel's table layout, the per-line arithmetic), executed on the host,
    equals the reference's cyclic LU solve."""
    so = os.path.join(ROOT, "semilagrangian.jl_b200", "lib", "libslb200_hosttest.so")
    L = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    L.slbt_bspsplit_solve_host.argtypes = [C.c_int, C.c_longlong, dp, dp, dp]
    rng = np.random.default_rng(77)
    nodes = np.array([float(x) for x in T.bspline_node_values_rat(order)])
    b = rng.random(n)
    x = np.empty(n)
    assert L.slbt_bspsplit_solve_host(order, n, nodes.ctypes.data_as(dp), b.ctypes.data_as(dp), x.ctypes.data_as(dp)) == 0
    ref = R.BSplineLU(order, n).sol(b)
    assert np.max(np.abs(x - ref)) <= 1e-13 * np.max(np.abs(ref))


@pytest.mark.parametrize("order,n", [(3, 16), (3, 128), (5, 64), (5, 128), (5, 1024), (7, 128), (9, 30), (9, 100), (11, 128), (11, 256), (13, 64), (13, 14)])
def test_recursive_filter_solver_host(order, n):
    """The constant-coefficient recursive-filter form of the periodic B-spline pre-solve
    (slb_bsprf.cuh: pole search, aliased start-up tables, per-line cascade), executed on the host,
    equals the reference's cyclic LU solve; odd and non-power-of-two n included."""
    so = os.path.join(ROOT, "semilagrangian.jl_b200", "lib", "libslb200_hosttest.so")
    L = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    L.slbt_bsprf_solve_host.argtypes = [C.c_int, C.c_longlong, dp, dp, dp, C.POINTER(C.c_int)]
    rng = np.random.default_rng(78)
    nodes = np.array([float(x) for x in T.bspline_node_values_rat(order)])
    b = rng.random(n)
    x = np.empty(n)
    K = (C.c_int * 6)()
    assert L.slbt_bsprf_solve_host(order, n, nodes.ctypes.data_as(dp), b.ctypes.data_as(dp), x.ctypes.data_as(dp), K) == 0
    ref = R.BSplineLU(order, n).sol(b)
    assert np.max(np.abs(x - ref)) <= 1e-13 * np.max(np.abs(ref))
    h = (order - 1) // 2
    assert all(1 <= K[k] <= n for k in range(h)) and list(K[:h]) == sorted(K[:h])


def test_driver_level_oracle_kats():
    """Oracle driver pinned by the reference's integration tests: rotation returns to the
    start (test/test_rotation.jl:237-252, err < 1e-3) and Poisson pieces are consistent
    (test/test_poisson.jl:38-108)."""
    sz = (200, 150)
    mx, my = R.UniformMesh(-5.0, 5.0, sz[0]), R.UniformMesh(-6.0, 4.5, sz[1])
    nbdt = 11
    dt = 2 * math.pi / nbdt
    adv = R.Advection((mx, my), [R.Lagrange(5), R.Lagrange(5)], dt, [([1, 2], 1, 1, True), ([2, 1], 1, 2, True)], tab_coef=R.magicsplit(dt))
    X, Y = np.meshgrid(mx.points, my.points, indexing="ij")
    f0 = np.asfortranarray(np.exp(-2 * (X**2 + (Y + 1.2) ** 2)))
    advd = R.AdvectionData(adv, f0, R.getrotationvar(adv))
    for _ in range(nbdt):
        while R.advection(advd):
            pass
    assert np.max(np.abs(advd.data - f0)) < 1e-3
    # Poisson: E from the plugin solves d/dx E = rho spectrally in 1-D
    nx, nv = 64, 32
    mx, mv = R.UniformMesh(0.0, 4 * math.pi, nx), R.UniformMesh(-6.0, 6.0, nv)
    adv = R.Advection((mx, mv), [R.Lagrange(5)] * 2, 0.1, [([1, 2], 1, 1, True), ([2, 1], 1, 2, True)])
    f = R.dotprod((1 + 0.1 * np.cos(0.5 * mx.points), np.exp(-mv.points**2 / 2) / math.sqrt(2 * math.pi)))
    pv = R.getpoissonvar(adv)
    advd = R.AdvectionData(adv, f, pv)
    pv.compute_charge(advd)
    pv.compute_elfield()
    assert abs(pv.rho.sum()) < 1e-12
    assert np.max(np.abs(pv.rho - 0.1 * np.cos(0.5 * mx.points))) < 1e-8  # int of the Gaussian ~ 1
    E = pv.t_elfield[0]
    # reference convention (src/poisson.jl:8,139-144): E_hat = (i k / |k|^2) rho_hat  ->  E = -(A/k) sin(kx)
    assert np.max(np.abs(E + 0.2 * np.sin(0.5 * mx.points) * (pv.rho.max() / 0.1))) < 1e-8
    assert abs(R.compute_ee(advd) - mx.step * np.sum(E**2)) < 1e-15


def test_inside_edge_is_rejected_by_advection_and_accepted_by_the_types():
    """InsideEdge exists at the kernel seam only: the reference's advection! has no method for it
    (src/advection.jl:627-631 passes one weight vector, src/interpolation.jl:250-256 wants one per offset)."""
    import slb200 as S

    it = S.Lagrange(7, edge=S.InsideEdge)
    assert it.edge == S.InsideEdge and S.Lagrange(7).edge == S.CircEdge
    with pytest.raises(ValueError):
        S.Lagrange(7, edge=3)
    m = S.UniformMesh(0.0, 1.0, 16)
    with pytest.raises(ValueError):
        S.Advection((m, m), [it, it], 0.1, [([1, 2], 1, 1, True), ([2, 1], 1, 2, True)])
    # the time-algorithm arguments of Advection (src/advection.jl:96-97)
    adv = S.Advection((m, m), [S.Lagrange(3)] * 2, 0.1, [([1, 2], 2, 1, False)], tab_coef=S.nosplit(0.1), timealg=S.ABTimeAlg_ip)
    assert adv.ordalg == 4 and len(adv.abcoef) == 5
    assert S.Advection((m, m), [S.Lagrange(3)] * 2, 0.1, [([1, 2], 2, 1, False)], tab_coef=S.nosplit(0.1)).ordalg == 0
    with pytest.raises(ValueError):
        S.Advection((m, m), [S.Lagrange(3)] * 2, 0.1, [([1, 2], 2, 1, False)], timealg=9)


def _program_barriers(ops):
    """ops: [(reads, writes)] with ranges (addr, len[, one_block]); returns the barrier flag of every op (the analysis
    slb_program_end runs, csrc/slb_program_host.h, compiled for the host)"""
    so = os.path.join(ROOT, "semilagrangian.jl_b200", "lib", "libslb200_hosttest.so")
    L = C.CDLL(so)
    ip, lp = C.POINTER(C.c_int), C.POINTER(C.c_longlong)
    L.slbt_program_barriers.argtypes = [C.c_int, ip, ip, lp, lp, ip, ip]
    nr = np.array([len(r) for r, _ in ops], dtype=np.int32)
    nw = np.array([len(w) for _, w in ops], dtype=np.int32)
    flat = [rg for r, w in ops for rg in list(r) + list(w)]
    addr = np.array([rg[0] for rg in flat], dtype=np.int64)
    ln = np.array([rg[1] for rg in flat], dtype=np.int64)
    b0 = np.array([int(len(rg) > 2 and rg[2]) for rg in flat], dtype=np.int32)
    out = np.zeros(len(ops), dtype=np.int32)
    assert L.slbt_program_barriers(len(ops), nr.ctypes.data_as(ip), nw.ctypes.data_as(ip), addr.ctypes.data_as(lp), ln.ctypes.data_as(lp),
                                   b0.ctypes.data_as(ip), out.ctypes.data_as(ip)) == 0
    return out.tolist()


def test_step_program_barrier_placement():
    """Grid barriers of a step program go exactly where an op touches what an earlier op of the same barrier interval
    wrote (or overwrites what it read); accesses made by one block only need none between themselves; the op list
    repeats, so the first op is checked against the last ones."""
    A, B, VT, P, RHO, E, EE = 0x1000000, 0x2000000, 0x3000000, 0x4000000, 0x5000000, 0x6000000, 0x7000000
    NF, NP, NE = 128 * 256 * 8, 4 * 128 * 8, 128 * 8
    sweep = lambda src, dst, tab=None: ([(src, NF)] + ([tab] if tab else []), [(dst, NF)])
    charge = lambda src: ([(src, NF)], [(P, NP)])
    field = ([(P, NP)], [(RHO, NE, True), (E, NE, True)])      # every block solves; block 0 stores rho and E
    ee = lambda slot: ([], [(EE + 8 * slot, 8, True)])          # reduced from a block-local copy of E
    # x - v - x Strang order, two steps (the velocity sweep reads E from the block-local copy: no range)
    two_steps = [sweep(A, B, (VT, 2048)), charge(B), field, sweep(B, A), sweep(A, B, (VT, 2048)), ee(0),
                 sweep(B, A, (VT, 2048)), charge(A), field, sweep(A, B), sweep(B, A, (VT, 2048)), ee(1)]
    #            x sweep | charge | field + v sweep | x sweep (+ ee) -- four barriers per step
    assert _program_barriers(two_steps) == [1, 1, 1, 0, 1, 0, 1, 1, 1, 0, 1, 0]
    # a velocity sweep that reads E from GLOBAL memory needs the barrier after the field solve
    glob = list(two_steps)
    glob[3] = sweep(B, A, (E, NE))
    assert _program_barriers(glob)[3] == 1
    # an energy reduction from global E by block 0 right after block 0's field solve: both one-block accesses, no barrier;
    # the same ranges without the one-block mark: barrier
    assert _program_barriers([charge(A), field, ([(E, NE, True)], [(EE, 8, True)])]) == [1, 1, 0]
    # (the barrier in front of the reduction also separates the field solve from the next repetition's charge pass)
    assert _program_barriers([charge(A), field, ([(E, NE)], [(EE, 8, True)])]) == [0, 1, 1]
    # write after read: the second op overwrites what the first one reads
    assert _program_barriers([([(A, 64)], [(B, 64)]), ([(VT, 64)], [(A, 64)])]) == [1, 1]
    # independent ops do not wait for each other; an op that writes conflicts with ITSELF in the next repetition
    # (conservative: one barrier per repetition)
    assert _program_barriers([([(A, 64)], [(B, 64)]), ([(VT, 64)], [(P, 64)])]) == [1, 0]
    assert _program_barriers([([(A, 64)], []), ([(VT, 64)], [])]) == [0, 0]
    # partial overlap counts; touching ranges do not
    assert _program_barriers([([], [(A, 64)]), ([(A + 63, 8)], [(B, 8)])]) == [1, 1]
    assert _program_barriers([([], [(A, 64)]), ([(A + 64, 8)], [(B, 8)])]) == [1, 0]


def test_segmented_bspline_correction_cut_table():
    """The segmented B-spline kernel drops, at compile time, the corrections z_k^(j+1) E of rows j >= slb_seg_cut(H, k, M)
    (csrc/slb_bspseg.cuh).  The poles of the order-(2H+1) symbol are universal constants: check the table against the
    poles computed here from the exact node values -- every dropped factor is below 1e-19, and the table is not wasteful
    (at most one pair of rows is kept beyond need)."""
    so = os.path.join(ROOT, "semilagrangian.jl_b200", "lib", "libslb200_hosttest.so")
    L = C.CDLL(so)
    L.slbt_seg_cut.argtypes = [C.c_int, C.c_int, C.c_int]
    for H in range(1, 6):
        order = 2 * H + 1
        vals = [float(x) for x in T.bspline_node_values_rat(order)]   # B(1) .. B(order) at the interior nodes (symmetric)
        sym = np.array(vals, dtype=np.float64)                         # coefficients of z^H a(z): a palindromic polynomial
        roots = np.roots(sym)
        poles = sorted(abs(r) for r in roots if abs(r) < 1.0)
        assert len(poles) == H and all(abs(r.imag) < 1e-9 for r in roots)
        for M in (8, 16, 32, 64):
            for k, z in enumerate(poles):
                cut = L.slbt_seg_cut(H, k, M)
                assert 0 < cut <= M and cut % 2 == 0
                if cut < M:
                    assert z ** (cut + 1) < 1e-19, (H, k, M, cut, z)          # first dropped row
                    assert cut <= 4 or z ** (cut - 3) >= 1e-19, (H, k, M, cut, z)   # at most one spare pair of rows
