"""Advection / AdvectionData / advection! -- host mirror of src/advection.jl.

The split-step state machine (StateAdv, state schedule, coefficient schedule, nextstate!)
is the reference's, unchanged in behaviour (src/advection.jl:10-20, :88-140, :152-163,
:358-367).  What changes is where the data lives: AdvectionData owns a device-resident grid
(slb_grid: f plus an equally sized scratch array, the counterpart of `data` + `bufdata`,
src/advection.jl:264-267) and `advection(advd)` (== advection!, :594-704) is a single
slb_sweep call -- no permutedims!, no per-line host loop, f never leaves HBM.

Displacement providers (AbstractExtDataAdv, src/advection.jl:172, docs/src/extdataadv.md):
`initcoef(advd)` is kept; the per-line callback `getalpha(parext, advd, indext)` cannot run
on the device, so a provider instead implements
    alpha_table(advd) -> (table, strides, scale, on_device)
meaning  alpha(line) = scale * table[sum_d idx_d * strides[d]]  over the line's other-dim
indices -- which is exactly how the reference's own providers index their `bufcur` arrays.
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from .interp import HERMITE, LAGRANGE
from .splitting import strangsplit
from . import unsplit2d
from .unsplit2d import NoTimeAlg


def modone(ind, n):
    """src/util.jl:73"""
    return (ind - 1) % n + 1


def invperm(p):
    q = [0] * len(p)
    for i, v in enumerate(p):
        q[v - 1] = i + 1
    return q


class StateAdv:
    """src/advection.jl:10-20 (perm is 1-based, as in the reference API)"""

    def __init__(self, ind, perm, ndims, stcoef, isconstdec):
        perm = [int(v) for v in perm]
        if sorted(perm) != list(range(1, len(perm) + 1)):
            raise ValueError(f"state {ind}: perm={perm} is not a permutation")
        self.ind, self.perm, self.invp = ind, perm, invperm(perm)
        self.ndims, self.stcoef, self.isconstdec = int(ndims), int(stcoef), bool(isconstdec)


class AbstractExtDataAdv:
    """Displacement provider interface (src/advection.jl:172, :221-224)."""

    def initcoef(self, advd):
        raise NotImplementedError

    def alpha_table(self, advd):
        raise NotImplementedError

    def alpha_table_nd(self, advd):
        """states with ndims > 1: one alpha_table descriptor per swept dim (getalpha returns a tuple
        of ndims shifts, src/advection.jl:221-224)"""
        raise NotImplementedError

    def initcoef_reads_data(self, advd):
        """True when initcoef of the CURRENT state reads advd's data (e.g. a charge density).
        A provider that returns False for a state lets advection() fuse that stage with the one
        before it into a single pass over HBM (slb_sweep_pair).  Default: assume it does."""
        return True


class Advection:
    """Advection(t_mesh, t_interp, dt_base, states; tab_coef=strangsplit(dt_base))
    -- src/advection.jl:73-141.  `states` = [(perm, ndims, stcoef, isconstdec), ...]."""

    def __init__(self, t_mesh, t_interp, dt_base, states, tab_coef=None, ctx=None, timealg=NoTimeAlg, ordalg=None):
        N = len(t_mesh)
        # timealg::TimeAlgorithm = NoTimeAlg, ordalg = timealg != NoTimeAlg ? 4 : 0, abcoef = ABcoef(ordalg + 1)
        # (src/advection.jl:96-97, :136)
        if timealg not in (unsplit2d.NoTimeAlg, unsplit2d.ABTimeAlg_ip, unsplit2d.ABTimeAlg_new, unsplit2d.ABTimeAlg_init):
            raise ValueError(f"unknown timealg {timealg}")
        self.timealg = timealg
        self.ordalg = (4 if timealg != NoTimeAlg else 0) if ordalg is None else int(ordalg)
        if timealg != NoTimeAlg and not (1 <= self.ordalg <= 7):
            raise ValueError(f"ordalg={self.ordalg} must be in 1..7 for the Adams-Bashforth time algorithms")
        self.abcoef = unsplit2d.abcoef(self.ordalg + 1)
        if len(t_interp) != N:
            raise ValueError(f"size of vector of Interpolation must be equal to N={N}")
        self.N = N
        self.sizeall = tuple(len(m) for m in t_mesh)
        self.t_mesh = tuple(t_mesh)
        self.t_interp = list(t_interp)
        if any(getattr(it, "edge", 1) != 1 for it in self.t_interp):
            # the reference's advection! has no method for InsideEdge interpolations (its per-line call passes
            # one weight vector, src/advection.jl:627-631); they exist at the kernel seam only
            raise ValueError("InsideEdge interpolations are available through interpolate() only, not through Advection")
        self.dt_base = float(dt_base)
        self.states = [StateAdv(i + 1, *s) for i, s in enumerate(states)]
        for s in self.states:
            if len(s.perm) != N:
                raise ValueError("state permutation length must equal the number of dimensions")
        self.tab_coef = [float(c) for c in (strangsplit(self.dt_base) if tab_coef is None else tab_coef)]
        self.maxcoef = max(s.stcoef for s in self.states)
        restcoef = len(self.tab_coef) % self.maxcoef
        nbstatesplus = sum(1 for s in self.states if s.stcoef == restcoef)
        self.nbstates = (len(self.tab_coef) // self.maxcoef) * len(self.states) + nbstatesplus
        self.ctx = ctx

    # src/advection.jl:150-163
    def getst(self, x):
        return self.states[modone(x, len(self.states)) - 1]

    def getstcoef(self, x):
        return ((x - 1) // len(self.states)) * self.maxcoef + self.getst(x).stcoef

    def getcur_t(self, x):
        return self.tab_coef[self.getstcoef(x) - 1]

    def getinterp(self, x):
        st = self.getst(x)
        return [self.t_interp[d - 1] for d in st.perm[: st.ndims]]


def sizeall(adv):
    return adv.sizeall


class AdvectionData:
    """AdvectionData(adv, data, parext; time_init=0) -- src/advection.jl:229-313.
    Copies `data` to the device (the reference copies it too, :264-267)."""

    def __init__(self, adv, data, parext, time_init=0.0, ctx=None, initdatas=None):
        data = np.asarray(data)
        if tuple(data.shape) != adv.sizeall:
            raise ValueError(f"size(data)={tuple(data.shape)} it must be {adv.sizeall}")
        self.adv = adv
        self.ctx = ctx or adv.ctx or _lib.default_context()
        self.state_gen = 1
        self.time_cur = float(time_init)
        self.parext = parext
        self.flags = 0
        # per-point displacement field (bufcur), its history (t_bufc) and the start-up data of
        # ABTimeAlg_init (initdatas): src/advection.jl:241-243
        self.bufcur = None
        self.t_bufc = []
        self.initdatas = initdatas
        h = C.c_void_p()
        _lib.check(_lib.lib().slb_grid_create(self.ctx.h, adv.N, _lib.i64(adv.sizeall), C.byref(h)))
        self.grid = h
        self._rhopart = None        # device buffer for the partial charge planes of a fused space pass (see poisson.py)
        self._rhopart_planes = 0    # planes currently held (0: none / stale)
        # off by default: measured slower (the x1x2 pass grows by more than the charge pass it replaces)
        self.use_rhopart = os.environ.get("SLB_RHOPART", "0") != "0"
        self._linesum = None        # device buffer for per-line output sums (see poisson.py)
        self._linesum_dim = None    # dim whose sweep produced the sums currently held, else None
        self.use_linesum = True
        # pair fusion: a stage may be held back one advection() call and then run together with the
        # next one in a single pass over HBM (slb_sweep_pair, bit-identical to two sweeps); anything
        # that looks at the data flushes it first
        self.fuse_pairs = os.environ.get("SLB_FUSE", "1") != "0"
        self._pending = None
        self.n_fused = 0
        self.upload(data)
        # mesh nodes stay on the device: shift tables for space sweeps (src/poisson.jl:191-203)
        self._points_dev = {}

    # -- data movement ---------------------------------------------------------------
    def flush(self):
        """run a stage that advection() held back for pair fusion (no-op otherwise)"""
        pend, self._pending = self._pending, None
        if pend is not None:
            _sweep_now(self, *pend)

    def upload(self, data):
        self._pending = None  # the data it would have swept is being replaced
        host = np.asfortranarray(data, dtype=np.float64)
        _lib.check(_lib.lib().slb_grid_upload(self.grid, host.ctypes.data_as(C.c_void_p)))
        self.ctx.sync()
        self._linesum_dim = None
        self._rhopart_planes = 0

    def getdata(self, out=None):
        """getdata(advd) (src/advection.jl:317): a host copy of the device-resident array."""
        self.flush()
        if out is None:
            out = np.empty(self.adv.sizeall, dtype=np.float64, order="F")
        assert out.flags.f_contiguous and out.dtype == np.float64
        _lib.check(_lib.lib().slb_grid_download(self.grid, out.ctypes.data_as(C.c_void_p)))
        return out

    def data_field(self):
        """the current f of a 2-D grid as a (non-owning) device field"""
        n1, n2 = self.adv.sizeall
        return unsplit2d.DeviceField.view(self.ctx, n1, n2, 1, C.c_void_p(_lib.lib().slb_grid_front(self.grid)))

    def interpolate_data(self):
        """data <- interpolate!(f, data, bufcur, t_interp) (src/advection.jl:607-619, :703): front ->
        back with the per-point kernel, then the grid's buffers swap roles"""
        L = _lib.lib()
        n1, n2 = self.adv.sizeall
        back = unsplit2d.DeviceField.view(self.ctx, n1, n2, 1, C.c_void_p(L.slb_grid_back(self.grid)))
        its = self.adv.t_interp
        h1, h2 = its[0].handle(self.ctx, n1), its[1].handle(self.ctx, n2)
        work = back.like() if any(unsplit2d._is_bspline(it) for it in its) else None
        try:
            _lib.check(L.slb_interp2d_points(self.ctx.h, h1, h2, n1, n2, 1, self.data_field().ptr, self.bufcur.ptr, back.ptr,
                                             work.ptr if work is not None else None, int(self.flags)))
        finally:
            if work is not None:
                work.free()
        _lib.check(L.slb_grid_swap(self.grid))
        self._linesum_dim = None
        self._rhopart_planes = 0

    def points_dev(self, dim0):
        p = self._points_dev.get(dim0)
        if p is None:
            p = self.ctx.to_device(self.adv.t_mesh[dim0].points)
            self._points_dev[dim0] = p
        return p

    # -- accessors, src/advection.jl:315-336 -------------------------------------------
    def getst(self):
        return self.adv.getst(self.state_gen)

    def getstcoef(self):
        return self.adv.getstcoef(self.state_gen)

    def getcur_t(self):
        return self.adv.getcur_t(self.state_gen)

    def getinterp(self):
        return self.adv.getinterp(self.state_gen)

    def _getcurrentindice(self):
        return self.getst().perm[0]

    def nextstate(self):
        """nextstate! -- src/advection.jl:358-367"""
        if self.state_gen < self.adv.nbstates:
            self.state_gen += 1
            return True
        self.state_gen = 1
        self.time_cur += self.adv.dt_base
        return False

    def close(self):
        self._pending = None
        if self.grid:
            _lib.lib().slb_grid_destroy(self.grid)
            self.grid = None
        for p in self._points_dev.values():
            self.ctx.free(p)
        self._points_dev = {}
        if self._linesum is not None:
            self.ctx.free(self._linesum)
            self._linesum = None
        if self._rhopart is not None:
            self.ctx.free(self._rhopart)
            self._rhopart = None
        for fld in [self.bufcur] + list(self.t_bufc):
            if fld is not None:
                fld.free()
        self.bufcur, self.t_bufc = None, []
        # the displacement provider is the caller's object (it may serve other AdvectionData): not closed here

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def getdata(advd):
    return advd.getdata()


def _table_args(table, on_device):
    if on_device:
        ptr, length = table
        return ptr, int(length), None
    arr = np.ascontiguousarray(table, dtype=np.float64)
    return arr.ctypes.data_as(C.c_void_p), arr.size, arr


def _linesum_begin(advd, dim0, interp, want_linesum):
    advd._linesum_dim = None  # any sweep invalidates sums held from an earlier one
    advd._rhopart_planes = 0
    ls_ok = want_linesum and advd.use_linesum and dim0 > 0 and interp.order + 1 <= 14 and interp.tabfct.shape[1] <= 14
    if ls_ok:
        if advd._linesum is None:
            advd._linesum = advd.ctx.malloc(int(np.prod(advd.adv.sizeall)) // min(advd.adv.sizeall[1:]) * 8)
        _lib.check(_lib.lib().slb_grid_set_linesum(advd.grid, advd._linesum))
    return ls_ok


def _sweep_now(advd, dim0, interp, table, strides, scale, on_device, flags=0, want_linesum=False):
    n = advd.adv.sizeall[dim0]
    h = interp.handle(advd.ctx, n)
    ls_ok = _linesum_begin(advd, dim0, interp, want_linesum)
    ptr, length, _keep = _table_args(table, on_device)
    try:
        _lib.check(
            _lib.lib().slb_sweep(advd.grid, int(dim0), h, ptr, int(length), _lib.i64(strides), float(scale), 1 if on_device else 0, int(flags))
        )
    finally:
        if ls_ok:
            _lib.check(_lib.lib().slb_grid_set_linesum(advd.grid, None))
    if ls_ok:
        advd._linesum_dim = dim0


def sweep(advd, dim0, interp, table, strides, scale, on_device, flags=0, want_linesum=False):
    """One slb_sweep on advd's grid (kernel seam).  table: (device pointer, length) when
    on_device else a float64 numpy array.  want_linesum: also store, per line, the sum of the
    line's outputs (consumed by the Poisson provider's next charge density)."""
    advd.flush()
    _sweep_now(advd, dim0, interp, table, strides, scale, on_device, flags, want_linesum)


FUSED_ORDERS = (3, 5, 7, 9, 11)  # order + 1 in {4, 6, 8, 10, 12}: instantiated in csrc/slb_pair.cu


def _pair_candidate(advd, stA, interpA, stridesA, on_device, stB):  # noqa: ARG001
    """Can stage A (about to run) be held back and fused with the next stage B?  Static part of
    the decision (slb_sweep_pair re-checks and may still answer SLB_E_UNSUPPORTED)."""
    adv = advd.adv
    if not advd.fuse_pairs or adv.N > 4:
        return False
    if stB.ndims != 1 or not stB.isconstdec:
        return False
    dA, dB = stA.perm[0] - 1, stB.perm[0] - 1
    if dB == 0 or dA == dB or stridesA[dB] != 0:
        return False
    interpB = adv.t_interp[dB]
    plain = lambda it: getattr(it, "kind", None) in (LAGRANGE, HERMITE) and it.tabfct.shape[1] <= 14
    if not (plain(interpA) and plain(interpB)) or interpA.order != interpB.order or interpA.order not in FUSED_ORDERS:
        return False
    return adv.sizeall[dA] >= interpA.order + 1


def sweep_pair(advd, stageA, stageB):
    """Two stages in one pass over HBM.  stage = (dim0, interp, table, strides, scale, on_device,
    flags, want_linesum).  Returns False when the library does not fuse this combination."""
    dA, itA, tabA, strA, scA, devA = stageA[:6]
    dB, itB, tabB, strB, scB, devB = stageB[:6]
    flags, want_ls = stageB[6], stageB[7]
    if devA != devB:
        return False
    hA = itA.handle(advd.ctx, advd.adv.sizeall[dA])
    hB = itB.handle(advd.ctx, advd.adv.sizeall[dB])
    pA, lA, _ka = _table_args(tabA, devA)
    pB, lB, _kb = _table_args(tabB, devB)
    ls_ok = _linesum_begin(advd, dB, itB, want_ls)
    L = _lib.lib()
    want_rho = advd.use_rhopart and bool(getattr(advd.parext, "wants_rhopart", lambda a, da, db: False)(advd, dA, dB))
    if want_rho:
        cap = int(np.prod(advd.adv.sizeall)) // 2
        if advd._rhopart is None:
            advd._rhopart = advd.ctx.malloc(cap * 8)
        _lib.check(L.slb_grid_set_rhopart(advd.grid, advd._rhopart, cap))
    try:
        rc = L.slb_sweep_pair(advd.grid, int(dA), hA, pA, lA, _lib.i64(strA), float(scA), int(dB), hB, pB, lB, _lib.i64(strB),
                              float(scB), 1 if devA else 0, int(flags))
        if want_rho and rc == 0:
            advd._rhopart_planes = int(L.slb_grid_rhopart_planes(advd.grid))
    finally:
        if want_rho:
            _lib.check(L.slb_grid_set_rhopart(advd.grid, None, 0))
        if ls_ok:
            _lib.check(_lib.lib().slb_grid_set_linesum(advd.grid, None))
    if rc == _lib.SLB_E_UNSUPPORTED:
        return False
    _lib.check(rc)
    if ls_ok:
        advd._linesum_dim = dB
    advd.n_fused += 1
    return True


def _advection_2d(advd):
    """A const-shift state with ndims >= 2 (e.g. ([1,2,3,4], 2, 1, true), test/test_poisson2d.jl:276,
    examples/vlasov-poisson-2d2v.jl:196): the reference shifts every 2-D slice by a constant pair
    (alpha_1, alpha_2) with the tensor stencil of src/interpolation.jl:212-231.  That operator is
    the product of the two 1-D stencils (and of the 1-D B-spline solves, :48-94), so it runs as ONE
    pair-fused pass (slb_sweep_pair), or as two sweeps where the pair is not fusable; results agree
    with the tensor form to rounding (the two forms add the same products in a different order)."""
    st = advd.getst()
    ext = advd.parext
    advd.flush()
    ext.initcoef(advd)
    descr = ext.alpha_table_nd(advd)  # one (table, strides, scale, on_device) per swept dim
    want = bool(getattr(ext, "wants_linesum", lambda a: False)(advd))
    stages = []
    for x in range(st.ndims):
        table, strides, scale, on_device = descr[x]
        d = st.perm[x] - 1
        stages.append((d, advd.adv.t_interp[d], table, list(strides), scale, on_device, advd.flags, False))
    if st.ndims == 2:
        if stages[1][0] == 0:  # the march of the fused pass cannot run along dim 0: the stencils commute
            stages.reverse()
    else:
        # ndims > 2 (the N-D tensor stencil of src/interpolation.jl:212-231 for any N): the 1-D stencils along
        # different dims commute, so they run in ascending dim order -- dim 0 first, where it can only be the
        # cross sweep of a fused pass -- and are paired greedily
        stages.sort(key=lambda sg: sg[0])
    stages[-1] = stages[-1][:7] + (want,)
    i = 0
    while i < len(stages):
        if i + 1 < len(stages) and advd.fuse_pairs and sweep_pair(advd, stages[i], stages[i + 1]):
            i += 2
        else:
            _sweep_now(advd, *stages[i])
            i += 1
    return advd.nextstate()


def advection(advd):
    """advection!(advd) -- src/advection.jl:594-704.  One split stage; returns True while
    more stages remain in the current time step.

    Pair fusion: when the NEXT stage can run in the same pass over HBM (slb_sweep_pair) and its
    initcoef does not look at the data, the current stage is only recorded; the next call runs
    both.  getdata / compute_ke / any charge density flush a recorded stage first, so every
    observable value is the one the reference's stage-by-stage execution produces."""
    st = advd.getst()
    if not st.isconstdec:
        return unsplit2d.advection_single_state(advd)
    if advd.adv.timealg != NoTimeAlg:
        raise NotImplementedError("the Adams-Bashforth time algorithms drive states with per-point shifts only")
    if st.ndims >= 2:
        return _advection_2d(advd)
    interp = advd.getinterp()[0]
    ext = advd.parext
    if advd._pending is not None and ext.initcoef_reads_data(advd):
        advd.flush()
    ext.initcoef(advd)  # src/advection.jl:407-408
    table, strides, scale, on_device = ext.alpha_table(advd)
    want = bool(getattr(ext, "wants_linesum", lambda a: False)(advd))
    cur = (st.perm[0] - 1, interp, table, list(strides), scale, on_device, advd.flags, want)
    if advd._pending is not None:
        pend, advd._pending = advd._pending, None
        if not sweep_pair(advd, pend, cur):
            _sweep_now(advd, *pend)
            _sweep_now(advd, *cur)
        return advd.nextstate()
    adv = advd.adv
    if advd.state_gen < adv.nbstates:  # never hold a stage back across the end of a time step
        nxt = adv.getst(advd.state_gen + 1)
        if _pair_candidate(advd, st, interp, strides, on_device, nxt):
            advd.state_gen += 1
            ok = not ext.initcoef_reads_data(advd)
            advd.state_gen -= 1
            if ok:
                advd._pending = cur
                return advd.nextstate()
    _sweep_now(advd, *cur)
    return advd.nextstate()


class StepGraph:
    """`nsteps` whole time steps of `advd` recorded ONCE (slb_capture_begin / _end) and replayed with one launch
    each (CUDA graph): for grids so small that a step is bound by launch latency, not by HBM -- the 1D1V example
    (128 x 256, examples/vlasov-poisson-1d1v.jl:60-64) issues 7 kernels for 256 KB of data.  The loop semantics stay
    the reference's: `launch()` advances `nsteps` steps, `energies()` returns the electric energy after each of them
    (compute_ee, src/util_poisson.jl:156-162, reduced on the device inside the graph).

    The recorded steps must leave the grid's front/back roles as they found them: use an even `nsteps` when a step
    has an odd number of passes.  Displacement providers must keep their tables on the device (the Poisson, rotation
    and translation providers do)."""

    def __init__(self, advd, nsteps=2):
        L = _lib.lib()
        adv, ctx, pv = advd.adv, advd.ctx, advd.parext
        if advd.state_gen != 1:
            raise ValueError("StepGraph records whole time steps: build it between two steps")
        advd.flush()
        self.advd, self.nsteps = advd, int(nsteps)
        # everything a step would allocate or upload lazily happens now, not inside the recording
        for d, it in enumerate(adv.t_interp):
            it.handle(ctx, adv.sizeall[d])
            advd.points_dev(d)
        if advd._linesum is None and adv.N > 1:
            advd._linesum = ctx.malloc(int(np.prod(adv.sizeall)) // min(adv.sizeall[1:]) * 8)
        self.E = list(getattr(pv, "E_dev", []))
        if hasattr(pv, "field_solve"):
            pv.field_solve(advd)  # sizes the scratch; the field of the current f (every step recomputes it anyway)
        self.nE = len(self.E)
        self.ee_dev = ctx.malloc(max(1, self.nsteps * max(1, self.nE)) * 8)
        dx = 1.0
        for m in adv.t_mesh[: getattr(pv, "Nsp", 0)]:
            dx *= m.step
        ctx.sync()
        front0, t0, nf0 = L.slb_grid_front(advd.grid), advd.time_cur, advd.n_fused
        _lib.check(L.slb_capture_begin(ctx.h))
        h = C.c_void_p()
        try:
            for s in range(self.nsteps):
                while advection(advd):
                    pass
                for d, e in enumerate(self.E):
                    _lib.check(L.slb_reduce_sumsq_async(ctx.h, e, pv.nsp_tot, dx, C.c_void_p(self.ee_dev.value + 8 * (s * self.nE + d))))
        finally:
            rc = L.slb_capture_end(ctx.h, C.byref(h))
        _lib.check(rc)
        self.h = h
        self.fused_per_launch = advd.n_fused - nf0
        advd.time_cur, advd.n_fused = t0, nf0  # nothing has run yet
        if L.slb_grid_front(advd.grid) != front0:
            L.slb_graph_destroy(self.h)
            self.h = None
            raise ValueError(f"{self.nsteps} step(s) swap the grid's buffers an odd number of times: record an even number of steps")

    def launch(self):
        """advance nsteps time steps (asynchronous)"""
        _lib.check(_lib.lib().slb_graph_launch(self.h))
        for _ in range(self.nsteps):  # the same additions as nextstate!, src/advection.jl:364
            self.advd.time_cur += self.advd.adv.dt_base
        self.advd.n_fused += self.fused_per_launch

    def energies(self):
        """electric energy after each of the last launch's steps (synchronises)"""
        if not self.nE:
            return []
        v = self.advd.ctx.to_host(self.ee_dev, self.nsteps * self.nE).reshape(self.nsteps, self.nE)
        return [float(sum(row)) for row in v]   # dx * sum_d sum(E_d .^ 2), components added left to right

    def close(self):
        if self.h:
            _lib.lib().slb_graph_destroy(self.h)
            self.h = None
            self.advd.ctx.free(self.ee_dev)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class StepProgram:
    """`nsteps` whole time steps of `advd` recorded ONCE as a step program (slb_program_begin / _end) and run
    `repeat` times over by ONE persistent cooperative kernel per `launch()` -- no launch per split stage at all, grid
    barriers only between dependent stages.  For grids that live in the L2 and are bound by launch latency: the 1D1V
    example (128 x 256, examples/vlasov-poisson-1d1v.jl:60-64) takes 7 kernels per step stepwise and as a CUDA graph
    (`StepGraph`).  Loop semantics as `StepGraph`: `launch()` advances nsteps * repeat steps, `energies()` returns the
    electric energy after each of them; data and energies are bit-identical to the step-by-step driver.

    Raises `SlbError` (SLB_E_UNSUPPORTED) when a step contains a call step programs do not record (B-splines, pair
    fusion, more than one space dim, host shift tables): nothing has run then and `StepGraph` is the fallback."""

    def __init__(self, advd, nsteps=2, repeat=1):
        L = _lib.lib()
        adv, ctx, pv = advd.adv, advd.ctx, advd.parext
        if advd.state_gen != 1:
            raise ValueError("StepProgram records whole time steps: build it between two steps")
        if int(nsteps) < 1 or int(repeat) < 1:
            raise ValueError("nsteps and repeat must be positive")
        advd.flush()
        self.advd, self.nsteps, self.repeat = advd, int(nsteps), int(repeat)
        for d, it in enumerate(adv.t_interp):
            it.handle(ctx, adv.sizeall[d])
            advd.points_dev(d)
        self.E = list(getattr(pv, "E_dev", []))
        self.nE = len(self.E)
        if hasattr(pv, "field_solve"):
            pv.field_solve(advd)  # the field of the current f, as the first recorded step will recompute it
        self.per_rep = self.nsteps * max(1, self.nE)
        self.ee_dev = ctx.malloc(self.per_rep * self.repeat * 8)
        dx = 1.0
        for m in adv.t_mesh[: getattr(pv, "Nsp", 0)]:
            dx *= m.step
        ctx.sync()
        if advd._linesum is None and adv.N > 1:
            advd._linesum = ctx.malloc(int(np.prod(adv.sizeall)) // min(adv.sizeall[1:]) * 8)
        t0, nf0, lsd0 = advd.time_cur, advd.n_fused, advd._linesum_dim
        self.h = None
        h = C.c_void_p()
        _lib.check(L.slb_program_begin(ctx.h))
        err = None
        try:
            for s in range(self.nsteps):
                while advection(advd):
                    pass
                for d, e in enumerate(self.E):
                    _lib.check(L.slb_reduce_sumsq_async(ctx.h, e, pv.nsp_tot, dx, C.c_void_p(self.ee_dev.value + 8 * (s * self.nE + d))))
            self._lsd_end = advd._linesum_dim  # line sums the last recorded sweep leaves behind (valid after every launch)
        except Exception as ex:   # SlbError (unsupported stage) or ValueError (bad argument): the recording is abandoned
            err = ex
        finally:
            rc = L.slb_program_end(ctx.h, C.byref(h))  # also puts the grids' front/back roles back: nothing has run
            advd.time_cur, advd.n_fused = t0, nf0
            advd.state_gen = 1
            advd._pending = None
            advd._linesum_dim = lsd0   # nothing ran: sums held from the last real sweep are still the current ones
        if err is not None or rc != 0:
            ctx.free(self.ee_dev)
            if err is not None:
                raise err
            _lib.check(rc)
        self.h = h
        no, nb, nblk = C.c_int(), C.c_int(), C.c_int()
        _lib.check(L.slb_program_info(self.h, C.byref(no), C.byref(nb), C.byref(nblk)))
        self.nops, self.nbarriers, self.nblocks = no.value, nb.value, nblk.value

    def launch(self):
        """advance nsteps * repeat time steps with one kernel launch (asynchronous)"""
        _lib.check(_lib.lib().slb_program_launch(self.h, self.repeat, self.per_rep))
        for _ in range(self.nsteps * self.repeat):  # the same additions as nextstate!, src/advection.jl:364
            self.advd.time_cur += self.advd.adv.dt_base
        self.advd._linesum_dim = self._lsd_end

    def profile(self):
        """one more launch with block 0's time stamps: [(kind, barrier-wait ns, run ns)] per recorded op (synchronises)"""
        n = self.nops
        kinds, wait, run = (C.c_int * n)(), (C.c_double * n)(), (C.c_double * n)()
        _lib.check(_lib.lib().slb_program_profile(self.h, self.repeat, self.per_rep, n, kinds, wait, run))
        for _ in range(self.nsteps * self.repeat):
            self.advd.time_cur += self.advd.adv.dt_base
        self.advd._linesum_dim = self._lsd_end
        names = {1: "sweep", 2: "charge", 3: "field", 4: "sumsq"}
        return [(names.get(kinds[k], "?"), wait[k], run[k]) for k in range(n)]

    def energies(self):
        """electric energy after each of the last launch's steps (synchronises)"""
        if not self.nE:
            return []
        n = self.nsteps * self.repeat
        v = self.advd.ctx.to_host(self.ee_dev, n * self.nE).reshape(n, self.nE)
        return [float(sum(row)) for row in v]

    def close(self):
        if self.h:
            _lib.lib().slb_program_destroy(self.h)
            self.h = None
            self.advd.ctx.free(self.ee_dev)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
