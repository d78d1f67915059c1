"""Halo-sharded 2D2V Vlasov-Poisson driver: one process per GPU, NO transposes, no MPI/NCCL in the data path.

Replaces the reference's MPI back-end (src/mpiinterface.jl:1-38; MPIOpt in src/advection.jl:116-123, :269-292,
:381-384): there every rank holds ALL of f and the whole array is broadcast P times after every sweep.  Here
f[x1,x2,v1,v2] is split ONCE, in slabs of c = n4/P points along v2, and stays that way for the whole run:

    rank r owns v2 in [r c, (r+1) c) and stores  [ low halo | slab | high halo ]  =  c + 2H planes along v2
    x1 x2 pass : purely local (slab view); it ALSO stores its outputs that lie within H planes of a slab boundary
                 into the neighbours' halo planes (NVLink peer stores inside the pass, slb_sweep_pair_halo
                 SLB_HALO_PASSIVE)
    v1 v2 pass : v1 is local; the v2 sweep marches once over the c + 2H rows and emits the slab's c rows
                 (SLB_HALO_MARCH).  The pass that ends a time step pushes its boundary rows as well (the next
                 step starts with another v1 v2 pass).
    rho        : every rank reduces its slab over (v1, v2) -> P partial charge densities of n1 n2 doubles,
                 all-gathered through the ranks' mailboxes (slb_comm_allgather, peer stores + flags) and summed in
                 rank order inside the one-kernel Poisson solve, which is replicated.  That all-gather is the
                 step's only synchronisation: it also orders the halo pushes against the passes that read them.

Why the all-gather suffices as the only synchronisation (passes of a step: V1 reads halos, X pushes, V2 reads halos
and pushes; buffers rotate B0 -> B1 -> B2 -> B0, pass k reads B[k % 3] and writes / pushes into B[(k + 1) % 3]):
  * read-after-write: the halos V1(s+1) reads were pushed by the neighbours' V2(s); the all-gather before V1(s+1)
    completes only when every rank's contribution has arrived, and a rank sends it after its own V2(s) (stream order).
    Same for V2(s) and the neighbours' X(s), with the mid-step all-gather.
  * write-after-read: X(s) pushes into the neighbours' B2, last read by their V2(s-1): finished before they contributed
    to the all-gather that precedes V1(s).  V2(s) pushes into the neighbours' B0, last read by their V1(s): finished
    before they contributed to the mid-step all-gather that precedes V2(s).  With two buffers the second case would race.
  * the mailbox slots alternate between two sets: a rank can be one all-gather ahead of a peer, never two.

H = order/2 + 1 + (ceil(max_shift) - 1): velocity shifts alpha = dt/dv E with |alpha| < max_shift cells.  A shift
outside the halo raises on the next compute_ee()/getdata_local().  Per exchange a rank sends 2 H planes instead
of (P-1)/P of its slab (the transposing driver, slb200/distributed.py), and the sends overlap the pass.
Lagrange/Hermite kinds with pair-fusable orders only: B-spline pre-solves couple whole lines and keep using the
transposing driver.

The host language's only job is to carry opaque handle bytes between the ranks once (`allgather_bytes`:
MPI.Allgather in Julia; torch.distributed.all_gather_object or anything else in Python).  `local_group` builds
P ranks inside one process (on one GPU or several): the tests use it to check the sharded arithmetic against the
single-grid driver on the driver's one-GPU box.
"""
import ctypes as C
import math

import numpy as np

from . import _lib
from .advection import FUSED_ORDERS
from .interp import HERMITE, LAGRANGE

HANDLE_BYTES = 128
NBUF = 3  # buffer rotation: a neighbour's pushes never land in a buffer its owner may still be reading


class HaloUnsupported(ValueError):
    """this configuration is not on the halo-sharded path (use slb200.distributed.ShardedAdvectionData)"""


def halo_width(order, max_shift):
    return order // 2 + 1 + max(0, int(math.ceil(max_shift)) - 1)


def estimate_max_shift(adv, rho, margin=1.5):
    """Largest velocity shift |dt/dv E| (in cells) the first steps will see, from the initial charge density
    (host-side numpy restatement of compute_elfield!, src/poisson.jl:139-144), times a safety margin: sizes the
    halo before any device buffer exists.  The passes still check every shift against the halo at run time."""
    from .poisson import _get_fctv_k_imag

    rho = np.asarray(rho, dtype=np.float64)
    rho = rho - rho.mean()
    spec = np.fft.fft2(rho)
    # time coefficients of the velocity stages only (Strang: dt/2 for v, dt for x)
    dtmax = max(abs(adv.getcur_t(k)) for k in range(1, adv.nbstates + 1) if adv.getst(k).perm[0] > 2)
    amax = 0.0
    for x, m in enumerate(_get_fctv_k_imag(adv)):
        e = np.real(np.fft.ifft2(1j * m * spec))
        amax = max(amax, dtmax / adv.t_mesh[2 + x].step * float(np.max(np.abs(e))))
    return max(1.0, margin * amax)


class HaloShardedAdvectionData:
    """Sharded counterpart of AdvectionData + PoissonVar for the 4 const-shift states of
    examples/vlasov-poisson-2d2v.jl (v1, v2, x1, x2).  `data_local`: this rank's slab
    f[:, :, :, rank*c : (rank+1)*c] (numpy, any order)."""

    def __init__(self, adv, data_local, rank, nranks, allgather_bytes=None, device=None, max_shift="auto", _defer_connect=False):
        if adv.N != 4:
            raise HaloUnsupported("the sharded driver covers 2D2V grids (N = 4)")
        for st in adv.states:
            if st.ndims != 1 or not st.isconstdec:
                raise HaloUnsupported("const-shift 1-D states only")
        dims = [adv.getst(k + 1).perm[0] - 1 for k in range(adv.nbstates)]
        if len(dims) % 2 or any((dims[i], dims[i + 1]) not in ((2, 3), (0, 1)) for i in range(0, len(dims), 2)):
            raise HaloUnsupported(f"state order {dims} is not a sequence of (v1, v2) / (x1, x2) pairs")
        ok = lambda it: getattr(it, "kind", None) in (LAGRANGE, HERMITE) and it.tabfct.shape[1] <= 14 and it.order in FUSED_ORDERS
        its = adv.t_interp
        if not (all(ok(it) for it in its) and its[0].order == its[1].order and its[2].order == its[3].order):
            raise HaloUnsupported("the halo path needs pair-fusable Lagrange/Hermite interpolations (B-splines: transposing driver)")
        self.adv, self.rank, self.P = adv, int(rank), int(nranks)
        n1, n2, n3, n4 = self.gshape = adv.sizeall
        if n4 % self.P:
            raise HaloUnsupported(f"n4={n4} must be divisible by the number of ranks {self.P}")
        self.c = n4 // self.P
        if allgather_bytes is None and not _defer_connect:
            if self.P != 1:
                raise ValueError("allgather_bytes is required for more than one rank")
            allgather_bytes = lambda b: [b]
        if isinstance(max_shift, str):
            if max_shift != "auto" or allgather_bytes is None:
                raise ValueError("max_shift must be a number of cells, or 'auto' (which needs allgather_bytes)")
            part = np.ascontiguousarray(np.asarray(data_local, dtype=np.float64).sum(axis=(2, 3))) * (adv.t_mesh[2].step * adv.t_mesh[3].step)
            rho = sum(np.frombuffer(b, dtype=np.float64).reshape(n1, n2) for b in allgather_bytes(part.tobytes()))
            max_shift = estimate_max_shift(adv, rho)
        self.max_shift = float(max_shift)
        self.H = halo_width(its[3].order, self.max_shift)
        if self.c < 2 * self.H:
            raise HaloUnsupported(f"slab of {self.c} planes is shorter than two halos of {self.H}")
        if tuple(data_local.shape) != (n1, n2, n3, self.c):
            raise ValueError(f"local slab shape {tuple(data_local.shape)} must be {(n1, n2, n3, self.c)}")
        L = _lib.lib()
        if device is None:
            device = self.rank % max(1, L.slb_device_count())
        self.ctx = _lib.Context(device)
        self.plane = n1 * n2 * n3
        self.nhalo = self.plane * (self.c + 2 * self.H)   # doubles per haloed buffer
        h = C.c_void_p()
        _lib.check(L.slb_comm_create(self.ctx.h, self.rank, self.P, n1 * n2, C.byref(h)))
        self.comm = h
        self._raw = [self.ctx.malloc(self.nhalo * 8) for _ in range(NBUF)]
        self.ptr = [p.value for p in self._raw]
        host = np.ascontiguousarray(np.asfortranarray(data_local, dtype=np.float64).reshape(-1, order="F"))
        _lib.check(L.slb_memcpy_h2d(self.ctx.h, C.c_void_p(self.ptr[0] + 8 * self.H * self.plane), host.ctypes.data_as(C.c_void_p), host.nbytes))
        self.ctx.sync()
        self.cur = 0
        self._grids = {}
        self.state_gen = 1
        self.time_cur = 0.0
        self._pending = None
        self.n_fused = 0
        self.linesum_valid = False
        self.use_linesum = True
        self.has_field = False
        # Poisson pieces (replicated solve)
        from .poisson import _get_fctv_k_imag

        self.fctv = [np.ascontiguousarray(a.reshape(-1, order="F")) for a in _get_fctv_k_imag(adv)]
        arr = (_lib.c_double_p * 2)(*[_lib.dptr(a) for a in self.fctv])
        h = C.c_void_p()
        _lib.check(L.slb_poisson_create(self.ctx.h, 2, _lib.i64((n1, n2)), arr, C.byref(h)))
        self.plan = h
        self.rho_part = self.ctx.malloc(n1 * n2 * 8)
        self.rho_dev = self.ctx.malloc(n1 * n2 * 8)
        self.E_dev = [self.ctx.malloc(n1 * n2 * 8) for _ in range(2)]
        self.points = [self.ctx.to_device(m.points) for m in adv.t_mesh]
        self.linesum = self.ctx.malloc(self.plane * 8)
        self.left = self.right = None
        # split pushes (SLB_HALO_SPLIT_PUSH=1, off by default): a pushing pass stores the boundary planes of its LOW side
        # into the lower neighbour itself; the HIGH side travels as a peer copy on a second stream while the main stream
        # does the field solve that follows.  Measured on 8 GPUs at 128^4 (thin slabs, NVLink-bound passes): the x pass
        # drops from 0.223 to 0.156 ms, but the copy does not hide behind the latency-bound field-solve chain (0.093 ->
        # 0.147 ms) and the next v pass waits for it (0.193 -> 0.217): 0.771 -> 0.827 ms per step
        # (profiles/r2_bench_8gpu_halo_splitpush_experiment.json).  Kept for larger grids / other link ratios.
        import os as _os

        self.split_push = _os.environ.get("SLB_HALO_SPLIT_PUSH", "0") not in ("0", "")
        self.side = _lib.Context(device)
        self.ev_pass = self.ctx.event()
        self.ev_copy = self.side.event()
        self.side.record(self.ev_copy)
        self._await_copy = False
        # every allocation happens here, before the first cross-rank kernel: cudaMalloc / cudaFree synchronise the
        # device, which must not happen while another in-process rank's kernel waits for this rank's flag
        for d in range(4):
            its[d].handle(self.ctx, self.gshape[d])
        _lib.check(L.slb_charge_density_raw(self._grid(False, 0, 1), 2, 1.0, self.rho_part))
        self.ctx.sync()
        if not _defer_connect:
            self._connect(allgather_bytes(self._export()))

    # ---- bootstrap: opaque handle bytes, carried once by the host language ----------------------------------
    def _export(self):
        L = _lib.lib()
        out = b""
        hb = C.create_string_buffer(HANDLE_BYTES)
        _lib.check(L.slb_comm_export(self.comm, hb))
        out += hb.raw
        for p in self._raw:
            _lib.check(L.slb_comm_export_buffer(self.comm, p, hb))
            out += hb.raw
        return out

    def _connect(self, blobs, barrier=True):
        L = _lib.lib()
        assert len(blobs) == self.P and all(len(b) == HANDLE_BYTES * (1 + NBUF) for b in blobs)
        mail = b"".join(b[:HANDLE_BYTES] for b in blobs)
        _lib.check(L.slb_comm_connect(self.comm, C.create_string_buffer(mail, len(mail))))

        def open_bufs(r):
            if r == self.rank:
                return list(self.ptr)
            out = []
            for i in range(NBUF):
                q = C.c_void_p()
                hb = blobs[r][HANDLE_BYTES * (1 + i):HANDLE_BYTES * (2 + i)]
                _lib.check(L.slb_comm_open_buffer(self.comm, C.create_string_buffer(hb, HANDLE_BYTES), C.byref(q)))
                out.append(q.value)
            return out

        lo, hi = (self.rank - 1) % self.P, (self.rank + 1) % self.P
        self.left = open_bufs(lo)
        self.right = self.left if hi == lo else open_bufs(hi)
        # halos of the initial data: the boundary planes of the slab go to the neighbours' halo planes
        self._push_initial_halos()
        if barrier:
            _lib.check(L.slb_comm_barrier(self.comm))

    def _push_initial_halos(self):
        L = _lib.lib()
        nb = 8 * self.H * self.plane
        base = self.ptr[self.cur]
        # my rows [H, 2H) -> left's rows [H + c, 2H + c);  my rows [c, c + H) -> right's rows [0, H)
        _lib.check(L.slb_memcpy_d2d(self.ctx.h, C.c_void_p(self.left[self.cur] + 8 * self.plane * (self.H + self.c)),
                                    C.c_void_p(base + 8 * self.plane * self.H), nb))
        _lib.check(L.slb_memcpy_d2d(self.ctx.h, C.c_void_p(self.right[self.cur]), C.c_void_p(base + 8 * self.plane * self.c), nb))

    # ---- grid handles over the buffers ----------------------------------------------------------------------
    def _grid(self, haloed, cur, out):
        key = (haloed, cur, out)
        g = self._grids.get(key)
        if g is None:
            n1, n2, n3, _ = self.gshape
            off = 0 if haloed else 8 * self.H * self.plane
            ext = (n1, n2, n3, self.c + 2 * self.H if haloed else self.c)
            g = C.c_void_p()
            _lib.check(_lib.lib().slb_grid_create_external(self.ctx.h, 4, _lib.i64(ext), C.c_void_p(self.ptr[cur] + off),
                                                           C.c_void_p(self.ptr[out] + off), C.byref(g)))
            self._grids[key] = g
        return g

    # ---- state machine (src/advection.jl:152-158, :358-367) ---------------------------------------------------
    def getst(self):
        return self.adv.getst(self.state_gen)

    def getcur_t(self):
        return self.adv.getcur_t(self.state_gen)

    def nextstate(self):
        if self.state_gen < self.adv.nbstates:
            self.state_gen += 1
            return True
        self.state_gen = 1
        self.time_cur += self.adv.dt_base
        return False

    def _dim_after(self, k):
        adv = self.adv
        return adv.getst(k + 1 if k < adv.nbstates else 1).perm[0] - 1

    # ---- field solve (src/poisson.jl:119-144) -----------------------------------------------------------------
    def compute_field(self):
        adv = self.adv
        n1, n2, n3, n4 = self.gshape
        dv = adv.t_mesh[2].step * adv.t_mesh[3].step
        L = _lib.lib()
        if self.linesum_valid:
            # the v1 v2 pass that ended the last step left, per (x1, x2, v1), the sum of its slab outputs over v2
            _lib.check(L.slb_charge_density_from(self.ctx.h, self.linesum, n1 * n2, n3, dv, self.rho_part, 0))
        else:
            nxt = (self.cur + 1) % NBUF
            _lib.check(L.slb_charge_density_raw(self._grid(False, self.cur, nxt), 2, dv, self.rho_part))
        slots = C.c_void_p()
        _lib.check(L.slb_comm_allgather(self.comm, self.rho_part, n1 * n2, C.byref(slots)))
        arr = (C.c_void_p * 2)(*[p.value for p in self.E_dev])
        _lib.check(L.slb_poisson_solve_partial(self.plan, slots, self.P, 1.0, 1, self.rho_dev, arr))
        self.has_field = True

    def compute_ee(self):
        """src/util_poisson.jl:156-162 (replicated: every rank returns the same value)"""
        adv = self.adv
        dx = adv.t_mesh[0].step * adv.t_mesh[1].step
        n = self.gshape[0] * self.gshape[1]
        tot = 0.0
        for e in self.E_dev:
            v = C.c_double()
            _lib.check(_lib.lib().slb_reduce_sumsq(self.ctx.h, e, n, C.byref(v)))
            tot += v.value
        self.check()
        return dx * tot

    def check(self):
        """raise when a velocity shift left the halo since the last check (synchronises)"""
        fl = C.c_int()
        _lib.check(_lib.lib().slb_halo_error(self.ctx.h, C.byref(fl)))
        if fl.value & 1:
            raise _lib.SlbError(f"a velocity shift exceeded the halo of {self.H} planes (max_shift too small): results of this step are invalid")

    # ---- one advection! call ------------------------------------------------------------------------------------
    def advection(self):
        """advection!(advd) for the sharded grid: the first stage of a (v1, v2) / (x1, x2) pair is only recorded
        (a v1 stage first solves for the field, src/poisson.jl:171-176), the second runs both sweeps in one pass."""
        d = self.getst().perm[0] - 1
        dt = self.getcur_t()
        if self._pending is None:
            if d == 2:
                self.compute_field()
            self._pending = (d, dt)
            return self.nextstate()
        dA, dtA = self._pending
        self._pending = None
        assert d == dA + 1
        self._pass(dA, dtA, d, dt, self._dim_after(self.state_gen))
        return self.nextstate()

    def _pass(self, dA, dtA, dB, dtB, nxt):
        adv = self.adv
        L = _lib.lib()
        n1, n2, n3, n4 = self.gshape
        out = (self.cur + 1) % NBUF
        vpass = dA == 2
        push = nxt == 2  # the next pass sweeps v2: it reads the halos of what this pass writes
        if vpass:
            if not self.has_field:
                raise RuntimeError("velocity state before any field solve")
            tA, lA, sA, scA = self.E_dev[0], n1 * n2, [1, n1, 0, 0], dtA / adv.t_mesh[2].step
            tB, lB, sB, scB = self.E_dev[1], n1 * n2, [1, n1, 0, 0], dtB / adv.t_mesh[3].step
        else:
            tA, lA, sA, scA = self.points[2], n3, [0, 0, 1, 0], -dtA / adv.t_mesh[0].step
            tB, lB, sB, scB = C.c_void_p(self.points[3].value + 8 * self.rank * self.c), self.c, [0, 0, 0, 1], -dtB / adv.t_mesh[1].step
        g = self._grid(vpass, self.cur, out)
        off = 0 if vpass else 8 * self.H * self.plane
        split = push and self.split_push
        hl = _lib.SlbHalo()
        hl.mode = _lib.SLB_HALO_MARCH if vpass else _lib.SLB_HALO_PASSIVE
        hl.halo = self.H
        hl.shard_dim = 3
        hl.push_lo = (self.left[out] + off) if push else None
        hl.push_hi = (self.right[out] + off) if (push and not split) else None
        hl.err_flag = None
        if vpass:
            # the LOW halo rows of this pass's input were copied by the lower neighbour's second stream: hold the pass
            # until its signal has arrived (every rank pushed, so every rank waits: signals and waits match one to one)
            self._consume_copy()
        if split:
            self.ctx.wait_event(self.ev_copy)  # the previous copy out of this rotation's buffers is long over; make it formal
        want_ls = vpass and self.use_linesum and nxt == 2
        self.linesum_valid = False
        if want_ls:
            _lib.check(L.slb_grid_set_linesum(g, self.linesum))
        try:
            hA = adv.t_interp[dA].handle(self.ctx, self.gshape[dA])
            hB = adv.t_interp[dB].handle(self.ctx, self.gshape[dB])
            _lib.check(L.slb_sweep_pair_halo(g, dA, hA, tA, lA, _lib.i64(sA), float(scA), dB, hB, tB, lB, _lib.i64(sB), float(scB), 1, 0,
                                             C.byref(hl)))
            _lib.check(L.slb_grid_swap(g))  # keep the handle's orientation; the driver tracks `cur`
            if split:
                # rows [c, c + H) of the output (haloed row numbers) -> the upper neighbour's low halo rows [0, H),
                # as a peer copy on the second stream, then its signal; overlaps the field solve on the main stream
                self.ctx.record(self.ev_pass)
                self.side.wait_event(self.ev_pass)
                _lib.check(L.slb_memcpy_d2d(self.side.h, C.c_void_p(self.right[out]), C.c_void_p(self.ptr[out] + 8 * self.plane * self.c),
                                            8 * self.H * self.plane))
                _lib.check(L.slb_comm_signal(self.comm, self.side.h, (self.rank + 1) % self.P, 0))
                self.side.record(self.ev_copy)
                self._await_copy = True
        finally:
            if want_ls:
                _lib.check(L.slb_grid_set_linesum(g, None))
        self.linesum_valid = want_ls
        self.cur = out
        self.n_fused += 1

    def _consume_copy(self):
        """enqueue the wait for the lower neighbour's halo copy, if one is outstanding"""
        if self._await_copy:
            _lib.check(_lib.lib().slb_comm_wait(self.comm, None, (self.rank - 1) % self.P, 0))
            self._await_copy = False

    # ---- data access --------------------------------------------------------------------------------------------
    def getdata_local(self):
        """this rank's slab f[:, :, :, rank*c:(rank+1)*c] (numpy, Fortran order); valid between steps"""
        n1, n2, n3, _ = self.gshape
        flat = self.ctx.to_host(C.c_void_p(self.ptr[self.cur] + 8 * self.H * self.plane), self.plane * self.c)
        self.check()
        return flat.reshape((n1, n2, n3, self.c), order="F")

    def upload_local(self, host_flat):
        """replace this rank's slab (flat Fortran order, e.g. pinned memory) and re-send its halo planes; every
        rank must call it, followed by sync_ranks()"""
        self._pending = None
        self._consume_copy()
        _lib.check(_lib.lib().slb_memcpy_h2d(self.ctx.h, C.c_void_p(self.ptr[self.cur] + 8 * self.H * self.plane),
                                             host_flat.ctypes.data_as(C.c_void_p), self.plane * self.c * 8))
        self._push_initial_halos()
        self.linesum_valid = False

    def download_local(self, host_flat):
        _lib.check(_lib.lib().slb_memcpy_d2h(self.ctx.h, host_flat.ctypes.data_as(C.c_void_p),
                                             C.c_void_p(self.ptr[self.cur] + 8 * self.H * self.plane), self.plane * self.c * 8))
        self.ctx.sync()

    # ---- streamed host I/O: uploads and read-backs on their own streams (bench.py's end-to-end leg) ---------------
    def stream_io_begin(self):
        """contexts (streams) and events for read-backs and uploads that run next to the passes"""
        if getattr(self, "_io", None) is None:
            up, down = _lib.Context(self.ctx.device), _lib.Context(self.ctx.device)
            self._io = {"up": up, "down": down, "ev_step": self.ctx.event(), "ev_up": up.event(), "ev_down": down.event()}
            up.record(self._io["ev_up"])
            down.record(self._io["ev_down"])
        return self._io

    def stream_io_exchange(self, host_out_flat, host_in_flat):
        """after a step: read this rank's slab back into host_out_flat and, at the same time, upload host_in_flat as
        the next step's slab into the free buffer of the rotation (PCIe is full duplex); the next step's passes wait
        for both.  Every rank calls it; the new slab's halo planes go to the neighbours and the ranks synchronise."""
        io, L = self.stream_io_begin(), _lib.lib()
        nb, off = self.plane * self.c * 8, 8 * self.H * self.plane
        self._consume_copy()
        self.ctx.record(io["ev_step"])
        nxt = (self.cur + 1) % NBUF     # not read by anybody any more: pushes into it ended with the last all-gather
        io["down"].wait_event(io["ev_step"])
        _lib.check(L.slb_memcpy_d2h(io["down"].h, host_out_flat.ctypes.data_as(C.c_void_p), C.c_void_p(self.ptr[self.cur] + off), nb))
        io["down"].record(io["ev_down"])
        io["up"].wait_event(io["ev_step"])
        _lib.check(L.slb_memcpy_h2d(io["up"].h, C.c_void_p(self.ptr[nxt] + off), host_in_flat.ctypes.data_as(C.c_void_p), nb))
        io["up"].record(io["ev_up"])
        self.ctx.wait_event(io["ev_up"])
        self.ctx.wait_event(io["ev_down"])
        self.cur = nxt
        self._pending = None
        self.linesum_valid = False
        self._push_initial_halos()
        self.sync_ranks()

    def sync_ranks(self):
        _lib.check(_lib.lib().slb_comm_barrier(self.comm))

    @property
    def rho(self):
        return self.ctx.to_host(self.rho_dev, self.gshape[0] * self.gshape[1]).reshape(self.gshape[:2], order="F")

    @property
    def t_elfield(self):
        n = self.gshape[0] * self.gshape[1]
        return tuple(self.ctx.to_host(p, n).reshape(self.gshape[:2], order="F") for p in self.E_dev)

    def close(self):
        L = _lib.lib()
        if self.ctx is None:
            return
        self.side.sync()
        self.ctx.sync()
        for g in self._grids.values():
            L.slb_grid_destroy(g)
        self._grids = {}
        if self.plan:
            L.slb_poisson_destroy(self.plan)
            self.plan = None
        if self.comm:
            L.slb_comm_destroy(self.comm)   # unmaps the neighbours' buffers and mailboxes
            self.comm = None
        for p in self._raw + [self.rho_part, self.rho_dev, self.linesum] + self.E_dev + self.points:
            self.ctx.free(p)
        self._raw = []
        self.side.close()
        self.ctx.close()
        self.ctx = None


# ---------------------------------------------------------------------------------------------------------------
# pure layout conventions (numpy) -- pinned by world_size-2 gloo tests on CPU (tests/test_sharded_cpu.py) and used
# by the GPU tests to check what the kernels pushed
# ---------------------------------------------------------------------------------------------------------------
def slab_with_halos(global_arr, rank, nranks, H):
    """rank's [low halo | slab | high halo] view of a periodic global array, sharded along its LAST dim"""
    n = global_arr.shape[-1]
    c = n // nranks
    idx = [(rank * c - H + j) % n for j in range(c + 2 * H)]
    return np.asfortranarray(global_arr[..., idx])


def halo_destinations(rank, nranks, c, H):
    """where a rank's boundary planes go: [(source rows, destination rank, destination rows)] in haloed row numbers.
    Rows [H, 2H) (the slab's first H planes) fill the lower neighbour's HIGH halo, rows [c, c + H) (its last H
    planes) the upper neighbour's LOW halo -- what slb_sweep_pair_halo's push_lo / push_hi stores implement."""
    lo, hi = (rank - 1) % nranks, (rank + 1) % nranks
    return [((H, 2 * H), lo, (H + c, 2 * H + c)), ((c, c + H), hi, (0, H))]


def exchange_halos_reference(local_haloed, rank, nranks, H, dist, torch, group=None):
    """host-side (CPU tensors, any backend) restatement of the halo exchange: fills the halo planes of
    `local_haloed` ([..., c + 2H], slab already in place) from the neighbours' boundary planes"""
    c = local_haloed.shape[-1] - 2 * H
    mine = [np.ascontiguousarray(local_haloed[..., a:b]) for (a, b), _, _ in halo_destinations(rank, nranks, c, H)]
    payload = torch.from_numpy(np.stack(mine))
    allp = [torch.empty_like(payload) for _ in range(nranks)]
    dist.all_gather(allp, payload, group=group)
    for src in range(nranks):
        for k, (_, dst, (a, b)) in enumerate(halo_destinations(src, nranks, c, H)):
            if dst == rank:
                local_haloed[..., a:b] = allp[src][k].numpy()
    return local_haloed


def torch_allgather_bytes(dist, group=None):
    """`allgather_bytes` over an initialised torch.distributed group (any backend: the handles are a few hundred
    bytes, exchanged once).  A Julia host passes MPI.Allgather instead."""
    def fn(blob):
        out = [None] * dist.get_world_size(group)
        dist.all_gather_object(out, blob, group=group)
        return out
    return fn


def local_group(adv, data, nranks, devices=None, max_shift="auto", host_sync=False):
    """P ranks inside ONE process (tests; also a single host thread driving several GPUs): returns the list of
    rank objects, wired to each other directly.  `data`: the full array [n1, n2, n3, n4]."""
    n4 = adv.sizeall[3]
    c = n4 // nranks
    if isinstance(max_shift, str):
        max_shift = estimate_max_shift(adv, np.asarray(data).sum(axis=(2, 3)) * (adv.t_mesh[2].step * adv.t_mesh[3].step))
    ranks = [HaloShardedAdvectionData(adv, np.asfortranarray(data[:, :, :, r * c:(r + 1) * c]), r, nranks,
                                      device=None if devices is None else devices[r], max_shift=max_shift, _defer_connect=True)
             for r in range(nranks)]
    blobs = [s._export() for s in ranks]
    for s in ranks:
        s.ctx.sync()
    # two phases: every rank's initial halo pushes are enqueued before anybody's barrier kernel starts to spin
    for s in ranks:
        s._connect(blobs, barrier=False)
    for s in ranks:
        if host_sync:   # under a profiler that serialises kernel launches, ranks cannot wait for each other on the device
            s.ctx.sync()
        else:
            s.sync_ranks()
    return ranks
