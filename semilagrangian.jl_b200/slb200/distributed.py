"""Sharded 2D2V driver: one process per GPU, torch.distributed (NCCL over NVLink) plumbing.

Replaces the reference's MPI back-end (src/mpiinterface.jl:1-38, src/advection.jl:116-123,
:269-292, :381-384), whose strategy is replicated data + work split + an all-gather by P
broadcasts after EVERY sweep.  Here the 4-D grid f[x1,x2,v1,v2] is domain-decomposed in slabs
and re-sharded with one all-to-all only when the sharded dimension must be swept:

    layout B (shard x2): local [n1, n2/P, n3, n4]  -> sweeps along v1, v2; rho is a local reduction
    layout A (shard v2): local [n1, n2, n3, n4/P]  -> sweeps along x1, x2

A Strang step  v1 v2 | x1 x2 | v1 v2  needs two exchanges.  Neither needs a pack/unpack pass:
  B -> A : a layout-B slab is already contiguous per destination (v2 is the slowest dim); the
           received blocks form an array that is block-major along x2, which the x1 sweep reads
           directly (slb_sweep_ex, SLB_RESHARD_IN_BLOCKED).
  A -> B : the x2 sweep writes its output block-major along x2 (SLB_RESHARD_OUT_BLOCKED), so
           each destination's block is contiguous; the received blocks concatenate along v2
           into the standard layout-B slab.
rho slabs are all-gathered (n1*n2/P doubles per rank) and the Poisson solve is replicated.  With
exchange = "p2p" nothing in the data path goes through torch.distributed / NCCL: the re-shards are peer stores
inside the passes, the barriers and the rho all-gather are the library's mailbox collectives (slb_comm_*).
Lagrange / Hermite runs do not need this driver any more: slb200/sharded.py keeps ONE slab layout and exchanges
halo planes only.  This one remains for the B-spline kinds, whose pre-solve couples whole lines.

The pure index bookkeeping (block-major packing, slab ranges, split sizes) is kept free of CUDA
so that world_size-2 gloo tests on CPU cover it (tests/test_distributed_cpu.py).
"""
import ctypes as C

import numpy as np

from . import _lib
from .advection import modone  # noqa: F401  (re-exported for symmetry with advection.py)

LAYOUT_A, LAYOUT_B = "A", "B"  # A: shard dim 3 (v2); B: shard dim 1 (x2)
SHARD_DIM = {LAYOUT_A: 3, LAYOUT_B: 1}


# ---------------------------------------------------------------------------------------------
# pure layout algebra (numpy) -- shared by the CPU tests and the GPU driver's checks
# ---------------------------------------------------------------------------------------------
def splititr(nb, lgtot):
    """src/util.jl:26-32: split 1..lgtot into nb contiguous ranges (1-based, inclusive)."""
    lg, r = divmod(lgtot, nb)
    out = [(x * (lg + 1) + 1, (x + 1) * (lg + 1)) for x in range(r)]
    out += [(x * lg + r + 1, (x + 1) * lg + r) for x in range(r, nb)]
    return out


def splitvec(nb, v):
    """src/util.jl:34-36"""
    return [v[a - 1:b] for a, b in splititr(nb, len(v))]


def slab(n, nranks, rank):
    """[lo, hi) of rank's slab along a dim of extent n (equal slabs)."""
    if n % nranks != 0:
        raise ValueError(f"extent {n} is not divisible by {nranks} ranks")
    c = n // nranks
    return rank * c, (rank + 1) * c


def local_shape(global_shape, layout, nranks):
    s = list(global_shape)
    d = SHARD_DIM[layout]
    if s[d] % nranks != 0:
        raise ValueError(f"extent {s[d]} of dim {d} is not divisible by {nranks} ranks")
    s[d] //= nranks
    return tuple(s)


def to_block_major(arr, bdim, nblocks):
    """Flat array holding `arr` (Fortran order) block-major along bdim: what a sweep with
    SLB_RESHARD_OUT_BLOCKED writes / what SLB_RESHARD_IN_BLOCKED reads."""
    c = arr.shape[bdim] // nblocks
    parts = []
    for s in range(nblocks):
        idx = [slice(None)] * arr.ndim
        idx[bdim] = slice(s * c, (s + 1) * c)
        parts.append(np.asarray(arr[tuple(idx)]).reshape(-1, order="F"))
    return np.concatenate(parts)


def from_block_major(flat, shape, bdim, nblocks):
    """inverse of to_block_major"""
    shape = tuple(shape)
    c = shape[bdim] // nblocks
    bshape = list(shape)
    bshape[bdim] = c
    bl = int(np.prod(bshape))
    out = np.empty(shape, dtype=flat.dtype, order="F")
    for s in range(nblocks):
        idx = [slice(None)] * len(shape)
        idx[bdim] = slice(s * c, (s + 1) * c)
        out[tuple(idx)] = flat[s * bl:(s + 1) * bl].reshape(bshape, order="F")
    return out


def exchange(dist, out_flat, in_flat, group=None):
    """The re-shard collective: equal-split all-to-all of flat buffers (torch tensors on any
    backend: nccl on the GPUs, gloo in the CPU tests)."""
    dist.all_to_all_single(out_flat, in_flat, group=group)


def reshard_B_to_A_reference(local_B, nranks, dist, torch, group=None):
    """Host-side (CPU tensor) restatement of the B -> A exchange: returns the standard
    layout-A slab.  Used by the gloo tests to pin the block conventions the kernels rely on."""
    n1, c2, n3, n4 = local_B.shape
    send = torch.from_numpy(np.ascontiguousarray(local_B.reshape(-1, order="F")))
    recv = torch.empty_like(send)
    exchange(dist, recv, send, group)
    # received: nranks blocks, block s = [n1, c2, n3, c4] of rank s -> block-major along x2
    return from_block_major(recv.numpy(), (n1, c2 * nranks, n3, n4 // nranks), 1, nranks)


def reshard_A_to_B_reference(local_A, nranks, dist, torch, group=None):
    """Host-side restatement of the A -> B exchange: returns the standard layout-B slab."""
    n1, n2, n3, c4 = local_A.shape
    send = torch.from_numpy(to_block_major(local_A, 1, nranks))
    recv = torch.empty_like(send)
    exchange(dist, recv, send, group)
    # received blocks r = [n1, c2, n3, c4(r)] concatenate along the slowest dim
    return recv.numpy().reshape((n1, n2 // nranks, n3, c4 * nranks), order="F")


# ---------------------------------------------------------------------------------------------
# GPU driver
# ---------------------------------------------------------------------------------------------
class ShardedAdvectionData:
    """Sharded counterpart of AdvectionData + PoissonVar for 2D2V Vlasov-Poisson
    (4 states of examples/vlasov-poisson-2d2v.jl: v1, v2, x1, x2 with ndims = 1).

    `data_local_B`: this rank's layout-B slab, shape [n1, n2/P, n3, n4] (numpy, Fortran order).

    exchange = "p2p"  : the sweep that precedes a layout change stores its output straight into
                        the destination ranks' buffers over NVLink (slb_sweep_peer, buffers
                        mapped with CUDA IPC); the only collective left is a tiny barrier.
    exchange = "nccl" : the sweep writes a block-major local array and an NCCL all-to-all
                        (torch.distributed.all_to_all_single) moves the blocks.
    """

    def __init__(self, adv, data_local_B, rank=None, world=None, group=None, device=None, exchange="p2p"):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.group = torch, dist, group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.P = dist.get_world_size(group) if world is None else world
        if adv.N != 4:
            raise ValueError("the sharded driver covers 2D2V grids (N = 4)")
        for st in adv.states:
            if st.ndims != 1 or not st.isconstdec:
                raise NotImplementedError("const-shift 1-D states only")
        if exchange not in ("p2p", "nccl"):
            raise ValueError("exchange must be 'p2p' or 'nccl'")
        self.exchange = exchange if self.P > 1 else "nccl"
        self.adv = adv
        self.gshape = adv.sizeall
        n1, n2, n3, n4 = self.gshape
        if n2 % self.P or n4 % self.P:
            raise ValueError(f"n2={n2} and n4={n4} must be divisible by the number of ranks {self.P}")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        # one dedicated stream carries the sweeps AND the collectives: torch.distributed orders its
        # NCCL work against the current stream, so every call below runs under this stream
        self.stream = torch.cuda.Stream(self.device)
        self.ctx = _lib.Context(self.device.index, stream=self.stream.cuda_stream)
        self.nloc = n1 * n2 * n3 * n4 // self.P
        L = _lib.lib()
        if self.exchange == "p2p":
            # three buffers: a peer never stores into a buffer its owner may still be reading
            self.nbuf = 3
            self._raw = [self.ctx.malloc(self.nloc * 8) for _ in range(self.nbuf)]
            self.ptr = [p.value for p in self._raw]
            self.bufs = None
            handles = []
            for p in self._raw:
                hb = C.create_string_buffer(64)
                _lib.check(L.slb_ipc_get_handle(self.ctx.h, p, hb))
                handles.append(hb.raw)
            allh = [None] * self.P
            dist.all_gather_object(allh, handles, group=group)
            self.peer = []  # peer[r][i]: address of rank r's buffer i in this process
            self._opened = []
            for r in range(self.P):
                if r == self.rank:
                    self.peer.append(list(self.ptr))
                    continue
                row = []
                for hb in allh[r]:
                    q = C.c_void_p()
                    _lib.check(L.slb_ipc_open_handle(self.ctx.h, C.create_string_buffer(hb, 64), C.byref(q)))
                    row.append(q.value)
                    self._opened.append(q)
                self.peer.append(row)
            # barrier and rho all-gather of the data path: the library's own mailbox collectives (slb_comm_*, peer
            # stores + flags on the driver's stream) -- torch.distributed only carried the handles above
            hc = C.c_void_p()
            _lib.check(L.slb_comm_create(self.ctx.h, self.rank, self.P, n1 * n2 // self.P, C.byref(hc)))
            self.comm = hc
            hb = C.create_string_buffer(128)
            _lib.check(L.slb_comm_export(self.comm, hb))
            allm = [None] * self.P
            dist.all_gather_object(allm, hb.raw, group=group)
            blob = b"".join(allm)
            _lib.check(L.slb_comm_connect(self.comm, C.create_string_buffer(blob, len(blob))))
        else:
            self.nbuf = 2
            self.bufs = [torch.empty(self.nloc, dtype=torch.float64, device=self.device) for _ in range(2)]
            self.ptr = [b.data_ptr() for b in self.bufs]
        if self.exchange != "p2p":
            self.comm = None
            self._flag = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.cur = 0  # index of the buffer holding f
        self.layout = LAYOUT_B
        self.since_barrier = 99  # sweeps since the last cross-rank barrier
        shp = local_shape(self.gshape, LAYOUT_B, self.P)
        if tuple(data_local_B.shape) != shp:
            raise ValueError(f"local slab shape {tuple(data_local_B.shape)} must be {shp}")
        host = np.ascontiguousarray(np.asfortranarray(data_local_B).reshape(-1, order="F"))
        _lib.check(L.slb_memcpy_h2d(self.ctx.h, C.c_void_p(self.ptr[0]), host.ctypes.data_as(C.c_void_p), host.nbytes))
        self.ctx.sync()
        self._grids = {}
        self.state_gen = 1
        self.time_cur = 0.0
        self.flags = 0
        # Poisson pieces (replicated solve)
        from .poisson import _get_fctv_k_imag

        self.Nsp = 2
        self.fctv = [np.ascontiguousarray(a.reshape(-1, order="F")) for a in _get_fctv_k_imag(adv)]
        arr = (_lib.c_double_p * 2)(*[_lib.dptr(a) for a in self.fctv])
        h = C.c_void_p()
        _lib.check(L.slb_poisson_create(self.ctx.h, 2, _lib.i64((n1, n2)), arr, C.byref(h)))
        self.plan = h
        self.rho_local = torch.empty(n1 * n2 // self.P, dtype=torch.float64, device=self.device)
        self.rho = torch.empty(n1 * n2, dtype=torch.float64, device=self.device)
        self.E = [torch.empty(n1 * n2, dtype=torch.float64, device=self.device) for _ in range(2)]
        self.points = [torch.from_numpy(np.ascontiguousarray(m.points)).to(self.device) for m in adv.t_mesh]
        self.linesum = torch.empty(self.nloc // min(n3, n4), dtype=torch.float64, device=self.device)
        self.linesum_dim = None
        self.use_linesum = True
        import os as _os

        self.fuse_pairs = _os.environ.get("SLB_FUSE", "1") != "0"   # v1v2 / x1x2 in one pass each (slb_sweep_pair_ex)
        self._pending = None
        self.n_fused = 0
        self.rotate_march = _os.environ.get("SLB_ROTATE", "1") != "0"  # ranks start their marches at different blocks
        self.has_field = False
        self.n_exchanges = 0
        self.n_barriers = 0
        torch.cuda.synchronize(self.device)  # buffers above were filled on torch's default stream
        if self.P > 1:
            dist.barrier(group=group)

    # ---- grid handles over the buffers ---------------------------------------------------
    def _grid(self, layout, out=None):
        out = (self.cur + 1) % self.nbuf if out is None else out
        key = (layout, self.cur, out)
        g = self._grids.get(key)
        if g is None:
            shp = local_shape(self.gshape, layout, self.P)
            g = C.c_void_p()
            _lib.check(_lib.lib().slb_grid_create_external(
                self.ctx.h, 4, _lib.i64(shp), C.c_void_p(self.ptr[self.cur]), C.c_void_p(self.ptr[out]), C.byref(g)))
            self._grids[key] = g
        return g

    def _barrier(self):
        """stream-ordered cross-rank barrier (called under the driver's stream)"""
        if self.comm is not None:
            _lib.check(_lib.lib().slb_comm_barrier(self.comm))
        else:
            self.dist.all_reduce(self._flag, group=self.group)
        self.since_barrier = 0
        self.n_barriers += 1

    def _exchange_nccl(self):
        # called under `with torch.cuda.stream(self.stream)`
        self.dist.all_to_all_single(self.bufs[1 - self.cur], self.bufs[self.cur], group=self.group)
        self.cur = 1 - self.cur
        self.n_exchanges += 1

    # ---- state machine (src/advection.jl:152-158, :358-367) ------------------------------
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

    # ---- field solve (src/poisson.jl:119-144) ---------------------------------------------
    def compute_field(self):
        with self.torch.cuda.stream(self.stream):
            self._compute_field()

    def _compute_field(self):
        self.flush()
        if self.layout != LAYOUT_B:
            raise RuntimeError("the charge density is a local reduction in layout B only")
        adv = self.adv
        n1, n2, n3, n4 = self.gshape
        dv = adv.t_mesh[2].step * adv.t_mesh[3].step
        L = _lib.lib()
        rl = C.c_void_p(self.rho_local.data_ptr())
        if self.linesum_dim is not None:
            # the last velocity sweep left sum over that dim per line: reduce the remaining one
            nv_rest = n3 * n4 // self.gshape[self.linesum_dim]
            _lib.check(L.slb_charge_density_from(self.ctx.h, C.c_void_p(self.linesum.data_ptr()), n1 * n2 // self.P, nv_rest, dv, rl, 0))
        else:
            _lib.check(L.slb_charge_density_raw(self._grid(LAYOUT_B), 2, dv, rl))
        arr = (C.c_void_p * 2)(*[e.data_ptr() for e in self.E])
        if self.comm is not None and self.P > 1:
            # the slabs of rho land side by side in this rank's mailbox: the gathered array IS rho (x2 is the slowest
            # space dim); the solve reads it from there and leaves the mean-free rho in self.rho
            slots = C.c_void_p()
            _lib.check(L.slb_comm_allgather(self.comm, rl, n1 * n2 // self.P, C.byref(slots)))
            _lib.check(L.slb_poisson_solve_partial(self.plan, slots, 1, 1.0, 1, C.c_void_p(self.rho.data_ptr()), arr))
            self.has_field = True
            return
        if self.P > 1:
            self.dist.all_gather_into_tensor(self.rho, self.rho_local, group=self.group)
        else:
            self.rho.copy_(self.rho_local)
        _lib.check(L.slb_poisson_solve_raw(self.plan, C.c_void_p(self.rho.data_ptr()), 1, arr))  # mean removal + all DFT passes: one kernel
        self.has_field = True

    def compute_ee(self):
        """src/util_poisson.jl:156-162 (replicated: every rank returns the same value)"""
        adv = self.adv
        dx = adv.t_mesh[0].step * adv.t_mesh[1].step
        tot = 0.0
        for e in self.E:
            v = C.c_double()
            _lib.check(_lib.lib().slb_reduce_sumsq(self.ctx.h, C.c_void_p(e.data_ptr()), e.numel(), C.byref(v)))
            tot += v.value
        return dx * tot

    # ---- one advection! call ---------------------------------------------------------------
    def advection(self):
        with self.torch.cuda.stream(self.stream):
            return self._advection()

    def _next_dim(self):
        adv = self.adv
        return adv.getst(self.state_gen + 1 if self.state_gen < adv.nbstates else 1).perm[0] - 1

    def _stage_table(self, d, dt):
        """alpha table of a stage along dim d for this rank's slab: (device pointer, length, strides, scale)"""
        adv = self.adv
        n1, n2, n3, n4 = self.gshape
        c2, c4 = n2 // self.P, n4 // self.P
        strides = [0, 0, 0, 0]
        if d >= 2:  # velocity sweep: alpha = (dt/dv_d) * E_{d-2}[x1, x2l + r*c2]   (src/poisson.jl:178-189)
            if not self.has_field:
                raise RuntimeError("velocity state before any field solve")
            strides[0], strides[1] = 1, n1
            tab = self.E[d - 2].data_ptr() + 8 * self.rank * c2 * n1
            return tab, c2 * n1, strides, dt / adv.t_mesh[d].step
        src = d + 2  # space sweep: alpha = (-dt/dx_d) * v_{d+2}   (src/poisson.jl:191-203)
        strides[src] = 1
        off = self.rank * c4 if src == 3 else 0
        tab = self.points[src].data_ptr() + 8 * off
        return tab, (c4 if src == 3 else self.gshape[src]), strides, -dt / adv.t_mesh[d].step

    def _pairable(self, dA, dB):
        from .advection import FUSED_ORDERS
        from .interp import HERMITE, LAGRANGE

        if not self.fuse_pairs or dB != dA + 1 or dA not in (0, 2):
            return False
        itA, itB = self.adv.t_interp[dA], self.adv.t_interp[dB]
        ok = lambda it: it.kind in (LAGRANGE, HERMITE) and it.tabfct.shape[1] <= 14
        return ok(itA) and ok(itB) and itA.order == itB.order and itA.order in FUSED_ORDERS

    def flush(self):
        """run a stage that was held back for pair fusion (no-op otherwise)"""
        pend, self._pending = self._pending, None
        if pend is not None:
            self._single(*pend)

    def _advection(self):
        st = self.getst()
        d = st.perm[0] - 1
        dt = self.getcur_t()
        nxt = self._next_dim()
        if self._pending is not None:
            dA, dtA, _ = self._pending
            if dA + 1 == d:
                self._pending = None
                self._pair(dA, dtA, d, dt, nxt)
                return self.nextstate()
            self.flush()
        if d == 2:
            self._compute_field()
        if self.state_gen < self.adv.nbstates and self._pairable(d, nxt):
            self._pending = (d, dt, nxt)  # held back: the next call runs both stages in one pass
            return self.nextstate()
        self._single(d, dt, nxt)
        return self.nextstate()

    def _enter_layout(self, d):
        """layout bookkeeping before a pass whose first sweep runs along dim d; returns True when the
        current buffer holds the blocks of a B -> A exchange (block-major along x2)"""
        need = LAYOUT_A if d < 2 else LAYOUT_B
        blocked_in = False
        if need != self.layout:
            # only B -> A is ever pending here: A -> B is completed by the x2 sweep itself
            if not (need == LAYOUT_A and d == 0):
                raise NotImplementedError("unsupported state order for the fused re-shard (needs x1 first after v-sweeps)")
            blocked_in = self.P > 1
            self.layout = LAYOUT_A
        return blocked_in

    def _pair(self, dA, dtA, dB, dtB, nxt):
        """stages dA, dB = dA + 1 in ONE pass over HBM (slb_sweep_pair_ex), with the re-shard that
        follows dB fused into its stores"""
        adv = self.adv
        L = _lib.lib()
        multi = self.P > 1
        blocked_in = self._enter_layout(dA)
        tA, lA, sA, scA = self._stage_table(dA, dtA)
        tB, lB, sB, scB = self._stage_table(dB, dtB)
        to_A = self.layout == LAYOUT_B and nxt < 2
        to_B = self.layout == LAYOUT_A and nxt >= 2
        if (to_A and dB != 3) or (to_B and dB != 1):
            raise NotImplementedError("unsupported state order for the fused re-shard")
        reshard = multi and (to_A or to_B)
        self.linesum_dim = None
        want_ls = self.use_linesum and dB == 3 and nxt == 2
        out = (self.cur + 1) % self.nbuf
        g = self._grid(self.layout, out)
        hA = adv.t_interp[dA].handle(self.ctx, self.gshape[dA])
        hB = adv.t_interp[dB].handle(self.ctx, self.gshape[dB])
        if want_ls:
            _lib.check(L.slb_grid_set_linesum(g, C.c_void_p(self.linesum.data_ptr())))
        try:
            args = (g, dA, hA, C.c_void_p(tA), lA, _lib.i64(sA), float(scA), dB, hB, C.c_void_p(tB), lB, _lib.i64(sB), float(scB), 1,
                    int(self.flags), self.P if blocked_in else 1)
            if reshard and self.exchange == "p2p":
                if self.since_barrier > 1:
                    self._barrier()  # every rank is done with the buffer we are about to store into
                blk = self.nloc // self.P
                bases = (C.c_void_p * self.P)(*[self.peer[q][out] + 8 * self.rank * blk for q in range(self.P)])
                _lib.check(L.slb_sweep_pair_ex(*args, self.P, bases, (self.rank + 1) % self.P if self.rotate_march else 0))
                self._barrier()      # all blocks have landed everywhere
                self.n_exchanges += 1
                self.cur = out
            else:
                _lib.check(L.slb_sweep_pair_ex(*args, self.P if reshard else 1, None, 0))
                _lib.check(L.slb_grid_swap(g))  # keep the handle's orientation; the driver tracks `cur`
                self.cur = out
                self.since_barrier += 1
                if reshard:
                    self._exchange_nccl()
        finally:
            if want_ls:
                _lib.check(L.slb_grid_set_linesum(g, None))
        if want_ls:
            self.linesum_dim = dB
        if to_B:
            self.layout = LAYOUT_B
        self.n_fused += 1

    def _single(self, d, dt, nxt):
        adv = self.adv
        interp = adv.t_interp[d]
        L = _lib.lib()
        multi = self.P > 1
        mode, bdim = _lib.SLB_RESHARD_NONE, 0
        if self._enter_layout(d):
            mode, bdim = _lib.SLB_RESHARD_IN_BLOCKED, 1   # cur holds the blocks received from every rank
        tab, tlen, strides, scale = self._stage_table(d, dt)
        # layout change AFTER this sweep: v2 followed by x1 (B -> A), x2 followed by a v-sweep (A -> B)
        to_A = self.layout == LAYOUT_B and nxt < 2
        to_B = self.layout == LAYOUT_A and nxt >= 2
        if (to_A and d != 3) or (to_B and d != 1):
            raise NotImplementedError("unsupported state order for the fused re-shard (v2 must precede x-sweeps, x2 must precede v-sweeps)")
        reshard = multi and (to_A or to_B)
        # line sums of a velocity sweep feed the next charge density (next state = v1)
        self.linesum_dim = None
        want_ls = self.use_linesum and d >= 2 and nxt == 2
        out = (self.cur + 1) % self.nbuf
        g = self._grid(self.layout, out)
        h = interp.handle(self.ctx, self.gshape[d])
        if want_ls:
            _lib.check(L.slb_grid_set_linesum(g, C.c_void_p(self.linesum.data_ptr())))
        try:
            if reshard and self.exchange == "p2p":
                if mode != _lib.SLB_RESHARD_NONE:
                    raise NotImplementedError("a sweep cannot both read and write re-sharded data")
                if self.since_barrier > 1:
                    self._barrier()  # every rank is done with the buffer we are about to store into
                blk = self.nloc // self.P
                bases = (C.c_void_p * self.P)(*[self.peer[q][out] + 8 * self.rank * blk for q in range(self.P)])
                _lib.check(L.slb_sweep_peer(g, d, h, C.c_void_p(tab), tlen, _lib.i64(strides), float(scale), 1, int(self.flags),
                                            self.P, bases))
                self._barrier()      # all blocks have landed everywhere
                self.n_exchanges += 1
                self.cur = out
            else:
                if reshard and to_B:
                    if mode != _lib.SLB_RESHARD_NONE:
                        raise NotImplementedError("a sweep cannot both read and write re-sharded data")
                    mode, bdim = _lib.SLB_RESHARD_OUT_BLOCKED, 1
                _lib.check(L.slb_sweep_ex(g, d, h, C.c_void_p(tab), tlen, _lib.i64(strides), float(scale), 1, int(self.flags),
                                          mode, bdim, self.P))
                _lib.check(L.slb_grid_swap(g))  # keep the handle's orientation; the driver tracks `cur`
                self.cur = out
                self.since_barrier += 1
                if reshard:
                    self._exchange_nccl()       # B -> A: contiguous slabs; A -> B: block-major output
        finally:
            if want_ls:
                _lib.check(L.slb_grid_set_linesum(g, None))
        if want_ls:
            self.linesum_dim = d
        if to_B:
            self.layout = LAYOUT_B

    # ---- data access ----------------------------------------------------------------------
    def upload_local(self, host_flat):
        """copy this rank's slab (current layout, flat Fortran order, e.g. pinned memory) to the device"""
        self._pending = None
        _lib.check(_lib.lib().slb_memcpy_h2d(self.ctx.h, C.c_void_p(self.ptr[self.cur]), host_flat.ctypes.data_as(C.c_void_p), self.nloc * 8))
        self.linesum_dim = None

    def download_local(self, host_flat):
        if self._pending is not None:
            with self.torch.cuda.stream(self.stream):
                self.flush()
        _lib.check(_lib.lib().slb_memcpy_d2h(self.ctx.h, host_flat.ctypes.data_as(C.c_void_p), C.c_void_p(self.ptr[self.cur]), self.nloc * 8))
        self.ctx.sync()

    def getdata_local(self):
        """this rank's slab in the current layout (numpy, Fortran order); valid between steps"""
        flat = np.empty(self.nloc, dtype=np.float64)
        self.download_local(flat)
        return flat.reshape(local_shape(self.gshape, self.layout, self.P), order="F")

    def gather_global(self):
        """full array on every rank (tests only)"""
        loc = self.getdata_local()
        t = self.torch.from_numpy(np.ascontiguousarray(loc.reshape(-1, order="F"))).to(self.device)
        out = [self.torch.empty_like(t) for _ in range(self.P)]
        if self.P > 1:
            self.dist.all_gather(out, t, group=self.group)
            self.torch.cuda.synchronize(self.device)
        else:
            out = [t]
        d = SHARD_DIM[self.layout]
        shp = local_shape(self.gshape, self.layout, self.P)
        parts = [o.cpu().numpy().reshape(shp, order="F") for o in out]
        return np.asfortranarray(np.concatenate(parts, axis=d))

    def close(self):
        L = _lib.lib()
        self.ctx.sync()
        for g in self._grids.values():
            L.slb_grid_destroy(g)
        self._grids = {}
        if self.plan:
            L.slb_poisson_destroy(self.plan)
            self.plan = None
        if self.exchange == "p2p" and self._raw:
            if self.comm is not None:
                L.slb_comm_destroy(self.comm)
                self.comm = None
            if self.P > 1:
                self.dist.barrier(group=self.group)
            for q in self._opened:
                L.slb_ipc_close_handle(self.ctx.h, q)
            self._opened = []
            if self.P > 1:
                self.dist.barrier(group=self.group)
            for p in self._raw:
                self.ctx.free(p)
            self._raw = []
