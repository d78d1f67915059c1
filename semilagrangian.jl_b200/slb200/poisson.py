"""Vlasov-Poisson displacement provider -- host mirror of src/poisson.jl + src/util_poisson.jl.

PoissonConst/PoissonVar (src/poisson.jl:35-95), initcoef! (:164-205), getalpha (:210-224),
compute_charge! (src/util_poisson.jl:68-79), compute_elfield! (src/poisson.jl:139-144),
compute_ee / compute_ke / getenergy* (src/util_poisson.jl:41-53, :156-183).
rho, E and the mesh nodes live on the device; a velocity sweep reads E directly
(alpha = (dt/dv) * E[x]) and a space sweep reads the velocity nodes (alpha = (-dt/dx) * v).
"""
import ctypes as C

import numpy as np

from . import _lib
from .advection import AbstractExtDataAdv
from .mesh import vec_k_fft
from .unsplit2d import DeviceField

# @enum TypePoisson, src/poisson.jl:2
StdPoisson, StdPoisson2d, StdABp = 1, 2, 3


def _get_fctv_k_imag(adv):
    """Imaginary parts of fctv_k (src/poisson.jl:7-15): k_x / |k|^2, zero mode 0."""
    nsp = adv.N // 2
    v_k = [vec_k_fft(m) for m in adv.t_mesh[:nsp]]
    sz = [len(m) for m in adv.t_mesh[:nsp]]
    s = np.zeros(sz, order="F")
    for x in range(nsp):
        shape = [1] * nsp
        shape[x] = sz[x]
        s = s + (v_k[x] ** 2).reshape(shape)
    with np.errstate(divide="ignore"):
        inv = 1.0 / s
    inv.reshape(-1, order="F")[0] = 0.0
    out = []
    for x in range(nsp):
        shape = [1] * nsp
        shape[x] = sz[x]
        out.append(np.asfortranarray(v_k[x].reshape(shape) * inv))
    return out


def dotprod(vs):
    """src/util.jl:59-67"""
    n = len(vs)
    res = np.ones([1] * n)
    for i, v in enumerate(vs):
        shape = [1] * n
        shape[i] = len(v)
        res = res * np.asarray(v, dtype=np.float64).reshape(shape)
    return np.asfortranarray(res)


class PoissonVar(AbstractExtDataAdv):
    """getpoissonvar(adv; type): PoissonConst + PoissonVar (src/poisson.jl:35-107).  type =
    StdPoisson: split sweeps (initcoef! :164-205); StdPoisson2d: one unsplit 2-D state with per-point
    shifts (initcoef! :229-247)."""

    def __init__(self, adv, ctx=None, type=StdPoisson):  # noqa: A002 (the reference's keyword)
        N = adv.N
        if type not in (StdPoisson, StdPoisson2d):
            raise ValueError("TypePoisson must be StdPoisson or StdPoisson2d (StdABp has no initcoef! in the reference)")
        if type == StdPoisson2d and N != 2:
            raise ValueError("StdPoisson2d needs a 1D1V grid")
        self.type = type
        if N % 2 != 0:
            raise ValueError(f"N={N} must be a multiple of 2")
        self.adv = adv
        self.ctx = ctx or adv.ctx or _lib.default_context()
        self.Nsp = self.Nv = N // 2
        self.sp_ext = adv.sizeall[: self.Nsp]
        self.nsp_tot = int(np.prod(self.sp_ext))
        self.fctv_imag = _get_fctv_k_imag(adv)
        self.v_square = dotprod([m.points for m in adv.t_mesh[self.Nsp:]]) ** 2  # src/poisson.jl:51
        L = _lib.lib()
        self._keep = [np.ascontiguousarray(a.reshape(-1, order="F")) for a in self.fctv_imag]
        arr = (_lib.c_double_p * self.Nsp)(*[_lib.dptr(a) for a in self._keep])
        h = C.c_void_p()
        _lib.check(L.slb_poisson_create(self.ctx.h, self.Nsp, _lib.i64(self.sp_ext), arr, C.byref(h)))
        self.plan = h
        self.rho_dev = self.ctx.malloc(self.nsp_tot * 8)
        self.E_dev = [self.ctx.malloc(self.nsp_tot * 8) for _ in range(self.Nsp)]
        self.vsq_dev = self.ctx.to_device(self.v_square.reshape(-1, order="F"))
        self.has_field = False
        self._sweep = None

    # src/poisson.jl:119-125 -> src/util_poisson.jl:68-79
    def compute_charge(self, advd):
        advd.flush()  # a stage held back for pair fusion must have run before f is reduced
        dv = 1.0
        for m in self.adv.t_mesh[self.Nsp:]:
            dv = dv * m.step
        d = getattr(advd, "_linesum_dim", None)
        if d is not None and d >= self.Nsp:
            # the previous sweep ran along velocity dim d and left sum_d f per line: rho needs only
            # the remaining velocity dims (1/n_d of the traffic of a pass over f)
            nv_rest = int(np.prod(self.adv.sizeall[self.Nsp:])) // self.adv.sizeall[d]
            _lib.check(_lib.lib().slb_charge_density_from(self.ctx.h, advd._linesum, self.nsp_tot, nv_rest, dv, self.rho_dev, 1))
            return
        _lib.check(_lib.lib().slb_charge_density(advd.grid, self.Nsp, dv, self.rho_dev))

    def field_solve(self, advd):
        """compute_charge! + compute_elfield! of one initcoef! (src/poisson.jl:171-176) as two
        launches: the reduction over f (or over the line sums the last velocity pass left) and one
        cooperative kernel for the rest (slb_vp_field_solve)."""
        advd.flush()
        dv = 1.0
        for m in self.adv.t_mesh[self.Nsp:]:
            dv = dv * m.step
        L = _lib.lib()
        d = getattr(advd, "_linesum_dim", None)
        nv = int(np.prod(self.adv.sizeall[self.Nsp:]))
        if d is not None and d >= self.Nsp:
            src, nv = advd._linesum, nv // self.adv.sizeall[d]
        elif getattr(advd, "_rhopart_planes", 0):
            # the fused space pass that just ran left partial planes of the charge density (1/4 of the bytes of f)
            src, nv = advd._rhopart, advd._rhopart_planes
        else:
            src = C.c_void_p(L.slb_grid_front(advd.grid))
        arr = (C.c_void_p * self.Nsp)(*[p.value for p in self.E_dev])
        _lib.check(L.slb_vp_field_solve(self.plan, src, nv, dv, self.rho_dev, arr))
        self.has_field = True

    def wants_rhopart(self, advd, dA, dB):
        """True when the fused pass over (dA, dB) sweeps exactly the two space dims of a 2D2V grid and the NEXT
        advection! call starts with compute_charge! (src/poisson.jl:171-174): the pass can then leave partial planes
        of rho (slb_grid_set_rhopart) and no separate reduction over f is needed."""
        adv = advd.adv
        if self.type != StdPoisson or self.Nsp != 2 or sorted((dA, dB)) != [0, 1]:
            return False
        nxt = adv.getst(advd.state_gen + 1 if advd.state_gen < adv.nbstates else 1)
        return nxt.perm[0] > self.Nsp and (self.Nsp + 1) in nxt.perm[: nxt.ndims]

    def wants_linesum(self, advd):
        """True when the NEXT advection! call starts with compute_charge! (src/poisson.jl:171-174)
        and the current sweep runs along a velocity dim, so its line sums can feed it."""
        adv = advd.adv
        st = advd.getst()
        if st.perm[0] <= self.Nsp:
            return False
        nxt = adv.getst(advd.state_gen + 1 if advd.state_gen < adv.nbstates else 1)
        return nxt.perm[0] > self.Nsp and (self.Nsp + 1) in nxt.perm[: nxt.ndims]

    # src/poisson.jl:139-144
    def compute_elfield(self):
        arr = (C.c_void_p * self.Nsp)(*[p.value for p in self.E_dev])
        _lib.check(_lib.lib().slb_poisson_solve(self.plan, self.rho_dev, arr))
        self.has_field = True

    @property
    def rho(self):
        return self.ctx.to_host(self.rho_dev, self.nsp_tot).reshape(self.sp_ext, order="F")

    @property
    def t_elfield(self):
        if not self.has_field:
            return None
        return tuple(self.ctx.to_host(p, self.nsp_tot).reshape(self.sp_ext, order="F") for p in self.E_dev)

    def isvelocity(self, advd):  # src/poisson.jl:155-158
        return advd.getst().perm[0] > self.Nsp

    def initcoef_reads_data(self, advd):
        """initcoef! computes the charge density only at a velocity state that contains dim Nsp+1
        (src/poisson.jl:171-176): every other state can be pair-fused with its predecessor."""
        st = advd.getst()
        return self.isvelocity(advd) and (self.Nsp + 1) in st.perm[: st.ndims]

    def _initcoef_2d(self, advd):
        """initcoef!(pv::PoissonVar{..,StdPoisson2d}, advd) -- src/poisson.jl:229-247:
        bufcur[i, j] = ((-dt/dx) * v_j, (dt/dv) * E_i), filled on the device from the resident E and
        velocity nodes."""
        adv = advd.adv
        self.field_solve(advd)
        dt = advd.getcur_t()
        n1, n2 = adv.sizeall
        if advd.bufcur is None:
            advd.bufcur = DeviceField(self.ctx, n1, n2, 2)
        _lib.check(_lib.lib().slb_fill_dec2d(self.ctx.h, advd.bufcur.ptr, n1, n2, advd.points_dev(1), -dt / adv.t_mesh[0].step,
                                             self.E_dev[0], dt / adv.t_mesh[1].step))

    def initcoef(self, advd):
        """initcoef!(pv, advd) -- src/poisson.jl:164-205"""
        if self.type == StdPoisson2d:
            return self._initcoef_2d(advd)
        st = advd.getst()
        adv = advd.adv
        dt = advd.getcur_t()
        Nsp, N = self.Nsp, adv.N
        strides = [0] * N
        if self.isvelocity(advd):
            if (Nsp + 1) in st.perm[: st.ndims]:
                self.field_solve(advd)
            if not self.has_field:
                raise RuntimeError("velocity state before any field solve (state order must start with dim Nsp+1)")
            # bufcur_v[x] = (dt / step(mesh_d)) * E_{d-Nsp}, d = perm[x]; indexed by ind.I[end-Nsp+1:end],
            # i.e. E's axis i is the grid dim perm[N-Nsp+i]  (src/poisson.jl:178-189, :216-219)
            stride = 1
            for i in range(Nsp):
                g = st.perm[N - Nsp + i]
                strides[g - 1] = stride
                stride *= self.sp_ext[i]
            self._sweeps = []
            for x in range(st.ndims):
                d = st.perm[x]
                self._sweeps.append(((self.E_dev[d - 1 - Nsp], self.nsp_tot), list(strides), dt / adv.t_mesh[d - 1].step, True))
        else:
            # tupleind / bufcur_sp, src/poisson.jl:191-203 (invp where perm is meant: identical for
            # the involutive permutations every reference driver uses)
            self._sweeps = []
            for x in range(st.ndims):
                tupleind = st.perm[st.invp[x] + Nsp - 1] - st.ndims
                g = st.perm[st.ndims + tupleind - 1]
                src_dim = st.invp[x] + Nsp
                sx = [0] * N
                sx[g - 1] = 1
                self._sweeps.append(((advd.points_dev(src_dim - 1), adv.sizeall[src_dim - 1]), sx,
                                     -dt / adv.t_mesh[st.invp[x] - 1].step, True))
        self._sweep = self._sweeps[0]

    def alpha_table_nd(self, advd):
        return self._sweeps

    def alpha_table(self, advd):
        return self._sweep

    def close(self):
        if getattr(self, "plan", None):
            _lib.lib().slb_poisson_destroy(self.plan)
            self.plan = None
            for p in [self.rho_dev, self.vsq_dev] + list(self.E_dev):
                self.ctx.free(p)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def getpoissonvar(adv, ctx=None, type=StdPoisson, typeadd=0):  # noqa: A002, ARG001
    """getpoissonvar(adv; type = StdPoisson, typeadd = 0) -- src/poisson.jl:100-103"""
    return PoissonVar(adv, ctx=ctx, type=type)


def compute_ee(advd):
    """compute_ee(advd) -- src/util_poisson.jl:156-162: dx * sum_d sum(E_d .^ 2), with the
    field of the last compute_elfield! (the reference does not recompute it)."""
    pv = advd.parext
    dx = 1.0
    for m in advd.adv.t_mesh[: pv.Nsp]:
        dx = dx * m.step
    tot = 0.0
    for p in pv.E_dev:
        v = C.c_double()
        _lib.check(_lib.lib().slb_reduce_sumsq(advd.ctx.h, p, pv.nsp_tot, C.byref(v)))
        tot = tot + v.value
    return dx * tot


def compute_ke(advd):
    """compute_ke(advd) -- src/util_poisson.jl:41-53"""
    pv = advd.parext
    adv = advd.adv
    dsp = 1.0
    for m in adv.t_mesh[: pv.Nsp]:
        dsp *= m.step
    dv = 1.0
    for m in adv.t_mesh[pv.Nsp:]:
        dv *= m.step
    advd.flush()
    v = C.c_double()
    _lib.check(_lib.lib().slb_kinetic_energy(advd.grid, pv.Nsp, pv.vsq_dev, dsp * dv, C.byref(v)))
    return v.value


def getenergy(advd):
    """src/util_poisson.jl:167-175"""
    pv = advd.parext
    pv.compute_charge(advd)
    pv.compute_elfield()
    ee = compute_ee(advd)
    ke = compute_ke(advd)
    return ee, ke, ee + ke


def getenergyall(advd):
    return getenergy(advd)[2]
