"""Unsplit 2-D advection: per-point N-D interpolation and the Adams-Bashforth time algorithms
 -- host mirror of
  interpolate!(fp, fi, bufdec, interp_t)        src/interpolation.jl:561-621 (closure form :401-429)
  autointerp! / interpbufc!                     src/interpolation.jl:626-682
  decbegin! / initcoef!(::AdvectionData)        src/advection.jl:391-580
  the single-state branch of advection!         src/advection.jl:607-619
  ABcoef                                        src/lagrange.jl:74-88
Every array lives on the device: a field of OpTuple{2,T} (src/util.jl:3-23) is a DeviceField with
two component planes [n1, n2, 2]; all arithmetic is done by libslb200 (slb_interp2d_points,
slb_lincomb, slb_fill_dec2d).  The host only sequences the calls, as the reference's driver does.
"""
import ctypes as C
from fractions import Fraction

import numpy as np

from . import _lib
from .interp import BSPLINE_FFT, BSPLINE_LU

# @enum TimeAlgorithm, src/advection.jl:2
NoTimeAlg, ABTimeAlg_ip, ABTimeAlg_new, ABTimeAlg_init = 1, 2, 3, 4


def abcoef(ordermax):
    """ABcoef(ordermax).tab (src/lagrange.jl:74-88) as exact fractions, 0-based [i][j]:
    tab[i][j] = integral over [-1, 0] of the Lagrange basis polynomial i on the nodes 0..j
    (the Adams-Bashforth weights of order j + 1).  Closed form through exact polynomial products."""
    tab = [[Fraction(0)] * ordermax for _ in range(ordermax)]
    for j in range(ordermax):
        for i in range(j + 1):
            poly = [Fraction(1)]  # ascending coefficients of prod_{l != i} (x - l) / (i - l)
            for l in range(j + 1):
                if l == i:
                    continue
                d = Fraction(i - l)
                nxt = [Fraction(0)] * (len(poly) + 1)
                for k, c in enumerate(poly):
                    nxt[k] += c * Fraction(-l) / d
                    nxt[k + 1] += c / d
                poly = nxt
            # integral from -1 to 0 of sum c_k x^k = - sum c_k (-1)^(k+1) / (k+1)
            tab[i][j] = -sum(c * Fraction((-1) ** (k + 1), k + 1) for k, c in enumerate(poly))
    return tab


class DeviceField:
    """[n1, n2, ncomp] Float64 planes on the device (ncomp = 1: scalar field, 2: OpTuple{2}).
    Storage comes from a per-context free list: the time algorithms create and drop many short-lived
    fields per step, all used on the context's one stream, so a block can be reused without
    synchronising (cudaMalloc / cudaFree would serialise every step)."""

    _pool = {}

    def __init__(self, ctx, n1, n2, ncomp, ptr=None):
        self.ctx, self.n1, self.n2, self.ncomp = ctx, int(n1), int(n2), int(ncomp)
        self.numel = self.n1 * self.n2 * self.ncomp
        self.owned = ptr is None
        if ptr is None:
            free = DeviceField._pool.get((ctx.serial, self.numel))
            ptr = free.pop() if free else ctx.malloc(self.numel * 8)
        self.ptr = ptr

    @classmethod
    def view(cls, ctx, n1, n2, ncomp, ptr):
        """non-owning view of device memory held elsewhere (the grid's front buffer)"""
        return cls(ctx, n1, n2, ncomp, ptr=ptr)

    @classmethod
    def from_host(cls, ctx, arr):
        arr = np.asarray(arr, dtype=np.float64)
        if arr.ndim == 2:
            arr = arr.reshape(arr.shape + (1,))
        f = cls(ctx, *arr.shape)
        f.upload(arr)
        return f

    def upload(self, arr):
        host = np.asfortranarray(arr, dtype=np.float64).reshape(-1, order="F")
        assert host.size == self.numel
        _lib.check(_lib.lib().slb_memcpy_h2d(self.ctx.h, self.ptr, host.ctypes.data_as(C.c_void_p), host.nbytes))
        self.ctx.sync()

    def to_host(self):
        out = self.ctx.to_host(self.ptr, self.numel).reshape((self.n1, self.n2, self.ncomp), order="F")
        return out[:, :, 0].copy(order="F") if self.ncomp == 1 else out

    def like(self):
        return DeviceField(self.ctx, self.n1, self.n2, self.ncomp)

    def assign(self, other):
        """self .= other"""
        assert other.numel == self.numel
        if other.ptr.value != self.ptr.value:
            _lib.check(_lib.lib().slb_memcpy_d2d(self.ctx.h, self.ptr, other.ptr, self.numel * 8))
        return self

    def copy(self):
        return self.like().assign(self)

    def swap(self, other):
        assert self.owned and other.owned
        self.ptr, other.ptr = other.ptr, self.ptr

    def free(self):
        """return the storage to the free list (stream-ordered reuse)"""
        if self.ptr is not None and self.owned:
            key = (self.ctx.serial, self.numel)
            if not any(k[0] == self.ctx.serial for k in DeviceField._pool):
                ctx = self.ctx
                ctx.on_close(lambda: DeviceField.release_pool(ctx))  # pooled blocks die with their context
            DeviceField._pool.setdefault(key, []).append(self.ptr)
        self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    @classmethod
    def release_pool(cls, ctx):
        """cudaFree every pooled block of a context"""
        for key in [k for k in cls._pool if k[0] == ctx.serial]:
            for p in cls._pool.pop(key):
                ctx.free(p)


def lincomb(out, coefs, fields):
    """out = sum_k coefs[k] * fields[k], rounded products summed left to right (slb_lincomb)"""
    n = len(coefs)
    assert n == len(fields) and n >= 1
    cf = (C.c_double * n)(*[float(c) for c in coefs])
    ptrs = (C.c_void_p * n)(*[f.ptr.value for f in fields])
    _lib.check(_lib.lib().slb_lincomb(out.ctx.h, out.ptr, n, cf, ptrs, out.numel))
    return out


def scaled(field, s):
    """s * field as a new field (e.g. `sens * copy.(t_ref)`, `-fmrdec`)"""
    return lincomb(field.like(), [s], [field])


def _is_bspline(it):
    return getattr(it, "kind", None) in (BSPLINE_LU, BSPLINE_FFT)


def interpolate_points(dst, src, dec, interps, flags=0):
    """interpolate!(dst, src, dec, interp_t) on device fields (src/interpolation.jl:561-621).
    dst must not be src; src and dec are left untouched."""
    if dst is src or dst.ptr.value == src.ptr.value:
        raise ValueError("fp and fi must not alias")
    if len(interps) != 2:
        raise ValueError(f"The number of Interpolation {len(interps)} is different of N=2")
    ctx = dst.ctx
    h1 = interps[0].handle(ctx, src.n1)
    h2 = interps[1].handle(ctx, src.n2)
    work = tmp = None
    inp = src
    if _is_bspline(interps[0]) or _is_bspline(interps[1]):
        tmp = src.copy()  # the pre-solves overwrite their input
        work = src.like()
        inp = tmp
    try:
        _lib.check(_lib.lib().slb_interp2d_points(ctx.h, h1, h2, src.n1, src.n2, src.ncomp, inp.ptr, dec.ptr, dst.ptr,
                                                  work.ptr if work is not None else None, int(flags)))
    finally:
        if tmp is not None:
            tmp.free()
            work.free()
    return dst


def autointerp(to, frm, nb, interps, flags=0):
    """autointerp!(to, from, nb, interp_t) -- src/interpolation.jl:626-655"""
    if nb < 1:
        to.assign(frm)
    fmr = frm.copy()
    for i in range(1, nb + 1):
        interpolate_points(to, frm, fmr, interps, flags)
        if i != nb:
            fmr.assign(to)
    fmr.free()


def interpbufc(t_buf, bufdec, interps, nb=None, flags=0):
    """interpbufc!(t_buf, bufdec, interp_t, nb = length(t_buf)) -- src/interpolation.jl:661-682:
    buf <- interpolate(copy(buf), bufdec) for the last nb fields; done into a scratch field whose
    storage is then swapped with buf's (no copy)."""
    nb = len(t_buf) if nb is None else nb
    if nb == 0:
        return
    scratch = t_buf[-1].like()
    for i in range(nb):
        buf = t_buf[len(t_buf) - 1 - i]
        interpolate_points(scratch, buf, bufdec, interps, flags)
        buf.swap(scratch)
    scratch.free()


def _c(adv, k, n):
    fr = adv.abcoef[k - 1][n - 1]  # c(st::ABcoef, k, n) = st.tab[k, n], src/lagrange.jl:88
    return fr.numerator / fr.denominator


def _ab_sum(adv, fields, n, ord_):
    """sum(map(k -> c(abcoef, k, ord) * fields[k], 1:n)) as a new field"""
    return lincomb(fields[0].like(), [_c(adv, k, ord_) for k in range(1, n + 1)], fields[:n])


def decbegin(t_trv, t_cal, interps, flags=0):
    """decbegin!(t_trv, t_cal, t_interp) -- src/advection.jl:391-399"""
    indice = len(t_trv)
    for i in range(1, indice):
        buf = t_cal[-1]
        frm = buf.copy()
        autointerp(buf, frm, indice - 1, interps, flags)
        frm.free()
        interpbufc(t_trv, buf, interps, i, flags)
        t_cal.pop().free()


def _free_all(fields):
    for f in fields:
        f.free()


def initcoef(advd):
    """initcoef!(self::AdvectionData) -- src/advection.jl:404-580: the provider's initcoef! (which
    fills advd.bufcur on the device) followed by the Adams-Bashforth extrapolation of the
    displacement field along the characteristics."""
    nbtours = 3
    adv = advd.adv
    timealg, ordalg = adv.timealg, adv.ordalg
    interps = adv.t_interp
    fl = advd.flags
    isbegin = advd.bufcur is None
    ext = advd.parext
    ext.initcoef(advd)
    if advd.bufcur is None:
        raise RuntimeError("the provider's initcoef must set advd.bufcur for states with per-point shifts")

    if timealg == ABTimeAlg_new and isbegin:  # src/advection.jl:414-481
        t_ref, t_cal = [advd.bufcur.copy()], []
        svdata = advd.data_field().copy()
        svbufcur = advd.bufcur.copy()
        sens = 1 if (ordalg * nbtours) % 2 == 1 else -1
        for indice in range(1, ordalg + 1):
            for nb in range(1, (nbtours - 1 if indice == ordalg else nbtours) + 1):
                t_ref = t_ref[::-1]
                t_trv = [scaled(a, sens) for a in t_ref]
                decbegin(t_trv, t_cal, interps, fl)
                _free_all(t_cal)
                t_cal = []
                advd.data_field().assign(svdata)
                for i in range(1, indice + 1):
                    fmrdec = _ab_sum(adv, t_trv, indice, indice)
                    if i != 1:
                        t_cal.append(scaled(fmrdec, -1.0))
                    if i != 1 or nb != 1:
                        t_trv.pop().free()
                        t_ref.pop().free()
                    autointerp(advd.bufcur, fmrdec, indice, interps, fl)
                    fmrdec.free()
                    interpbufc(advd.t_bufc, advd.bufcur, interps, None, fl)
                    advd.interpolate_data()
                    ext.initcoef(advd)
                    t_trv.insert(0, advd.bufcur.copy())
                    t_ref.insert(0, advd.bufcur.copy())
                s = _ab_sum(adv, t_trv, indice + 1, indice + 1)
                t_cal.append(scaled(s, -1.0))
                s.free()
                _free_all(t_trv)
                sens = -sens
        assert sens == 1, "sens must be positive at this place"
        t_ref = t_ref[::-1]
        t_trv = [scaled(a, sens) for a in t_ref]
        _free_all(t_ref)
        t_trv.pop(0).free()
        decbegin(t_trv, t_cal, interps, fl)
        _free_all(t_cal)
        advd.t_bufc = t_trv
        advd.data_field().assign(svdata)
        advd.bufcur.assign(svbufcur)
        svdata.free()
        svbufcur.free()

    if timealg == ABTimeAlg_ip and isbegin:  # src/advection.jl:484-508
        for indice in range(1, ordalg):
            advd.t_bufc.insert(0, advd.bufcur.copy())
            fmrdec = _ab_sum(adv, advd.t_bufc, indice, indice)
            frm = fmrdec.copy()
            autointerp(fmrdec, frm, indice - 1, interps, fl)
            frm.free()
            interpbufc(advd.t_bufc, fmrdec, interps, None, fl)
            fmrdec.free()

    if timealg == ABTimeAlg_init and isbegin:  # src/advection.jl:509-543
        if advd.initdatas is None:
            raise ValueError("ABTimeAlg_init needs AdvectionData(...; initdatas)")
        for indice in range(1, len(advd.initdatas) + 1):
            advd.t_bufc.insert(0, advd.bufcur.copy())
            ord_ = min(indice, ordalg)
            fmrdec = _ab_sum(adv, advd.t_bufc, ord_, ord_)
            if ord_ == ordalg:
                advd.t_bufc.pop().free()
            frm = fmrdec.copy()
            autointerp(fmrdec, frm, ordalg - 1, interps, fl)
            frm.free()
            interpbufc(advd.t_bufc, fmrdec, interps, None, fl)
            fmrdec.free()
            advd.upload(advd.initdatas[indice - 1])
            advd.time_cur += advd.getcur_t()
            ext.initcoef(advd)

    if timealg in (ABTimeAlg_ip, ABTimeAlg_new, ABTimeAlg_init):  # src/advection.jl:545-578
        advd.t_bufc.insert(0, advd.bufcur.copy())
        bufc = _ab_sum(adv, advd.t_bufc, ordalg, ordalg)
        autointerp(advd.bufcur, bufc, ordalg - 1, interps, fl)
        bufc.free()
        advd.t_bufc.pop().free()
        interpbufc(advd.t_bufc, advd.bufcur, interps, None, fl)


_IDENTITY = None


def _identity_interp():
    """Lagrange(1): weights (1 - t, t) -- with a zero shift exactly (1, 0), i.e. the identity along that dim"""
    global _IDENTITY
    if _IDENTITY is None:
        from .interp import Lagrange

        _IDENTITY = Lagrange(1)
    return _IDENTITY


def advection_split_points(advd):
    """advection! for a SPLIT state whose shifts vary per point, e.g. [([1, 2], 1, 1, false), ([2, 1], 1, 2, false)] --
    the split form of the quasi-geostrophic driver (src/advection.jl:633-645 with the 4-argument getalpha,
    src/quasigeostrophic.jl:126-135): every line along dim perm[1] is interpolated with the shifts
    bufcur[ind][invp[1]] of its own points.  On the device this is the per-point 2-D kernel (slb_interp2d_points)
    with the swept dim's displacement plane and an exact identity (Lagrange(1) at zero shift: weights (1, 0)) along
    the other dim -- the same sum of the same products, so SLB_SWEEP_EXACT stays bit-identical to the 1-D stencil."""
    adv = advd.adv
    st = advd.getst()
    if adv.N != 2 or st.ndims != 1:
        raise NotImplementedError("split states with per-point shifts are on the B200 path for 2-D grids (SURVEY.md 8f)")
    if adv.timealg != NoTimeAlg:
        raise NotImplementedError("the Adams-Bashforth time algorithms drive the unsplit state only")
    L = _lib.lib()
    advd.flush()
    advd.parext.initcoef(advd)
    if advd.bufcur is None:
        raise RuntimeError("the provider's initcoef must set advd.bufcur for states with per-point shifts")
    n1, n2 = adv.sizeall
    d, comp = st.perm[0] - 1, st.invp[0] - 1
    pb = n1 * n2 * 8
    dec = DeviceField(advd.ctx, n1, n2, 2)
    src_plane = C.c_void_p(advd.bufcur.ptr.value + comp * pb)
    _lib.check(L.slb_memcpy_d2d(advd.ctx.h, C.c_void_p(dec.ptr.value + d * pb), src_plane, pb))
    zero = (C.c_double * 1)(0.0)
    _lib.check(L.slb_lincomb(advd.ctx.h, C.c_void_p(dec.ptr.value + (1 - d) * pb), 1, zero, (C.c_void_p * 1)(src_plane.value), n1 * n2))
    interps = list(adv.t_interp)
    interps[1 - d] = _identity_interp()
    back = DeviceField.view(advd.ctx, n1, n2, 1, C.c_void_p(L.slb_grid_back(advd.grid)))
    try:
        interpolate_points(back, advd.data_field(), dec, interps, advd.flags)
    finally:
        dec.free()
    _lib.check(L.slb_grid_swap(advd.grid))
    advd._linesum_dim = None
    return advd.nextstate()


def advection_single_state(advd):
    """advection! for an Advection with ONE state whose shifts vary per point, e.g.
    [([1, 2], 2, 1, false)] (src/advection.jl:594-619, :703-704; test/test_poisson2d.jl:197,
    test/test_swirling.jl:207): data <- interpolate(data, bufcur); nextstate!."""
    adv = advd.adv
    st = advd.getst()
    if len(adv.states) != 1 and st.ndims == 1:
        return advection_split_points(advd)
    if len(adv.states) != 1 or adv.N != 2 or st.ndims != 2 or st.perm != [1, 2]:
        raise NotImplementedError(
            "per-point shifts are on the B200 path for one unsplit 2-D state ([1, 2], 2, 1, false) (SURVEY.md 8f-1)")
    initcoef(advd)
    advd.interpolate_data()
    return advd.nextstate()
