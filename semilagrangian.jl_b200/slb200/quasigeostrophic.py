"""Surface quasi-geostrophic displacement provider -- host mirror of src/quasigeostrophic.jl:1-121
(GeoConst, GeoVar, getgeovar, initdata!, initcoef!) for the unsplit 2-D state ([1, 2], 2, 1, false).

The velocity is a spectral multiplier of the advected buoyancy itself,
    u_x = real(ifft(coefrsqk[x] .* fft(b))),  coefrsqk[1] = i k_y / |k|,  coefrsqk[2] = -i k_x / |k|,
the same shape as the Poisson field solve (purely imaginary multipliers, real output), so it runs on the
library's DFT kernels (slb_poisson_create / slb_poisson_solve) straight from the device-resident data;
bufcur = dt * (u_1, u_2) is formed by slb_lincomb.  Nothing leaves the device.
"""
import ctypes as C

import numpy as np

from . import _lib
from .advection import AbstractExtDataAdv
from .mesh import start, stop, vec_k_fft
from .unsplit2d import DeviceField


class GeoVar(AbstractExtDataAdv):
    def __init__(self, adv, ctx=None, odg_b=1e-3):
        if adv.N != 2:
            raise ValueError("the number of dimension must be 2")  # src/quasigeostrophic.jl:26
        self.adv = adv
        self.ctx = ctx or adv.ctx or _lib.default_context()
        self.odg_b = float(odg_b)
        n1, n2 = adv.sizeall
        kx = vec_k_fft(adv.t_mesh[0]).reshape(n1, 1)
        ky = vec_k_fft(adv.t_mesh[1]).reshape(1, n2)
        k = np.sqrt(kx * kx + ky * ky)
        with np.errstate(divide="ignore"):
            v = np.where(k == 0, 0.0, 1.0 / k)
        # imaginary parts of coefrsqk (src/quasigeostrophic.jl:34-42)
        self.coef_imag = (np.asfortranarray(v * ky), np.asfortranarray(-(v * kx)))
        self._keep = [np.ascontiguousarray(a.reshape(-1, order="F")) for a in self.coef_imag]
        arr = (_lib.c_double_p * 2)(*[_lib.dptr(a) for a in self._keep])
        h = C.c_void_p()
        _lib.check(_lib.lib().slb_poisson_create(self.ctx.h, 2, _lib.i64((n1, n2)), arr, C.byref(h)))
        self.plan = h
        self.u_dev = [self.ctx.malloc(n1 * n2 * 8) for _ in range(2)]

    def initdata(self, advd):
        """initdata!(geoc, advd) -- src/quasigeostrophic.jl:79-104: two warm anticyclones, two cold cyclones"""
        mx, my = advd.adv.t_mesh
        lx, ly = stop(mx) - start(mx), stop(my) - start(my)  # src/quasigeostrophic.jl:82-83
        x, y = mx.points, my.points
        ee, sig = 4, lx / 15

        def anticyclone(cx, cy):
            return np.exp(-ee * (x - cx) ** 2 / (2 * sig**2))[:, None] * np.exp(-((y - cy) ** 2) / (2 * sig**2))[None, :]

        d = anticyclone(lx / 4, ly / 4)
        d = d + anticyclone(3 * lx / 4, ly / 4)
        d = d - anticyclone(lx / 4, 3 * ly / 4)
        d = d - anticyclone(3 * lx / 4, 3 * ly / 4)
        advd.upload(d * self.odg_b)

    def initcoef(self, advd):
        """initcoef!(geoc, advd) -- src/quasigeostrophic.jl:109-121"""
        L = _lib.lib()
        n1, n2 = advd.adv.sizeall
        dt = advd.getcur_t()
        advd.flush()
        arr = (C.c_void_p * 2)(*[p.value for p in self.u_dev])
        _lib.check(L.slb_poisson_solve(self.plan, C.c_void_p(L.slb_grid_front(advd.grid)), arr))
        if advd.bufcur is None:
            advd.bufcur = DeviceField(self.ctx, n1, n2, 2)
        coef = (C.c_double * 1)(float(dt))
        for x in range(2):
            src = (C.c_void_p * 1)(self.u_dev[x].value)
            plane = C.c_void_p(advd.bufcur.ptr.value + x * n1 * n2 * 8)
            _lib.check(L.slb_lincomb(self.ctx.h, plane, 1, coef, src, n1 * n2))

    def initcoef_reads_data(self, advd):
        return True

    def close(self):
        if getattr(self, "plan", None):
            _lib.lib().slb_poisson_destroy(self.plan)
            self.plan = None
            for p in self.u_dev:
                self.ctx.free(p)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def getgeovar(adv, ctx=None, **kw):
    """getgeovar(adv; kwargs...) -- src/quasigeostrophic.jl:72-74"""
    return GeoVar(adv, ctx=ctx, **kw)
