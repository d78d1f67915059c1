"""Shared helpers for the parity tests: run the ORACLE on the same inputs as the CUDA path."""
import numpy as np

from oracle import clib, refmodel as R

KINDS = {"lagrange": 0, "bspline_lu": 1, "bspline_fft": 2, "hermite": 3}


def make_pair(kind, order, n):
    """(product interp, oracle interp) of the same type."""
    import slb200 as S

    if kind == "lagrange":
        return S.Lagrange(order), R.Lagrange(order)
    if kind == "bspline_lu":
        return S.BSplineLU(order, n), R.BSplineLU(order, n)
    if kind == "bspline_fft":
        return S.BSplineFFT(order, n), R.BSplineFFT(order, n)
    if kind == "hermite":
        return S.Hermite(order), R.Hermite(order)
    raise ValueError(kind)


def oracle_sweep(f, dim, ointerp, tab, astride, nthreads=4):
    """orc_sweep on a copy of f (Fortran order); tab is the pre-scaled alpha table."""
    g = np.array(f, dtype=np.float64, order="F", copy=True)
    scratch = np.empty(g.size)
    tab = np.ascontiguousarray(tab, dtype=np.float64)
    rc = clib.lib().orc_sweep(
        g.ctypes.data_as(clib.c_double_p), clib.dp(scratch), g.ndim, clib.lp(g.shape), dim, ointerp._h,
        clib.dp(tab), clib.lp(astride), nthreads,
    )
    assert rc == 0
    return g


def relerr(a, b):
    """max_i |a_i - b_i| / max_i |b_i|  (SURVEY.md 8d parity figure)"""
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


class DeviceGrid:
    """Thin test wrapper: an N-D grid on the device, swept through the C ABI."""

    def __init__(self, f, ctx=None):
        import ctypes as C
        from slb200 import _lib

        self._lib, self.C = _lib, C
        self.ctx = ctx or _lib.default_context()
        f = np.asfortranarray(f, dtype=np.float64)
        self.shape = f.shape
        self.h = C.c_void_p()
        _lib.check(_lib.lib().slb_grid_create(self.ctx.h, f.ndim, _lib.i64(f.shape), C.byref(self.h)))
        _lib.check(_lib.lib().slb_grid_upload(self.h, f.ctypes.data_as(C.c_void_p)))

    def sweep(self, dim, interp, tab, astride, scale=1.0, flags=0):
        _lib, C = self._lib, self.C
        tab = np.ascontiguousarray(tab, dtype=np.float64)
        h = interp.handle(self.ctx, self.shape[dim])
        _lib.check(_lib.lib().slb_sweep(self.h, dim, h, tab.ctypes.data_as(C.c_void_p), tab.size, _lib.i64(astride), float(scale), 0, flags))

    def sweep_pair(self, dimA, interpA, tabA, astrA, dimB, interpB, tabB, astrB, flags=0):
        """slb_sweep_pair with host alpha tables"""
        _lib, C = self._lib, self.C
        tabA = np.ascontiguousarray(tabA, dtype=np.float64)
        tabB = np.ascontiguousarray(tabB, dtype=np.float64)
        hA = interpA.handle(self.ctx, self.shape[dimA])
        hB = interpB.handle(self.ctx, self.shape[dimB])
        _lib.check(_lib.lib().slb_sweep_pair(
            self.h, dimA, hA, tabA.ctypes.data_as(C.c_void_p), tabA.size, _lib.i64(astrA), 1.0,
            dimB, hB, tabB.ctypes.data_as(C.c_void_p), tabB.size, _lib.i64(astrB), 1.0, 0, flags))

    def get(self):
        out = np.empty(self.shape, dtype=np.float64, order="F")
        self._lib.check(self._lib.lib().slb_grid_download(self.h, out.ctypes.data_as(self.C.c_void_p)))
        return out

    def close(self):
        if self.h:
            self._lib.lib().slb_grid_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
