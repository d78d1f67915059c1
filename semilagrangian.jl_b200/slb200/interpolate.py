"""Kernel seam: interpolate!(fp, fi, dec, interp) on host vectors / batches of lines.

Reference: src/interpolation.jl:175-193 (explicit decint/precal) and :302-315 (dec).
Runs on the device through the same slb_sweep as the driver (a [n, nlines] grid swept
along dim 0); host buffers in, host buffers out.
"""
import ctypes as C

import numpy as np

from . import _lib


def interpolate_lines(fi, dec, interp, ctx=None, flags=0, axis=0):
    """Shift every line of the 2-D array `fi` along `axis` by dec (scalar or one value per
    line, in grid units); returns the interpolated array."""
    ctx = ctx or _lib.default_context()
    fi = np.asfortranarray(fi, dtype=np.float64)
    if fi.ndim == 1:
        fi = fi.reshape(-1, 1, order="F")
        axis = 0
    other = 1 - axis
    nl = fi.shape[other]
    dec = np.ascontiguousarray(np.broadcast_to(np.asarray(dec, dtype=np.float64), (nl,)))
    L = _lib.lib()
    g = C.c_void_p()
    _lib.check(L.slb_grid_create(ctx.h, 2, _lib.i64(fi.shape), C.byref(g)))
    try:
        _lib.check(L.slb_grid_upload(g, fi.ctypes.data_as(C.c_void_p)))
        strides = [0, 0]
        strides[other] = 1
        h = interp.handle(ctx, fi.shape[axis])
        if getattr(interp, "edge", 1) == 2:  # InsideEdge: src/interpolation.jl:308-314
            flags = int(flags) | _lib.SLB_SWEEP_INSIDE_EDGE
        _lib.check(L.slb_sweep(g, axis, h, dec.ctypes.data_as(C.c_void_p), nl, _lib.i64(strides), 1.0, 0, int(flags)))
        out = np.empty(fi.shape, dtype=np.float64, order="F")
        _lib.check(L.slb_grid_download(g, out.ctypes.data_as(C.c_void_p)))
    finally:
        L.slb_grid_destroy(g)
    return out


def interpolate(fp, fi, dec, interp, ctx=None, flags=0):
    """interpolate!(fp, fi, dec, interp) -- src/interpolation.jl:302-315 (CircEdge)."""
    fi = np.asarray(fi, dtype=np.float64)
    if fp is fi:
        raise ValueError("fp and fi must not alias")
    fp[...] = interpolate_lines(fi, dec, interp, ctx=ctx, flags=flags)[:, 0]
    return None


def interpolate_nd(fp, fi, dec, interp_t, ctx=None, flags=0):
    """interpolate!(fp, fi, dec, interp_t) for N = 2 with per-point shifts -- src/interpolation.jl:401-429
    (dec a function of the 0-based index tuple) and :561-621 (dec an [n1, n2, 2] array of shifts in
    grid units, the OpTuple field).  fi: [n1, n2] or [n1, n2, ncomp] host array; fp receives the
    result.  Host buffers in and out; the arithmetic runs in slb_interp2d_points."""
    from .unsplit2d import DeviceField, interpolate_points

    ctx = ctx or _lib.default_context()
    fi = np.asarray(fi, dtype=np.float64)
    if fp is fi:
        raise ValueError("fp and fi must not alias")
    if len(interp_t) != fi.ndim and not (fi.ndim == 3 and len(interp_t) == 2):
        raise ValueError(f"The number of Interpolation {len(interp_t)} is different of N={fi.ndim}")
    n1, n2 = fi.shape[:2]
    if callable(dec):
        d = np.empty((n1, n2, 2), order="F")
        for j in range(n2):
            for i in range(n1):
                d[i, j, :] = dec((i, j))
        dec = d
    dec = np.asarray(dec, dtype=np.float64)
    if dec.shape != (n1, n2, 2):
        raise ValueError(f"dec must have shape {(n1, n2, 2)}")
    src = DeviceField.from_host(ctx, fi)
    dfl = DeviceField.from_host(ctx, dec)
    dst = src.like()
    try:
        interpolate_points(dst, src, dfl, interp_t, flags)
        fp[...] = dst.to_host()
    finally:
        for f in (src, dfl, dst):
            f.free()
    return None


def sol(interp, b, ctx=None):
    """sol(interp, b) -- src/interpolation.jl:40, src/bsplinelu.jl:282-284, src/bsplinefft.jl:49-51;
    b: vector or [n, nlines] array (lines along axis 0)."""
    ctx = ctx or _lib.default_context()
    b = np.asfortranarray(b, dtype=np.float64)
    one = b.ndim == 1
    if one:
        b = b.reshape(-1, 1, order="F")
    L = _lib.lib()
    g = C.c_void_p()
    _lib.check(L.slb_grid_create(ctx.h, 2, _lib.i64(b.shape), C.byref(g)))
    try:
        _lib.check(L.slb_grid_upload(g, b.ctypes.data_as(C.c_void_p)))
        _lib.check(L.slb_presolve(g, 0, interp.handle(ctx, b.shape[0])))
        out = np.empty(b.shape, dtype=np.float64, order="F")
        _lib.check(L.slb_grid_download(g, out.ctypes.data_as(C.c_void_p)))
    finally:
        L.slb_grid_destroy(g)
    return out[:, 0] if one else out
