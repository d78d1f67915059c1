"""ORACLE (test infrastructure) -- ctypes binding of oracle/liboracle.so (oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

c_long_p = C.POINTER(C.c_long)
c_double_p = C.POINTER(C.c_double)


def build(force=False):
    """Compile oracle.c -> liboracle.so (gcc).  Building the checker is not using it."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        build()
    L = C.CDLL(_SO)
    L.orc_getprecal.argtypes = [c_double_p, C.c_int, C.c_int, C.c_double, c_double_p]
    L.orc_getprecal.restype = None
    L.orc_split_alpha.argtypes = [C.c_double, c_double_p]
    L.orc_split_alpha.restype = C.c_long
    L.orc_interpolate_circ.argtypes = [c_double_p, c_double_p, C.c_long, C.c_long, c_double_p, C.c_int]
    L.orc_interpolate_circ.restype = None
    L.orc_lu_create.argtypes = [C.c_long, c_double_p, C.c_int, C.c_int]
    L.orc_lu_create.restype = C.c_void_p
    L.orc_lu_destroy.argtypes = [C.c_void_p]
    L.orc_lu_dims.argtypes = [C.c_void_p, c_long_p]
    L.orc_lu_band.argtypes = [C.c_void_p]
    L.orc_lu_band.restype = c_double_p
    L.orc_lu_lastrows.argtypes = [C.c_void_p]
    L.orc_lu_lastrows.restype = c_double_p
    L.orc_lu_lastcols.argtypes = [C.c_void_p]
    L.orc_lu_lastcols.restype = c_double_p
    L.orc_lu_sol.argtypes = [C.c_void_p, c_double_p, c_double_p]
    L.orc_lu_sol.restype = None
    L.orc_fft.argtypes = [c_double_p, c_double_p, C.c_long, C.c_int]
    L.orc_fft.restype = C.c_int
    L.orc_interp_create.argtypes = [C.c_int, C.c_int, C.c_long, c_double_p, C.c_int, c_double_p]
    L.orc_interp_create.restype = C.c_void_p
    L.orc_interp_destroy.argtypes = [C.c_void_p]
    L.orc_interpolate_alpha.argtypes = [C.c_void_p, c_double_p, c_double_p, C.c_long, C.c_double]
    L.orc_interpolate_alpha.restype = None
    L.orc_sol_line.argtypes = [C.c_void_p, c_double_p, c_double_p]
    L.orc_sol_line.restype = None
    L.orc_interpolate_points2d.argtypes = [C.c_void_p, C.c_void_p, c_double_p, c_double_p, c_double_p, C.c_long, C.c_long, C.c_int, C.c_int]
    L.orc_interpolate_points2d.restype = None
    L.orc_interpolate_inside.argtypes = [c_double_p, c_double_p, C.c_long, C.c_long, C.c_double, c_double_p, C.c_int, C.c_int]
    L.orc_interpolate_inside.restype = C.c_int
    L.orc_sweep.argtypes = [c_double_p, c_double_p, C.c_int, c_long_p, C.c_int, C.c_void_p, c_double_p, c_long_p, C.c_int]
    L.orc_sweep.restype = C.c_int
    L.orc_compute_charge.argtypes = [c_double_p, c_double_p, C.c_long, C.c_long, C.c_double, C.c_int]
    L.orc_compute_charge.restype = None
    L.orc_max_threads.restype = C.c_int
    _lib = L
    return L


def dp(a):
    assert a.dtype == np.float64
    return a.ctypes.data_as(c_double_p)


def lp(seq):
    arr = (C.c_long * len(seq))(*[int(x) for x in seq])
    return arr
