"""Times slb_sweep_pair against two slb_sweep calls on the 2D2V n^4 grid (CUDA events).
Knobs: SLB_PAIR_CHUNK_MB, SLB_PAIR_RING (read by the library)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "semilagrangian.jl_b200"))
import ctypes as C

import slb200 as S
from slb200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=128)
ap.add_argument("--order", type=int, default=7)
ap.add_argument("--dims", default="2,3")
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
n = a.size
dA, dB = [int(x) for x in a.dims.split(",")]
ctx = S.default_context()
L = _lib.lib()
h = C.c_void_p()
_lib.check(L.slb_grid_create(ctx.h, 4, _lib.i64((n,) * 4), C.byref(h)))
host = np.random.default_rng(1).random(n**4)
_lib.check(L.slb_grid_upload(h, host.ctypes.data_as(C.c_void_p)))
it = S.Lagrange(a.order).handle(ctx, n)
if dA >= 2:   # velocity sweeps: alpha = (dt/dv) E(x1, x2)
    tab = ctx.to_device(np.linspace(-0.4, 0.4, n * n)); tlen = n * n; strides = [1, n, 0, 0]
else:         # space sweeps: alpha = -(dt/dx) v
    tab = ctx.to_device(np.linspace(-6, 6, n)); tlen = n; strides = [0, 0, 0, 1] if 3 not in (dA, dB) else [0, 0, 1, 0]
e0, e1 = ctx.event(), ctx.event()


def timed(fn):
    fn(); ctx.sync()
    ctx.record(e0)
    for _ in range(a.reps):
        fn()
    ctx.record(e1)
    return _lib.Context.elapsed_ms(e0, e1) / a.reps


def two():
    _lib.check(L.slb_sweep(h, dA, it, tab, tlen, _lib.i64(strides), 1.0, 1, 0))
    _lib.check(L.slb_sweep(h, dB, it, tab, tlen, _lib.i64(strides), 1.0, 1, 0))


ms2 = timed(two)
print(f"two sweeps dims {dA},{dB}: {ms2:.3f} ms  ({2*n**4/ms2/1e6:.1f} Gcell/s)")


def pair():
    _lib.check(L.slb_sweep_pair(h, dA, it, tab, tlen, _lib.i64(strides), 1.0, dB, it, tab, tlen, _lib.i64(strides), 1.0, 1, 0))


ms = timed(pair)
print(f"fused pair dims {dA},{dB}: {ms:.3f} ms  ({2*n**4/ms/1e6:.1f} Gcell/s, {n**4*16/ms/1e6:.0f} GB/s of HBM traffic)  speed-up {ms2/ms:.2f}")
