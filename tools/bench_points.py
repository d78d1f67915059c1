"""Times the per-point 2-D interpolation kernel (slb_interp2d_points) with CUDA
events on the library's stream, on grids larger than L2, and one unsplit Vlasov-Poisson step.
Algorithmic traffic: read f (8 B) + the two displacement planes (16 B) + write f (8 B) = 32 B per
point."""
import argparse
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "semilagrangian.jl_b200"))
import slb200 as S  # noqa: E402
from slb200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()
ctx = S.default_context()
n = a.n
x = np.arange(n) / n
dec = np.empty((n, n, 2), order="F")
dec[:, :, 0] = 3.7 * (-np.cos(math.pi * x[:, None]) ** 2 * np.sin(2 * math.pi * x[None, :]))
dec[:, :, 1] = 3.7 * (np.cos(math.pi * x[None, :]) ** 2 * np.sin(2 * math.pi * x[:, None]))
f = np.asfortranarray(np.exp(-10 * ((0.25 - x[:, None]) ** 2 + (0.5 - x[None, :]) ** 2)))
src, dfl = S.DeviceField.from_host(ctx, f), S.DeviceField.from_host(ctx, dec)
dst = src.like()
e0, e1 = ctx.event(), ctx.event()
res = {"n": n, "bytes_per_point": 32, "kernels": {}}
for name, its in (("lagrange5", [S.Lagrange(5)] * 2), ("lagrange9", [S.Lagrange(9)] * 2), ("hermite9", [S.Hermite(9)] * 2),
                  ("bspline_lu5", [S.BSplineLU(5, n), S.BSplineLU(5, n)])):
    for _ in range(3):
        S.interpolate_points(dst, src, dfl, its)
    ctx.sync()
    ctx.record(e0)
    for _ in range(a.reps):
        S.interpolate_points(dst, src, dfl, its)
    ctx.record(e1)
    ms = ctx.elapsed_ms(e0, e1) / a.reps
    res["kernels"][name] = {"ms": ms, "Gpoint_s": n * n / ms / 1e6, "GBps": 32 * n * n / ms / 1e6}

# one unsplit 1D1V Vlasov-Poisson run (StdPoisson2d, ABTimeAlg_ip order 3), 1024 x 1024, Lagrange 7
m = 1024
mesh_sp, mesh_v = S.UniformMesh(0.0, 4 * math.pi, m), S.UniformMesh(-9.0, 9.0, m)
dt = 0.01
adv = S.Advection((mesh_sp, mesh_v), [S.Lagrange(7)] * 2, dt, [([1, 2], 2, 1, False)], tab_coef=S.nosplit(dt), timealg=S.ABTimeAlg_ip, ordalg=3)
xx, yy = mesh_sp.points[:, None], mesh_v.points[None, :]
advd = S.AdvectionData(adv, 1 / math.sqrt(2 * math.pi) * np.exp(-0.5 * yy**2) * (1 + 0.5 * np.cos(xx / 2)), S.getpoissonvar(adv, type=S.StdPoisson2d))
for _ in range(3):
    while S.advection(advd):
        pass
ctx.sync()
l0 = ctx.launch_count()
ctx.record(e0)
nst = 20
for _ in range(nst):
    while S.advection(advd):
        pass
ctx.record(e1)
ms = ctx.elapsed_ms(e0, e1) / nst
res["vp1d1v_unsplit_ab3_1024"] = {"ms_per_step": ms, "launches_per_step": (ctx.launch_count() - l0) / nst, "Mpoint_steps_s": m * m / ms / 1e3,
                                  "energy": S.getenergy(advd)[2]}
print(json.dumps(res))
