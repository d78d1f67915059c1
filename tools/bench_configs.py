"""Times the small BASELINE.json configurations on one GPU (CUDA events on the library's stream):
  C1  Vlasov-Poisson 1D1V Landau damping 128 x 256, Lagrange 9, Strang splitting
  C2  2-D rigid rotation 1024 x 1024, periodic B-spline order 5 (BSplineLU), magic splitting
  C3  Vlasov-Poisson 2D2V 64^4, Lagrange 7, Strang splitting
These grids are cache-resident or launch-bound (C1: 256 KB, C2: 8 MB, C3: 134 MB); the headline
metric is measured by bench.py on 128^4.  One JSON line."""
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "semilagrangian.jl_b200"))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import slb200 as S  # noqa: E402

ctx = S.default_context()
e0, e1 = ctx.event(), ctx.event()


def timed(advd, nsteps, warm=5):
    for _ in range(warm):
        while S.advection(advd):
            pass
    ctx.sync()
    l0 = ctx.launch_count()
    ctx.record(e0)
    for _ in range(nsteps):
        while S.advection(advd):
            pass
    ctx.record(e1)
    ms = ctx.elapsed_ms(e0, e1) / nsteps
    return ms, (ctx.launch_count() - l0) / nsteps


res = {}
# C1: examples/vlasov-poisson-1d1v.jl:24-45
nx, nv = 128, 256
mx, mv = S.UniformMesh(0.0, 2 * math.pi / 0.5, nx), S.UniformMesh(-6.0, 6.0, nv)
adv = S.Advection((mx, mv), [S.Lagrange(9)] * 2, 0.1, [([2, 1], 1, 1, True), ([1, 2], 1, 2, True)])
f = S.dotprod((1 + 0.001 * np.cos(0.5 * mx.points), np.exp(-mv.points**2 / 2) / math.sqrt(2 * math.pi)))
advd = S.AdvectionData(adv, f, S.getpoissonvar(adv))
ms, nl = timed(advd, 200)
res["C1_vp1d1v_128x256_L9_strang"] = {"ms_per_step": ms, "launches_per_step": nl, "sweeps_per_step": 3,
                                      "Gcell_s_per_sweep": 3 * nx * nv / ms / 1e6, "note": "256 KB grid: launch-bound"}
# C2: examples/run_rotation.jl shape, test/test_rotation.jl:43-62
n = 1024
m1, m2 = S.UniformMesh(-5.0, 5.0, n), S.UniformMesh(-5.0, 5.0, n)
dt = 2 * math.pi / 100
adv = S.Advection((m1, m2), [S.BSplineLU(5, n), S.BSplineLU(5, n)], dt, [([1, 2], 1, 1, True), ([2, 1], 1, 2, True)], tab_coef=S.magicsplit(dt))
x, y = m1.points[:, None], m2.points[None, :]
advd = S.AdvectionData(adv, np.exp(-13 * (x**2 + (y + 1.2) ** 2)), S.getrotationvar(adv))
ms, nl = timed(advd, 50)
res["C2_rotation_1024x1024_bsplinelu5_magic"] = {"ms_per_step": ms, "launches_per_step": nl, "sweeps_per_step": 3,
                                                 "Gcell_s_per_sweep": 3 * n * n / ms / 1e6, "note": "8 MB grid: L2-resident"}
# C3
adv, vecs = bench.vp2d2v_setup(S, 64, 7, "lagrange")
f = np.empty((64,) * 4, order="F")
bench.fill_product(f, vecs)
advd = S.AdvectionData(adv, f, S.getpoissonvar(adv))
ms, nl = timed(advd, 50)
res["C3_vp2d2v_64^4_L7_strang"] = {"ms_per_step": ms, "launches_per_step": nl, "sweeps_per_step": 6,
                                   "Gcell_s_per_sweep": 6 * 64**4 / ms / 1e6, "note": "134 MB grid: just above L2"}
print(json.dumps(res))
