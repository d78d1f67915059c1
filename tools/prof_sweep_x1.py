"""A few stand-alone sweeps along dim 0 of a 128^4 grid (Lagrange order from argv) for ncu: python tools/prof_sweep_x1.py [order]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "semilagrangian.jl_b200"))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import slb200 as S  # noqa: E402

order = int(sys.argv[1]) if len(sys.argv) > 1 else 7
n = 128
adv, vecs = bench.vp2d2v_setup(S, n, order, "lagrange")
f = np.empty((n,) * 4, order="F")
bench.fill_product(f, vecs)
g = S.AdvectionData(adv, f, S.getpoissonvar(adv))
tab = (g.points_dev(2), n)
for _ in range(4):
    S.sweep(g, 0, adv.t_interp[0], tab, [0, 0, 1, 0], -0.1 / adv.t_mesh[0].step, True)
g.ctx.sync()
print("ok")
