"""Profiling driver: one warm-up Strang step + N profiled steps of the 2D2V n^4 workload.
Run under ncu on the GPU box (see profiles/README.md); numbers printed under a profiler are
never bench values."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "semilagrangian.jl_b200"))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import slb200 as S  # noqa: E402
from slb200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=128)
ap.add_argument("--order", type=int, default=7)
ap.add_argument("--interp", default="lagrange")
ap.add_argument("--steps", type=int, default=1)
a = ap.parse_args()
ctx = S.default_context()
adv, vecs = bench.vp2d2v_setup(S, a.size, a.order, a.interp)
host, _ = _lib.pinned_empty((a.size,) * 4)
bench.fill_product(host, vecs)
advd = S.AdvectionData(adv, host, S.getpoissonvar(adv))
for _ in range(1 + a.steps):
    while S.advection(advd):
        pass
print("ee", S.compute_ee(advd), "launches", ctx.launch_count())
