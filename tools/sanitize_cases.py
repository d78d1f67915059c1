"""Small invocations of every hot kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
  python tools/sanitize_cases.py            (run under:  compute-sanitizer --tool <t> python tools/sanitize_cases.py)
Shapes are tiny: the tools slow kernels down 10-100x.  The halo-sharded passes run with ONE rank (its own
neighbour on both sides: windowed march, row / plane pushes, mailbox all-gather and barrier kernels all execute) --
the sanitizer serialises kernel launches, which ranks that wait for each other's flags cannot survive."""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "semilagrangian.jl_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import slb200 as S  # noqa: E402
from helpers import DeviceGrid  # noqa: E402
from slb200.sharded import local_group  # noqa: E402

rng = np.random.default_rng(3)


def vp(sz, interps, steps=1):
    ms = (S.UniformMesh(0.0, 4 * math.pi, sz[0]), S.UniformMesh(0.0, 4 * math.pi, sz[1]), S.UniformMesh(-6.0, 6.0, sz[2]), S.UniformMesh(-6.0, 6.0, sz[3]))
    tabst = [([3, 4, 1, 2], 1, 1, True), ([4, 3, 1, 2], 1, 1, True), ([1, 2, 4, 3], 1, 2, True), ([2, 1, 3, 4], 1, 2, True)]
    adv = S.Advection(ms, interps, 0.1, tabst)
    fsp = lambda x: 0.5 * np.cos(x / 2) + 1
    fv = lambda v: np.exp(-v**2 / 2) / math.sqrt(2 * math.pi)
    f = S.dotprod((fsp(ms[0].points), fsp(ms[1].points), fv(ms[2].points), fv(ms[3].points)))
    return adv, f


# 1. fused pair passes + single sweeps + field solve (FFT and DFT forms) + reductions: one Strang step, 16^4, Lagrange 7
adv, f = vp((16, 16, 16, 16), [S.Lagrange(7)] * 4)
a = S.AdvectionData(adv, f, S.getpoissonvar(adv))
while S.advection(a):
    pass
print("fused step ee", S.compute_ee(a), "ke", S.compute_ke(a))
a.fuse_pairs = False
while S.advection(a):
    pass
print("unfused step ee", S.compute_ee(a))
# 2. segmented B-spline sweeps (strided and dim 0), orders 3 and 11, n = 32; long-line kernel n = 512
for order in (3, 11):
    adv, f = vp((32, 8, 32, 8), [S.BSplineLU(order, n) for n in (32, 8, 32, 8)])
    g = DeviceGrid(rng.random((32, 6, 32, 4)))
    it = S.BSplineLU(order, 32)
    g.sweep(0, it, rng.uniform(-3, 3, 32), [0, 0, 1, 0])
    g.sweep(2, it, rng.uniform(-3, 3, 32 * 6), [1, 32, 0, 0])
    print("bspline seg order", order, float(np.sum(g.get())))
    g.close()
# 128-point lines: the (32 rows x 4 warps) variants with capped registers; dim 0: the tile shares the exchange buffers' memory
for order in (5, 11):
    it = S.BSplineLU(order, 128)
    g = DeviceGrid(rng.random((128, 40)))
    g.sweep(0, it, rng.uniform(-3, 3, 40), [0, 1])
    g.close()
    g = DeviceGrid(rng.random((40, 128)))
    g.sweep(1, it, rng.uniform(-3, 3, 40), [1, 0])
    print("bspline seg n=128 order", order, float(np.sum(g.get())))
    g.close()
g = DeviceGrid(rng.random((512, 8)))
g.sweep(0, S.BSplineLU(5, 512), rng.uniform(-3, 3, 8), [0, 1])
g.close()
g = DeviceGrid(rng.random((8, 512)))
g.sweep(1, S.BSplineLU(5, 512), rng.uniform(-3, 3, 8), [1, 0])
print("bspline wline", float(np.sum(g.get())))
g.close()
# 3. halo-sharded passes, one rank
adv, f = vp((32, 8, 16, 16), [S.Lagrange(7)] * 4)
(s,) = local_group(adv, f, 1)
while s.advection():
    pass
print("halo step ee", s.compute_ee(), float(np.sum(s.getdata_local())))
s.close()
# 4. small-grid paths: chunked strided sweep and a CUDA-graph replay (1D1V)
mx, mv = S.UniformMesh(0.0, 4 * math.pi, 64), S.UniformMesh(-6.0, 6.0, 64)
adv = S.Advection((mx, mv), [S.Lagrange(9)] * 2, 0.1, [([2, 1], 1, 1, True), ([1, 2], 1, 2, True)])
f = S.dotprod((1 + 0.001 * np.cos(0.5 * mx.points), np.exp(-mv.points**2 / 2) / math.sqrt(2 * math.pi)))
a = S.AdvectionData(adv, f, S.getpoissonvar(adv))
sg = S.StepGraph(a, 2)
sg.launch()
print("graph ee", sg.energies())
sg.close()
# 5. step program (one persistent cooperative kernel, own grid barrier): v - x - v order with line sums, after one real step
while S.advection(a):
    pass
sp = S.StepProgram(a, nsteps=2, repeat=2)
sp.launch()
print("program ee", sp.energies(), [k for k, _, _ in sp.profile()][:4])
sp.close()
# 6. tile-staged dim-0 Lagrange / Hermite sweeps: orders 3, 7, 11, line lengths 32 and 64, a line count that leaves a partial tile
for order in (3, 7, 11):
    for n0, rest in ((32, (5, 7)), (64, (3, 11))):
        g = DeviceGrid(rng.random((n0,) + rest))
        g.sweep(0, S.Lagrange(order), rng.uniform(-40, 40, rest[0]), [0, 1, 0])
        print("contig tile", order, n0, float(np.sum(g.get())))
        g.close()
print("sanitize cases done")
