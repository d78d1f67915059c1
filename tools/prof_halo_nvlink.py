"""NVLink traffic of the halo-pushing passes, measured by ncu on ONE process that drives two GPUs (two in-process
ranks, host synchronisation between the passes: a profiler serialises kernel launches, so the ranks must not wait for
each other on the device):
    ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,gpu__time_duration.sum,dram__bytes_write.sum -k regex:k_sweep_fused \
        --csv --log-file gpurun_out/halo_nvlink.csv python tools/prof_halo_nvlink.py [n] [P]
Expected per pushing pass and rank: 2 H planes out (n = 128, H = 4, P = 2: 8 x 16.8 MB = 134 MB)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "semilagrangian.jl_b200"))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import slb200 as S  # noqa: E402
from slb200.sharded import local_group  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
P = int(sys.argv[2]) if len(sys.argv) > 2 else 2
adv, vecs = bench.vp2d2v_setup(S, n, 7, "lagrange")
f = np.empty((n,) * 4, order="F")
bench.fill_product(f, vecs)
ranks = local_group(adv, f, P, devices=list(range(P)), max_shift=1.0, host_sync=True)
del f
E = np.linspace(-0.6, 0.6, n * n)
for s in ranks:
    s.has_field = True
    s.tabE = s.ctx.to_device(E)
    s.E_dev = [s.tabE, s.tabE]


def sync():
    for s in ranks:
        s.ctx.sync()


dt = adv.dt_base
for rep in range(2):
    for s in ranks:
        s._pass(2, dt / 2, 3, dt / 2, 0)      # v1 v2, no pushes
    sync()
    for s in ranks:
        s._pass(0, dt, 1, dt, 2)              # x1 x2 with plane pushes
    sync()
    for s in ranks:
        s._pass(2, dt / 2, 3, dt / 2, 2)      # v1 v2 with row pushes
    sync()
for s in ranks:
    s.check()
print("done; H =", ranks[0].H, "c =", ranks[0].c)
