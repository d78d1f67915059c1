"""torchrun --nproc-per-node P tools/check_sharded.py : the sharded driver over P GPUs must
reproduce the single-GPU driver (every rank recomputes the plain run on its own GPU)."""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "semilagrangian.jl_b200"))
import torch
import torch.distributed as dist

import slb200 as S
from slb200.distributed import ShardedAdvectionData, slab

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
exchange = sys.argv[2] if len(sys.argv) > 2 else "p2p"
kind = sys.argv[3] if len(sys.argv) > 3 else "lagrange"
order = int(sys.argv[4]) if len(sys.argv) > 4 else 7
mk = {"lagrange": lambda k: S.Lagrange(order), "bspline_lu": lambda k: S.BSplineLU(order, k),
      "bspline_fft": lambda k: S.BSplineFFT(order, k), "hermite": lambda k: S.Hermite(order)}[kind]
ms = (S.UniformMesh(0.0, 4 * math.pi, n), S.UniformMesh(0.0, 4 * math.pi, n), S.UniformMesh(-6.0, 6.0, n), S.UniformMesh(-6.0, 6.0, n))
tabst = [([3, 4, 1, 2], 1, 1, True), ([4, 3, 1, 2], 1, 1, True), ([1, 2, 4, 3], 1, 2, True), ([2, 1, 3, 4], 1, 2, True)]
adv = S.Advection(ms, [mk(n) for _ in range(4)], 0.1, tabst)
fsp = lambda x: 0.5 * np.cos(x / 2) + 1
fv = lambda v: np.exp(-v**2 / 2) / math.sqrt(2 * math.pi)
f = S.dotprod((fsp(ms[0].points), fsp(ms[1].points), fv(ms[2].points), fv(ms[3].points)))
lo, hi = slab(n, world, rank)
sh = ShardedAdvectionData(adv, np.asfortranarray(f[:, lo:hi, :, :]), exchange=exchange)
plain = S.AdvectionData(adv, f, S.getpoissonvar(adv), ctx=S.Context(local))
worst = 0.0
for step in range(3):
    while S.advection(plain):
        pass
    while sh.advection():
        pass
    ee_s, ee_p = sh.compute_ee(), S.compute_ee(plain)
    g = sh.gather_global()
    p = plain.getdata()
    err = float(np.max(np.abs(g - p)) / np.max(np.abs(p)))
    worst = max(worst, err, abs(ee_s - ee_p) / abs(ee_p))
ok = worst <= 1e-12
print(f"rank {rank}/{world}: {kind} {order} exchange={sh.exchange} fused_passes={sh.n_fused} barriers={sh.n_barriers} n={n} max rel err vs single-GPU driver = {worst:.3e} exchanges={sh.n_exchanges} {'OK' if ok else 'FAIL'}", flush=True)
sh.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
