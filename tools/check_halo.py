"""torchrun --nproc-per-node P tools/check_halo.py [n] [order] [steps]: the halo-sharded driver over P GPUs (one process
per GPU; torch.distributed only carries the handle bytes and the verdict) must reproduce the single-GPU driver, which
every rank recomputes on its own GPU."""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "semilagrangian.jl_b200"))
import torch
import torch.distributed as dist

import slb200 as S
from slb200.sharded import HaloShardedAdvectionData, torch_allgather_bytes

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo", rank=rank, world_size=world)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
order = int(sys.argv[2]) if len(sys.argv) > 2 else 7
nsteps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ms = (S.UniformMesh(0.0, 4 * math.pi, n), S.UniformMesh(0.0, 4 * math.pi, n), S.UniformMesh(-6.0, 6.0, n), S.UniformMesh(-6.0, 6.0, n))
tabst = [([3, 4, 1, 2], 1, 1, True), ([4, 3, 1, 2], 1, 1, True), ([1, 2, 4, 3], 1, 2, True), ([2, 1, 3, 4], 1, 2, True)]
ctx1 = S.Context(local)   # ONE context (stream) for the single-GPU reference run: its provider and its grid must share it
adv = S.Advection(ms, [S.Lagrange(order)] * 4, 0.1, tabst, ctx=ctx1)
fsp = lambda x: 0.5 * np.cos(x / 2) + 1
fv = lambda v: np.exp(-v**2 / 2) / math.sqrt(2 * math.pi)
f = S.dotprod((fsp(ms[0].points), fsp(ms[1].points), fv(ms[2].points), fv(ms[3].points)))
c = n // world
sh = HaloShardedAdvectionData(adv, np.asfortranarray(f[:, :, :, rank * c:(rank + 1) * c]), rank, world, torch_allgather_bytes(dist), device=local)
plain = S.AdvectionData(adv, f, S.getpoissonvar(adv, ctx=ctx1), ctx=ctx1)
worst = 0.0
for step in range(nsteps):
    while S.advection(plain):
        pass
    while sh.advection():
        pass
    ee_s, ee_p = sh.compute_ee(), S.compute_ee(plain)
    g = sh.getdata_local()
    p = plain.getdata()[:, :, :, rank * c:(rank + 1) * c]
    err = float(np.max(np.abs(g - p)) / np.max(np.abs(p)))
    worst = max(worst, err, abs(ee_s - ee_p) / abs(ee_p))
ok = worst <= 1e-12
print(f"rank {rank}/{world}: halo driver L{order} n={n} H={sh.H} c={sh.c} fused_passes={sh.n_fused} max rel err vs single-GPU driver = {worst:.3e} "
      f"ee={ee_s:.15e} {'OK' if ok else 'FAIL'}", flush=True)
dist.barrier()
sh.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
