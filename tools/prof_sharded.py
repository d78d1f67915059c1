"""torchrun --nproc-per-node P tools/prof_sharded.py : where does a sharded Strang step spend its time?
CUDA events around the field solve (rho reduction + all-gather + Poisson), the exchange passes
(including their barriers) and the local pass, max over ranks."""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "semilagrangian.jl_b200"))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import bench
import slb200 as S
from slb200.distributed import ShardedAdvectionData, slab

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
adv, vecs = bench.vp2d2v_setup(S, n, 7, "lagrange")
lo, hi = slab(n, world, rank)
a, b, c, d = vecs
loc = np.empty((n, hi - lo, n, n), order="F")
bench.fill_product(loc, (a, b[lo:hi], c, d))
sh = ShardedAdvectionData(adv, loc)
for _ in range(3):
    while sh.advection():
        pass
torch.cuda.synchronize()
dist.barrier()
names, evs = [], []


def mark(name):
    e = torch.cuda.Event(enable_timing=True)
    e.record(sh.stream)
    names.append(name)
    evs.append(e)


orig_field, orig_pair, orig_barrier = sh._compute_field, sh._pair, sh._barrier


def field():
    mark("field<")
    orig_field()
    mark("field>")


def pair(*a_):
    mark("pair<")
    orig_pair(*a_)
    mark("pair>")


def barrier():
    mark("barrier<")
    orig_barrier()
    mark("barrier>")


sh._compute_field, sh._pair, sh._barrier = field, pair, barrier
steps = 5
with torch.cuda.stream(sh.stream):
    mark("start")
    for _ in range(steps):
        while sh.advection():
            pass
    mark("end")
torch.cuda.synchronize()
tot = {}
stack = []
for i, nm in enumerate(names):
    if nm.endswith("<"):
        stack.append((nm[:-1], evs[i]))
    elif nm.endswith(">"):
        k, e0 = stack.pop()
        tot[k] = tot.get(k, 0.0) + e0.elapsed_time(evs[i])
total = evs[0].elapsed_time(evs[-1])
vals = torch.tensor([total] + [tot.get(k, 0.0) for k in ("field", "pair", "barrier")], dtype=torch.float64, device="cuda")
dist.all_reduce(vals, op=dist.ReduceOp.MAX)
if rank == 0:
    t, f, p, bq = [float(x) / steps for x in vals]
    print(f"P={world} n={n}: step {t:.3f} ms | field solves {f:.3f} | fused passes incl. their barriers {p:.3f} (barriers alone {bq:.3f}) | rest {t - f - p:.3f}")
sh.close()
dist.destroy_process_group()
