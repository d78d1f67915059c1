"""PCIe probe for the e2e leg: pinned H2D / D2H bandwidth alone and concurrently (two streams), with the process's
default CPU affinity and pinned to each NUMA node in turn (pinned host memory is placed by first touch)."""
import glob
import os
import subprocess
import sys
import time

import torch

print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
nodes = sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))
print("numa nodes:", [os.path.basename(n) for n in nodes], "affinity:", len(os.sched_getaffinity(0)), "cpus")


def cpus_of(node):
    out = []
    for part in open(os.path.join(node, "cpulist")).read().strip().split(","):
        if "-" in part:
            a, b = part.split("-")
            out += list(range(int(a), int(b) + 1))
        elif part:
            out.append(int(part))
    return out


def probe(tag, nbytes=1 << 31):
    h1 = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h2 = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h1.fill_(1)
    h2.fill_(2)
    d1 = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    d2 = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    def h2d():
        with torch.cuda.stream(s1):
            d1.copy_(h1, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)

    def both():
        h2d()
        d2h()

    def both_chunked(nch=8):
        c = nbytes // nch
        for k in range(nch):
            with torch.cuda.stream(s1):
                d1[k * c:(k + 1) * c].copy_(h1[k * c:(k + 1) * c], non_blocking=True)
            with torch.cuda.stream(s2):
                h2[k * c:(k + 1) * c].copy_(d2[k * c:(k + 1) * c], non_blocking=True)

    gb = nbytes / 1e9
    print(f"{tag}: H2D {gb / timed(h2d):.1f} GB/s, D2H {gb / timed(d2h):.1f} GB/s, both {gb / timed(both):.1f} GB/s each way, "
          f"both in 8 chunks {gb / timed(both_chunked):.1f}", flush=True)
    del h1, h2, d1, d2


all_cpus = sorted(os.sched_getaffinity(0))
probe("default affinity")
for nd in nodes:
    cp = [c for c in cpus_of(nd) if c in all_cpus]
    if not cp:
        print(os.path.basename(nd), "has none of our cpus")
        continue
    os.sched_setaffinity(0, cp)
    probe(f"pinned to {os.path.basename(nd)} ({len(cp)} cpus)")
os.sched_setaffinity(0, all_cpus)
