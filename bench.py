#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 semi-Lagrangian sweep path.

Metric (BASELINE.json): Gcell-updates/s per FP64 advection sweep (2D2V 128^4).
A "step" is one full Strang-split Vlasov-Poisson time step on the 2D2V 128^4 grid with
Lagrange order 7: 6 one-dimensional sweeps (v1 v2 | x1 x2 | v1 v2) plus the charge-density
reductions and Poisson solves that feed the velocity sweeps (examples/vlasov-poisson-2d2v.jl).
value = (6 * 128^4 cell-updates * n_steps) / device time, i.e. the average per-sweep rate with
every direction and the field solves included.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Under torchrun (N > 1) the 128^4 grid is sharded over the N ranks (strong scaling) and
re-sharded with all-to-all exchanges (slb200.distributed).  `--impl reference` times the
reference's CPU algorithm (the oracle restatement, oracle/) on the host cores.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "semilagrangian.jl_b200"))
sys.path.insert(0, ROOT)

import numpy as np

METRIC = "Gcell-updates/s per FP64 advection sweep (2D2V 128^4)"
UNIT = "Gcell/s"
BYTES_PER_CELL = 16.0  # read f once (8 B) + write once (8 B)


def read_peaks():
    """measured HBM peak (driver-written MEASURED_PEAKS.json) else the profiling guide's fallback"""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._pump, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(self.NAMES, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------
def vp2d2v_setup(M, n, order, kind="lagrange", **kw):
    """2D2V Vlasov-Poisson of examples/vlasov-poisson-2d2v.jl:48-83 on an n^4 grid."""
    sz = (n, n, n, n)
    m1 = M.UniformMesh(0.0, 4 * math.pi, n)
    m2 = M.UniformMesh(0.0, 4 * math.pi, n)
    v1 = M.UniformMesh(-6.0, 6.0, n)
    v2 = M.UniformMesh(-6.0, 6.0, n)
    mk = {"lagrange": lambda k: M.Lagrange(order), "bspline_lu": lambda k: M.BSplineLU(order, k),
          "bspline_fft": lambda k: M.BSplineFFT(order, k), "hermite": lambda k: M.Hermite(order)}[kind]
    tabst = [([3, 4, 1, 2], 1, 1, True), ([4, 3, 1, 2], 1, 1, True), ([1, 2, 4, 3], 1, 2, True), ([2, 1, 3, 4], 1, 2, True)]
    adv = M.Advection((m1, m2, v1, v2), [mk(k) for k in sz], 0.1, tabst, **kw)
    eps = 0.5
    fsp = lambda x: eps * np.cos(x / 2) + 1
    fv = lambda v: np.exp(-v**2 / 2) / math.sqrt(2 * math.pi)
    vecs = (fsp(m1.points), fsp(m2.points), fv(v1.points), fv(v2.points))
    return adv, vecs


def fill_product(out, vecs):
    """out[i,j,k,l] = a[i] b[j] c[k] d[l] without a second full-size temporary"""
    a, b, c, d = vecs
    plane = np.multiply.outer(b, a).T.copy(order="F")  # [x1, x2]
    for l in range(len(d)):
        for k in range(len(c)):
            out[:, :, k, l] = plane * (c[k] * d[l])


# ------------------------------------------------------------------------------------------
# reference arm: the reference's CPU algorithm (oracle restatement) on the host cores
# ------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import refmodel as R
    from oracle import clib

    ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    n = args.size
    adv, vecs = vp2d2v_setup(R, n, args.order, args.interp, nthreads=ncores)
    f = np.empty((n,) * 4, order="F")
    fill_product(f, vecs)
    advd = R.AdvectionData(adv, f, R.getpoissonvar(adv))
    del f
    cells_per_step = 6 * n**4

    def step():
        while R.advection(advd):
            pass
        return R.compute_ee(advd)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = cells_per_step * args.steps / dt / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": ncores, "kind": "port",
                         "sample": f"full Strang step (6 sweeps + field solves) on the full 2D2V {n}^4 grid, oracle C/OpenMP port "
                                   f"of the reference algorithm (allocation-free, hence optimistic for the Julia original)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def workload_config(args):
    return {
        "workload": f"2D2V {args.size}^4 Vlasov-Poisson Strang step, {args.interp} order {args.order}: 6 sweeps + 2 charge/Poisson solves",
        "grid": [args.size] * 4, "interp": args.interp, "order": args.order, "dt": 0.1,
        "cache": "inputs larger than L2 (f = %.2f GB, ping-pong buffer of the same size)" % (args.size**4 * 8 / 1e9),
    }


def cpu_baseline_sample(args):
    """Bounded CPU sample for the `cpu_baseline` object of our own arm: one Strang step of the
    oracle on the full grid (a few seconds on the box's cores)."""
    from oracle import refmodel as R

    ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    n = args.size
    adv, vecs = vp2d2v_setup(R, n, args.order, args.interp, nthreads=ncores)
    f = np.empty((n,) * 4, order="F")
    fill_product(f, vecs)
    advd = R.AdvectionData(adv, f, R.getpoissonvar(adv))
    del f
    R.advection(advd)  # first-touch warm-up of the scratch array (one stage)
    advd.state_gen = 1
    t0 = time.perf_counter()
    while R.advection(advd):
        pass
    R.compute_ee(advd)
    dt = time.perf_counter() - t0
    return {"value": 6 * n**4 / dt / 1e9, "unit": UNIT, "cores": ncores, "kind": "port",
            "sample": f"1 full Strang step (6 sweeps + field solves) of the 2D2V {n}^4 workload, {dt:.1f} s, oracle C/OpenMP port"}


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def run_ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 or args.gpus > 1:
        from bench_dist import run_distributed  # sharded path (torch.distributed + NCCL)

        return run_distributed(args, sys.modules[__name__])
    import slb200 as S
    from slb200 import _lib

    ctx = S.default_context()
    n = args.size
    adv, vecs = vp2d2v_setup(S, n, args.order, args.interp)
    host, hptr = _lib.pinned_empty((n,) * 4)
    fill_product(host, vecs)
    pv = S.getpoissonvar(adv)
    advd = S.AdvectionData(adv, host, pv)
    cells_per_step = 6 * n**4
    L = _lib.lib()

    def step():
        while S.advection(advd):
            pass

    # ---- warm-up ----------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step()
    ctx.sync()

    # ---- timed region: K resident steps, events around every sweep ----------------------
    nev = 6 * args.steps + 1
    evs = [ctx.event() for _ in range(nev)]
    sampler = ClockSampler(ctx.device)
    sampler.start()
    time.sleep(0.3)
    launches0 = ctx.launch_count()
    ctx.sync()
    ctx.timer_start()
    i = 0
    ctx.record(evs[0])
    stage_dims = []
    for _ in range(args.steps):
        more = True
        while more:
            stage_dims.append(advd.getst().perm[0] - 1)
            more = S.advection(advd)
            i += 1
            ctx.record(evs[i])
    ms_total = ctx.timer_stop()
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop()
    value = cells_per_step * args.steps / (ms_total * 1e-3) / 1e9
    # per-stage durations (include the field solve for the v1 stages)
    stage_ms = [_lib.Context.elapsed_ms(evs[j], evs[j + 1]) for j in range(i)]
    per_dim = {}
    for d, ms in zip(stage_dims, stage_ms):
        per_dim.setdefault(d, []).append(ms)

    # ---- roofline of the dominant kernel: the strided sweep (5 of the 6 stages) ---------
    # timed alone with events, same stream, inside this run (no profiler): v2 sweep = pure kernel
    peak, peak_src = read_peaks()
    reps = 5
    e0, e1 = ctx.event(), ctx.event()
    kern = {}
    table = np.linspace(-0.4, 0.4, n * n)  # |alpha| < 0.5 like (dt/dv) E
    vtab = adv.t_mesh[2].points
    for name, dim, tab, strides, scale in (
        ("k_sweep_strided/v2", 3, table, [1, n, 0, 0], 1.0),
        ("k_sweep_strided/v1", 2, table, [1, n, 0, 0], 1.0),
        ("k_sweep_strided/x2", 1, vtab, [0, 0, 0, 1], -0.1 / adv.t_mesh[1].step),
        ("k_sweep_contig/x1", 0, vtab, [0, 0, 1, 0], -0.1 / adv.t_mesh[0].step),
    ):
        tdev = ctx.to_device(tab)
        S.sweep(advd, dim, adv.t_interp[dim], (tdev, len(tab)), strides, scale, True)
        ctx.sync()
        ctx.record(e0)
        for _ in range(reps):
            S.sweep(advd, dim, adv.t_interp[dim], (tdev, len(tab)), strides, scale, True)
        ctx.record(e1)
        ms = _lib.Context.elapsed_ms(e0, e1) / reps
        kern[name] = {"ms": ms, "GBps": n**4 * BYTES_PER_CELL / (ms * 1e-3) / 1e9, "Gcell_s": n**4 / (ms * 1e-3) / 1e9}
        ctx.free(tdev)
    dom = "k_sweep_strided/v2"
    ach = kern[dom]["GBps"]
    roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": ach / peak, "traffic": None, "bytes_per_launch": n**4 * BYTES_PER_CELL,
                "ms_per_launch": kern[dom]["ms"], "all_kernels": kern}

    # ---- e2e: host buffers in, host buffers out, every step ------------------------------
    # restore a physical state first (the roofline sweeps above used synthetic shifts)
    fill_product(host, vecs)
    e2e_steps = max(1, min(args.steps, 5))
    advd.upload(host)
    advd.state_gen = 1
    ctx.sync()
    t0 = time.perf_counter()
    ctx.timer_start()
    ee = 0.0
    for _ in range(e2e_steps):
        _lib.check(L.slb_grid_upload(advd.grid, host.ctypes.data_as(_lib.C.c_void_p)))   # H2D of this step's input f
        step()
        ee = S.compute_ee(advd)                                                          # D2H scalar (the step's metric)
        advd.getdata(out=host)                                                           # D2H of the step's result f
    ms_e2e = ctx.timer_stop()
    wall_e2e = time.perf_counter() - t0
    e2e_val = cells_per_step * e2e_steps / max(ms_e2e * 1e-3, wall_e2e) / 1e9
    nbytes = n**4 * 8
    # resident variant (how the reference API is normally driven: f stays inside AdvectionData,
    # only the electric energy comes back each step)
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step()
        ee = S.compute_ee(advd)
    ctx.sync()
    wall_res = time.perf_counter() - t0

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args), "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes + 8,
                "steps": e2e_steps, "note": "upload f from pinned host memory, full Strang step, read back ee and f, every step"},
        "e2e_resident": {"value": cells_per_step * e2e_steps / wall_res / 1e9, "unit": UNIT,
                         "note": "f resident in HBM across steps (AdvectionData semantics), ee read back per step; wall clock"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "stage_ms": {f"dim{d}": float(np.mean(v)) for d, v in sorted(per_dim.items())},
        "last_ee": ee,
    }
    if not args.no_cpu:
        try:
            line["cpu_baseline"] = cpu_baseline_sample(args)
        except Exception as exc:  # the checker must never take the product down
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {exc}"}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--order", type=int, default=7)
    ap.add_argument("--interp", default="lagrange", choices=["lagrange", "bspline_lu", "bspline_fft", "hermite"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline sample")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="multi-GPU re-shard: peer stores fused into the sweep (p2p) or NCCL all-to-all")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
