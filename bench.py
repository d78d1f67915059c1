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
# dram__bytes_read.sum + dram__bytes_write.sum of one k_sweep_fused launch (ncu --set full), by grid size
NCU_TRAFFIC = {128: 4.383e9}


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

    warm = max(args.warmup, 3)  # same step count as the GPU arm, so that ee_after_timed is comparable
    ee = None
    for _ in range(warm):
        ee = step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ee = step()
    dt = time.perf_counter() - t0
    val = cells_per_step * args.steps / dt / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": warm, "ee_after_timed": ee, "steps_done": warm + args.steps, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
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
        "workload": f"2D2V {args.size}^4 Vlasov-Poisson Strang step, {args.interp} order {args.order}: 6 sweeps + 2 charge/Poisson solves "
                    "(value = 6 * n^4 cell-updates per step / device time per step)",
        "grid": [args.size] * 4, "interp": args.interp, "order": args.order, "dt": 0.1,
        "cache": "inputs larger than L2 (f = %.2f GB, ping-pong buffer of the same size)" % (args.size**4 * 8 / 1e9),
    }


def cpu_baseline_sample(args, steps_done):
    """CPU leg of our own arm: the oracle runs the SAME steps as the GPU arm (warm-up + timed, capped at 30: about
    10-40 s on the box's cores), recording the electric energy after every step.  Returns the `cpu_baseline`
    object (timed over those steps) and the oracle's ee history for the parity object."""
    from oracle import refmodel as R

    ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    n = args.size
    adv, vecs = vp2d2v_setup(R, n, args.order, args.interp, nthreads=ncores)
    f = np.empty((n,) * 4, order="F")
    fill_product(f, vecs)
    advd = R.AdvectionData(adv, f, R.getpoissonvar(adv))
    del f
    nsteps = min(steps_done, 30)
    hist = []
    t0 = time.perf_counter()
    for _ in range(nsteps):
        while R.advection(advd):
            pass
        hist.append(R.compute_ee(advd))
    dt = time.perf_counter() - t0
    cb = {"value": 6 * n**4 * nsteps / dt / 1e9, "unit": UNIT, "cores": ncores, "kind": "port",
          "sample": f"{nsteps} full Strang steps (6 sweeps + field solves each) of the 2D2V {n}^4 workload from the same initial "
                    f"condition as the GPU arm, {dt:.1f} s, oracle C/OpenMP port on {ncores} threads"}
    return cb, hist


def gpu_ee_history(S, args, nsteps):
    """ee after each of the first nsteps Strang steps on a fresh device-resident grid (same initial condition)"""
    n = args.size
    adv, vecs = vp2d2v_setup(S, n, args.order, args.interp)
    f = np.empty((n,) * 4, order="F")
    fill_product(f, vecs)
    advd = S.AdvectionData(adv, f, S.getpoissonvar(adv))
    del f
    hist = []
    for _ in range(nsteps):
        while S.advection(advd):
            pass
        hist.append(S.compute_ee(advd))
    advd.close()
    return hist


# ------------------------------------------------------------------------------------------
# the other BASELINE.json configurations (C1-C4), each timed and parity-checked in the same run
# ------------------------------------------------------------------------------------------
def _cfg_c1(M, **kw):
    """C1: Vlasov-Poisson 1D1V Landau damping 128 x 256, Lagrange 9, Strang (examples/vlasov-poisson-1d1v.jl:24-45)"""
    nx, nv = 128, 256
    mx, mv = M.UniformMesh(0.0, 2 * math.pi / 0.5, nx), M.UniformMesh(-6.0, 6.0, nv)
    adv = M.Advection((mx, mv), [M.Lagrange(9)] * 2, 0.1, [([2, 1], 1, 1, True), ([1, 2], 1, 2, True)], **kw)
    f = M.dotprod((1 + 0.001 * np.cos(0.5 * mx.points), np.exp(-mv.points**2 / 2) / math.sqrt(2 * math.pi)))
    return M.AdvectionData(adv, f, M.getpoissonvar(adv)), 3 * nx * nv


def _cfg_c2(M, **kw):
    """C2: 2-D rigid rotation 1024 x 1024, periodic B-spline order 5 (BSplineLU), magic splitting
    (examples/run_rotation.jl shape, test/test_rotation.jl:43-62)"""
    n = 1024
    m1, m2 = M.UniformMesh(-5.0, 5.0, n), M.UniformMesh(-5.0, 5.0, n)
    dt = 2 * math.pi / 100
    adv = M.Advection((m1, m2), [M.BSplineLU(5, n), M.BSplineLU(5, n)], dt, [([1, 2], 1, 1, True), ([2, 1], 1, 2, True)],
                      tab_coef=M.magicsplit(dt), **kw)
    x, y = m1.points[:, None], m2.points[None, :]
    return M.AdvectionData(adv, np.asfortranarray(np.exp(-13 * (x**2 + (y + 1.2) ** 2))), M.getrotationvar(adv)), 3 * n * n


def _cfg_vp4(n, order, kind):
    def build(M, **kw):
        adv, vecs = vp2d2v_setup(M, n, order, kind, **kw)
        f = np.empty((n,) * 4, order="F")
        fill_product(f, vecs)
        return M.AdvectionData(adv, f, M.getpoissonvar(adv)), 6 * n**4
    return build


def run_configs(S, ctx, args, peak, only=None, with_oracle=True):
    """C1-C4 on this GPU: ms/step (CUDA events), Gcell/s per sweep, the HBM fraction of the per-sweep algorithmic
    traffic (16 B per cell-update; C1/C2/C3 are cache-resident or launch-bound, flagged), and a
    parity figure: the grid after `psteps` Strang steps against the oracle on the same inputs (max-abs relative)."""
    from slb200 import _lib

    R = None
    if with_oracle:   # the oracle is the CHECKER of this leg (part of the cpu_baseline / parity leg; --no-cpu skips it)
        from oracle import refmodel as R

    ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    e0, e1 = ctx.event(), ctx.event()
    cases = [
        ("C1_vp1d1v_128x256_L9_strang", _cfg_c1, 200, 10, "256 KB grid: launch-bound, cache-resident", True),
        ("C2_rotation_1024x1024_bsplinelu5_magic", _cfg_c2, 50, 3, "8 MB grid: L2-resident", False),
        ("C3_vp2d2v_64^4_L7_strang", _cfg_vp4(64, 7, "lagrange"), 30, 2, "134 MB grid: just above L2", True),
        ("C4_vp2d2v_128^4_bsplinefft11_strang", _cfg_vp4(128, 11, "bspline_fft"), 5, 1, "2.15 GB grid: HBM-bound", True),
    ]
    out = {}
    for name, build, nsteps, psteps, note, has_ee in cases:
        if (args.size < 128 and name.startswith("C4")) or (only and not name.startswith(tuple(only))):
            continue  # reduced-size smoke runs of bench.py skip the big case
        try:
            g, cells = build(S)
            par = None
            if R is not None:
                o, _ = build(R, nthreads=ncores)
                for _ in range(psteps):
                    while S.advection(g):
                        pass
                    while R.advection(o):
                        pass
                a, b = g.getdata(), o.data
                par = {"steps": psteps, "f_rel_maxabs": float(np.max(np.abs(a - b)) / np.max(np.abs(b)))}
                if has_ee:
                    ee_g, ee_o = S.compute_ee(g), R.compute_ee(o)
                    par["ee_rel"] = abs(ee_g - ee_o) / abs(ee_o)
                del a, b, o
            for _ in range(3):
                while S.advection(g):
                    pass
            ctx.sync()
            l0 = ctx.launch_count()
            ctx.record(e0)
            for _ in range(nsteps):
                while S.advection(g):
                    pass
            ctx.record(e1)
            ms = _lib.Context.elapsed_ms(e0, e1) / nsteps
            nl = (ctx.launch_count() - l0) / nsteps
            extra = {}
            if name.startswith("C1") and hasattr(S, "StepGraph"):
                # launch-bound: whole steps recorded once and replayed as one CUDA-graph launch per two steps, with the
                # electric energy of every step still reduced on the device (the example's loop reads it per step)
                sg = S.StepGraph(g, nsteps=2)
                for _ in range(3):
                    sg.launch()
                ctx.sync()
                ctx.record(e0)
                for _ in range(nsteps // 2):
                    sg.launch()
                ctx.record(e1)
                extra["ms_per_step_graph"] = _lib.Context.elapsed_ms(e0, e1) / (nsteps // 2 * 2)
                t0 = time.perf_counter()
                for _ in range(nsteps // 2):
                    sg.launch()
                    sg.energies()
                extra["ms_per_step_graph_ee_readback_wall"] = (time.perf_counter() - t0) * 1e3 / (nsteps // 2 * 2)
                sg.close()
            if name.startswith("C1") and hasattr(S, "StepProgram"):
                # the same steps as ONE persistent cooperative kernel per launch (slb_program_*): 2 recorded steps
                # repeated nsteps / 2 times inside the kernel, ee of every step reduced on the device; its history is
                # compared bit for bit with the stepwise driver's from the same state
                try:
                    rep = max(1, nsteps // 2)
                    ga, gb = build(S)[0], build(S)[0]
                    for gg in (ga, gb):   # one real step: its last sweep leaves the line sums the next field solve starts from
                        while S.advection(gg):
                            pass
                    sp = S.StepProgram(gb, nsteps=2, repeat=rep)
                    sp.launch()
                    ee_prog = sp.energies()
                    ee_step = []
                    for _ in range(2 * rep):
                        while S.advection(ga):
                            pass
                        ee_step.append(S.compute_ee(ga))
                    extra["program_equals_stepwise_bitwise"] = bool(np.array_equal(gb.getdata(), ga.getdata()) and ee_prog == ee_step)
                    for _ in range(2):
                        sp.launch()
                    ctx.sync()
                    ctx.record(e0)
                    for _ in range(3):
                        sp.launch()
                    ctx.record(e1)
                    extra["ms_per_step_program"] = _lib.Context.elapsed_ms(e0, e1) / (3 * 2 * rep)
                    extra["program"] = {"ops_per_2_steps": sp.nops, "grid_barriers_per_2_steps": sp.nbarriers, "blocks": sp.nblocks,
                                        "steps_per_launch": 2 * rep,
                                        "block0_ns_per_op_kind_wait_run": [(k, round(w), round(r)) for k, w, r in sp.profile()]}
                    sp.close()
                    ga.close()
                    gb.close()
                except Exception as exc:
                    extra["program_error"] = f"{type(exc).__name__}: {exc}"
            g.close()
            out[name] = {"ms_per_step": ms, "Gcell_s": cells / ms / 1e6, "launches_per_step": nl,
                         "hbm_frac_per_sweep_bytes": cells * BYTES_PER_CELL / (ms * 1e-3) / 1e9 / peak, "parity": par, "note": note, **extra}
        except Exception as exc:  # a side measurement must not take the headline down
            out[name] = {"error": f"{type(exc).__name__}: {exc}"}
    return out


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
    nfused0 = advd.n_fused
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
    advd_nfused = advd.n_fused - nfused0
    clocks = sampler.stop()
    # electric energy right after the timed steps (nothing has touched the grid since): the cross-check
    # value every arm prints -- N = 1/2/4/8 and the oracle after the same number of steps must agree
    steps_done = max(args.warmup, 3) + args.steps
    ee_after_timed = S.compute_ee(advd)
    value = cells_per_step * args.steps / (ms_total * 1e-3) / 1e9
    # per-stage durations (include the field solve for the v1 stages)
    stage_ms = [_lib.Context.elapsed_ms(evs[j], evs[j + 1]) for j in range(i)]
    per_dim = {}
    for d, ms in zip(stage_dims, stage_ms):
        per_dim.setdefault(d, []).append(ms)

    # ---- roofline of the dominant kernel: the pair-fused pass (all 6 sweeps of a step run as 3 such
    # launches).  Timed alone with events, same stream, inside this run (no profiler). -------------
    peak, peak_src = read_peaks()
    reps = 5
    e0, e1 = ctx.event(), ctx.event()
    kern = {}
    table = np.linspace(-0.4, 0.4, n * n)  # |alpha| < 0.5 like (dt/dv) E
    vtab = adv.t_mesh[2].points
    tdev_E = ctx.to_device(table)
    tdev_v = ctx.to_device(vtab)
    it = adv.t_interp[0]
    sx = -0.1 / adv.t_mesh[0].step

    def timed(fn):
        fn()
        ctx.sync()
        ctx.record(e0)
        for _ in range(reps):
            fn()
        ctx.record(e1)
        return _lib.Context.elapsed_ms(e0, e1) / reps

    def stage(dim, tab, ln, strides, scale):
        return (dim, it, (tab, ln), strides, scale, True, 0, False)

    vE = lambda d: stage(d, tdev_E, n * n, [1, n, 0, 0], 1.0)
    pairs = (("k_sweep_fused/v1v2", vE(2), vE(3)),
             ("k_sweep_fused/x1x2", stage(0, tdev_v, n, [0, 0, 1, 0], sx), stage(1, tdev_v, n, [0, 0, 0, 1], sx)))
    for name, sa, sb in pairs:
        if not S.sweep_pair(advd, sa, sb):  # not pair-fused for this interpolation kind (B-spline pre-solves): nothing launched
            continue
        ms = timed(lambda: S.sweep_pair(advd, sa, sb))
        kern[name] = {"ms": ms, "GBps": n**4 * BYTES_PER_CELL / (ms * 1e-3) / 1e9, "Gcell_s": 2 * n**4 / (ms * 1e-3) / 1e9,
                      "cell_updates_per_cell": 2}
    singles = (("sweep/v2", vE(3)), ("sweep/v1", vE(2)), ("sweep/x2", stage(1, tdev_v, n, [0, 0, 0, 1], sx)),
               ("sweep/x1", stage(0, tdev_v, n, [0, 0, 1, 0], sx)))
    for name, sg in singles:
        ms = timed(lambda: S.sweep(advd, *sg[:6]))
        kern[name] = {"ms": ms, "GBps": n**4 * BYTES_PER_CELL / (ms * 1e-3) / 1e9, "Gcell_s": n**4 / (ms * 1e-3) / 1e9,
                      "cell_updates_per_cell": 1}
    ctx.free(tdev_E)
    ctx.free(tdev_v)
    fused_step = advd_nfused > 0
    if fused_step and "k_sweep_fused/v1v2" in kern:
        # the step runs as pair-fused passes: the v1v2 pass is 2 of its 3 big launches
        dom = "k_sweep_fused/v1v2"
        dom_note = ("one launch reads f once and writes it once (16 B per cell) and performs TWO sweeps (2 cell-updates per cell); "
                    "in SURVEY.md 8(d)'s per-sweep unit (16 B per cell-update) that is twice the single-sweep roofline rate")
        traffic, traffic_src = NCU_TRAFFIC.get(n), "profiles/r1_ncu_full_fused_L7_128_s4f.txt (ncu --set full, per launch)"
    else:
        # no pair fusion for this kind: the step is six single sweeps; the dominant kernel is the slowest of them
        dom = max((k for k in kern if k.startswith("sweep/")), key=lambda k: kern[k]["ms"])
        dom_note = "single sweep (pre-solve + stencil in one pass for the B-spline kinds): reads f once, writes it once, 16 B per cell-update"
        traffic, traffic_src = None, None
    ach = kern[dom]["GBps"]
    roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": ach / peak, "traffic": traffic, "traffic_source": traffic_src,
                "bytes_per_launch": n**4 * BYTES_PER_CELL, "note": dom_note,
                "ms_per_launch": kern[dom]["ms"], "all_kernels": kern}

    # ---- e2e: host buffers in, host buffers out, every step --------------------------------------
    # ONE grid.  Every step uploads its input f from pinned host memory and reads back its result (f and the
    # electric energy).  Uploads and read-backs run on their own streams: after step k, the read-back of its
    # result (front buffer) and the upload of step k+1's input (into the back buffer, which then becomes the
    # front) overlap -- PCIe is full duplex -- and step k+1 waits for both.
    fill_product(host, vecs)
    e2e_steps = max(4, min(args.steps, 8))
    nbytes = n**4 * 8
    host_out, hptr2 = _lib.pinned_empty((n,) * 4)
    cup, cdown = _lib.Context(ctx.device), _lib.Context(ctx.device)
    ev_step, ev_up, ev_down = ctx.event(), cup.event(), cdown.event()
    vp = _lib.C.c_void_p

    def streamed_steps(nsteps):
        advd.flush()
        advd.state_gen = 1
        _lib.check(L.slb_memcpy_h2d(cup.h, vp(L.slb_grid_front(advd.grid)), host.ctypes.data_as(vp), nbytes))  # input of step 0
        cup.record(ev_up)
        cdown.record(ev_down)
        ee_ = 0.0
        for _ in range(nsteps):
            ctx.wait_event(ev_up)      # this step's input has landed in the front buffer
            ctx.wait_event(ev_down)    # the previous result has left the buffer that is scratch now
            advd._linesum_dim = None
            step()
            ee_ = S.compute_ee(advd)   # D2H scalar; waits for the step
            ctx.record(ev_step)
            front, back = L.slb_grid_front(advd.grid), L.slb_grid_back(advd.grid)
            cdown.wait_event(ev_step)
            _lib.check(L.slb_memcpy_d2h(cdown.h, host_out.ctypes.data_as(vp), vp(front), nbytes))     # this step's result
            cdown.record(ev_down)
            cup.wait_event(ev_step)
            _lib.check(L.slb_memcpy_h2d(cup.h, vp(back), host.ctypes.data_as(vp), nbytes))            # the next step's input
            cup.record(ev_up)
            _lib.check(L.slb_grid_swap(advd.grid))
        cup.sync()
        cdown.sync()
        ctx.sync()
        return ee_

    streamed_steps(2)  # warm-up
    t0 = time.perf_counter()
    ee = streamed_steps(e2e_steps)
    wall_e2e = time.perf_counter() - t0
    e2e_val = cells_per_step * e2e_steps / wall_e2e / 1e9
    # serial variant: one grid, upload -> step -> read back, nothing overlapped
    ser_steps = max(1, min(args.steps, 3))
    advd.state_gen = 1
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(ser_steps):
        _lib.check(L.slb_grid_upload(advd.grid, host.ctypes.data_as(_lib.C.c_void_p)))
        advd._linesum_dim = None
        step()
        ee = S.compute_ee(advd)
        advd.getdata(out=host)
    wall_ser = time.perf_counter() - t0
    # host-side cost of issuing one step (ctypes + Python driver logic; the launches are asynchronous)
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step()
    host_issue_ms = (time.perf_counter() - t0) / e2e_steps * 1e3
    ctx.sync()
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
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes + 8, "steps": e2e_steps,
                "note": "ONE grid: every step uploads its input f from pinned host memory, runs the full Strang step, reads back ee "
                        "and f; copies run on their own streams, so the read-back of step k's result overlaps the upload of step "
                        "k+1's input (front/back buffers) and step k+1 waits for both; wall clock; PCIe-bound by construction: "
                        "2 x 2.15 GB per 2.7 ms step"},
        "e2e_serial": {"value": cells_per_step * ser_steps / wall_ser / 1e9, "unit": UNIT, "steps": ser_steps,
                       "note": "the same with nothing overlapped: AdvectionData.upload, step, compute_ee, getdata, one after the other"},
        "e2e_resident": {"value": cells_per_step * e2e_steps / wall_res / 1e9, "unit": UNIT,
                         "note": "f resident in HBM across steps (AdvectionData semantics), ee read back per step; wall clock"},
        "gpu_launches": int(launches),
        "host_issue_ms_per_step": host_issue_ms,
        "roofline": roofline,
        "gpu_fused_passes_per_step": advd_nfused / args.steps,
        "advection_call_ms": {f"dim{d}": float(np.mean(v)) for d, v in sorted(per_dim.items())},
        "advection_call_note": "device time between successive advection() calls: the first stage of a fused pair is only recorded "
                               "(dim2 = charge density + Poisson solve), the second runs both sweeps",
        "ee_after_timed": ee_after_timed, "steps_done": steps_done,
    }
    if not args.no_cpu:
        try:
            cb, hist_o = cpu_baseline_sample(args, steps_done)
            line["cpu_baseline"] = cb
            hist_g = gpu_ee_history(S, args, len(hist_o))
            ho, hg = np.array(hist_o), np.array(hist_g)
            line["parity"] = {
                "what": "electric-energy history of the first %d Strang steps from the same initial condition: this GPU path vs the oracle "
                        "(CPU restatement of the reference; itself pinned by the reference's KATs only -- parity unpinned at the ulp "
                        "level, no Julia runtime here)" % len(hist_o),
                "steps": len(hist_o), "ee_hist_rel_vs_oracle": float(np.max(np.abs(hg - ho)) / np.max(np.abs(ho))),
                "ee_oracle_last": hist_o[-1], "ee_gpu_last": hist_g[-1],
                "ee_after_timed_equals_fresh_run": (bool(hist_g[-1] == ee_after_timed) if len(hist_g) == steps_done else None),
                "tolerance": 1e-10,
            }
        except Exception as exc:  # the checker must never take the product down
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {exc}"}
    if not args.no_configs:
        advd.close()
        line["configs"] = run_configs(S, ctx, args, peak, with_oracle=not args.no_cpu)
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
    ap.add_argument("--no-configs", action="store_true", help="skip the C1-C4 side measurements")
    ap.add_argument("--no-e2e", action="store_true", help="sharded runs: skip the host-buffer end-to-end leg (large grids)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="multi-GPU re-shard: peer stores fused into the sweep (p2p) or NCCL all-to-all")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
