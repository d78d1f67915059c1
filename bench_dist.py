"""Sharded arm of bench.py (N > 1): the 2D2V 128^4 grid is split over the ranks (strong scaling).  Launched by
torchrun; rank 0 prints the JSON line.  torch.distributed is the launcher-side plumbing only (carrying the IPC
handles once, the max over ranks of the timings): the data path is libslb200's own kernels.

  halo driver (slb200/sharded.py, Lagrange/Hermite): slabs along v2 for the whole run, NO transposes; the halo
      planes travel as NVLink peer stores inside the passes, rho as a 131 KB mailbox all-gather.
  transposing driver (slb200/distributed.py; B-spline kinds, or SLB_SHARD=transpose): two re-shards per step fused
      into the stores of the passes before them.
"""
import json
import os
import time

import numpy as np


def run_distributed(args, B):
    import torch
    import torch.distributed as dist

    import slb200 as S
    from slb200 import _lib
    from slb200.distributed import ShardedAdvectionData, slab
    from slb200.sharded import HaloShardedAdvectionData, HaloUnsupported, torch_allgather_bytes

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    n = args.size
    adv, vecs = B.vp2d2v_setup(S, n, args.order, args.interp)
    lo, hi = slab(n, world, rank)
    a, b, c, d = vecs
    mode = os.environ.get("SLB_SHARD", "halo")
    max_shift = os.environ.get("SLB_MAX_SHIFT", "auto")   # cells; "auto": from the initial field, with a 1.5x margin
    max_shift = max_shift if max_shift == "auto" else float(max_shift)
    sh = None
    if mode == "halo":
        try:
            loc = np.empty((n, n, n, hi - lo), order="F")
            B.fill_product(loc, (a, b, c, d[lo:hi]))
            sh = HaloShardedAdvectionData(adv, loc, rank, world, torch_allgather_bytes(dist), device=local, max_shift=max_shift)
            driver = "halo"
        except HaloUnsupported as exc:
            if rank == 0:
                print(f"[bench] halo driver not applicable ({exc}); using the transposing driver", flush=True)
            sh = None
    if sh is None:
        loc = np.empty((n, hi - lo, n, n), order="F")
        B.fill_product(loc, (a, b[lo:hi], c, d))
        exchange = getattr(args, "exchange", "p2p")
        driver = "transpose"
        try:
            sh = ShardedAdvectionData(adv, loc, exchange=exchange)
        except Exception as exc:  # CUDA IPC unavailable (container restrictions): NCCL all-to-all instead
            if exchange != "p2p":
                raise
            if rank == 0:
                print(f"[bench] p2p exchange unavailable ({exc}); falling back to NCCL all-to-all", flush=True)
            sh = ShardedAdvectionData(adv, loc, exchange="nccl")
    del loc
    cells_per_step = 6 * n**4
    ctx = sh.ctx

    def step():
        while sh.advection():
            pass

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    sampler = B.ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count()
    nf0 = sh.n_fused
    e0, e1 = ctx.event(), ctx.event()   # CUDA events on the stream the passes run on
    torch.cuda.synchronize()
    dist.barrier()
    ctx.record(e0)
    for _ in range(args.steps):
        step()
    ctx.record(e1)
    torch.cuda.synchronize()
    dist.barrier()
    # electric energy right after the timed steps: equal on every rank, across N = 1/2/4/8 and to the oracle's
    ee_after_timed = sh.compute_ee()
    steps_done = warm + args.steps
    # where the step goes (outside the timed region): device time between successive advection() calls
    nst = adv.nbstates
    evs = [ctx.event() for _ in range(3 * nst + 1)]
    ctx.record(evs[0])
    dims = []
    for k in range(3 * nst):
        dims.append(sh.getst().perm[0] - 1)
        sh.advection()
        ctx.record(evs[k + 1])
    torch.cuda.synchronize()
    per_dim = {}
    for k, dd in enumerate(dims):
        per_dim.setdefault(dd, []).append(_lib.Context.elapsed_ms(evs[k], evs[k + 1]))
    call_ms = {f"dim{dd}": float(np.mean(v)) for dd, v in sorted(per_dim.items())}
    dist.barrier()
    ms = torch.tensor([_lib.Context.elapsed_ms(e0, e1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    launches = ctx.launch_count() - launches0
    nfused = sh.n_fused - nf0
    clocks = sampler.stop() if rank == 0 else None

    # e2e: host slab in, host slab out, every step (pinned host memory)
    nbytes_local = n**4 * 8 // world
    e2e_steps = max(1, min(args.steps, 3))
    if getattr(args, "no_e2e", False):
        e2e_val = None
    else:
        host, _hp = _lib.pinned_empty((n**4 // world,))
        torch.cuda.synchronize()
        sh.download_local(host)
        torch.cuda.synchronize()
        streamed = driver == "halo"
        if streamed:
            host_out, _hp2 = _lib.pinned_empty((n**4 // world,))
            sh.upload_local(host)
            sh.sync_ranks()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            if streamed:
                # this step's input is already on its way (or resident); after the step the read-back of its result
                # overlaps the upload of the next step's input, both on their own streams
                step()
                _ = sh.compute_ee()                    # D2H scalar
                sh.stream_io_exchange(host_out, host)  # D2H of the slab | H2D of the next slab, halos re-sent
            else:
                sh.upload_local(host)       # H2D of this rank's slab (pinned), on the driver's stream
                step()
                _ = sh.compute_ee()         # D2H scalar
                sh.download_local(host)     # D2H of the slab; synchronises
        sh.ctx.sync()
        torch.cuda.synchronize()
        dist.barrier()
        wall = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(wall, op=dist.ReduceOp.MAX)
        e2e_val = cells_per_step * e2e_steps / float(wall.item()) / 1e9

    if rank == 0:
        peak, peak_src = B.read_peaks()
        value = cells_per_step * args.steps / (ms_total * 1e-3) / 1e9
        ms_step = ms_total / args.steps
        if driver == "halo":
            H, cs = sh.H, sh.c
            plane = n**3 * 8
            out_bytes = 2 * 2 * H * plane     # two pushing passes per step, H planes to either neighbour
            hbm_bytes = (2 * (cs + 2 * H) + cs) * plane + 3 * cs * plane + cs * plane  # reads of 2 v passes + x pass, 3 writes, 1 rho pass
            par = (f"v2 slabs of {cs} planes + {H} halo planes per side (velocity shifts below {sh.max_shift:.2f} cells) over {world} GPUs, "
                   "no transposes: halo planes pushed as peer stores inside the passes")
            link_note = "halo planes ride inside the x1x2 pass and the step's last v1v2 pass as NVLink peer stores; rho: one 131 KB mailbox all-gather per field solve"
            kern = "whole step per rank: 3 fused passes (v passes read c + 2H rows) + 1 rho pass"
        else:
            payload = nbytes_local * (world - 1) // world
            out_bytes = 2 * payload
            fused = nfused > 0
            hbm_bytes = (3 * 16 + 8 if fused else 6 * 16 + 8) * n**4 / world
            par = f"x2/v2 slabs over {world} GPUs, 2 all-to-all re-shards per step"
            link_note = "two re-shards per Strang step; their stores ride inside the v2 / x2 passes (peer memory), no separate collective"
            kern = "whole step per rank: %s" % ("3 fused passes (16 B/cell each) + 1 rho pass (8 B/cell)" if fused else "6 sweeps (16 B/cell) + 1 rho pass (8 B/cell)")
        line = {
            "metric": B.METRIC, "value": value, "unit": B.UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": dict(B.workload_config(args), parallelism=par, driver=driver),
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": B.UNIT, "h2d_bytes_per_step": nbytes_local * world, "d2h_bytes_per_step": nbytes_local * world + 8 * world,
                    "steps": e2e_steps, "note": "every step: every rank uploads its slab from pinned host memory, full Strang step, reads back ee and its slab; halo driver: "
                                                "the read-back of step k overlaps the upload of step k+1 (own streams), transposing driver: one after the other"},
            "gpu_launches": int(launches), "gpu_fused_passes_per_step": nfused / args.steps,
            "roofline": {"bound": "hbm", "kernel": kern, "achieved": hbm_bytes / (ms_step * 1e-3) / 1e9, "peak": peak, "peak_source": peak_src,
                         "unit": "GB/s", "frac": hbm_bytes / (ms_step * 1e-3) / 1e9 / peak, "traffic": None, "bytes_per_step_per_gpu": hbm_bytes},
            "nvlink": {"bytes_out_per_gpu_per_step": out_bytes, "peak_GBps": 770.0, "peak_source": "B200_PROFILING.md peer copy",
                       "bound_ms": out_bytes / 770e9 * 1e3, "frac_of_step": (out_bytes / 770e9 * 1e3) / ms_step, "note": link_note},
            "advection_call_ms": call_ms,
            "advection_call_note": "rank 0, device time between successive advection() calls: dim2 = charge density + mailbox all-gather + "
                                   "Poisson solve (its sweep is deferred), dim3 = the v1v2 pass, dim1 = the x1x2 pass",
            "ee_after_timed": ee_after_timed, "steps_done": steps_done,
        }
        print(json.dumps(line))
    torch.cuda.synchronize()
    dist.barrier()
    sh.close()
    dist.destroy_process_group()
    return 0
