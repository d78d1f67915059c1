"""Sharded arm of bench.py (N > 1): 2D2V 128^4 grid domain-decomposed over the ranks
(strong scaling), re-sharded twice per Strang step with an NCCL all-to-all over NVLink
(slb200.distributed).  Launched by torchrun; rank 0 prints the JSON line."""
import json
import os
import time

import numpy as np


def run_distributed(args, B):
    import torch
    import torch.distributed as dist

    import slb200 as S
    from slb200.distributed import ShardedAdvectionData, slab

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
    loc = np.empty((n, hi - lo, n, n), order="F")
    B.fill_product(loc, (a, b[lo:hi], c, d))
    exchange = getattr(args, "exchange", "p2p")
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

    def step():
        while sh.advection():
            pass

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    sampler = B.ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = sh.ctx.launch_count()
    ex0 = sh.n_exchanges
    nf0 = sh.n_fused
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    dist.barrier()
    e0.record(sh.stream)  # events on the stream the sweeps and collectives run on
    for _ in range(args.steps):
        step()
    e1.record(sh.stream)
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    launches = sh.ctx.launch_count() - launches0
    nex = sh.n_exchanges - ex0
    nfused = sh.n_fused - nf0
    clocks = sampler.stop() if rank == 0 else None
    ee = sh.compute_ee()

    # sweep-only and exchange-only timings (explain the step time)
    def timed(fn, reps=5):
        torch.cuda.synchronize()
        dist.barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(sh.stream):
            a0.record(sh.stream)
            for _ in range(reps):
                fn()
            a1.record(sh.stream)
        torch.cuda.synchronize()
        t = torch.tensor([a0.elapsed_time(a1) / reps], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    nbytes_local = n**4 * 8 // world
    if sh.exchange == "nccl":
        ms_a2a = timed(lambda: dist.all_to_all_single(sh.bufs[1 - sh.cur], sh.bufs[sh.cur]))
    else:
        ms_a2a = None

    # e2e: host slab in, host slab out, every step (pinned host memory)
    from slb200 import _lib

    e2e_steps = max(1, min(args.steps, 3))
    if getattr(args, "no_e2e", False):
        e2e_val = None
    else:
        host, _hp = _lib.pinned_empty((n**4 // world,))
        torch.cuda.synchronize()
        sh.download_local(host)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            sh.upload_local(host)       # H2D of this rank's slab (pinned), on the driver's stream
            step()
            _ = sh.compute_ee()         # D2H scalar
            sh.download_local(host)     # D2H of the slab; synchronises
        dist.barrier()
        wall = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(wall, op=dist.ReduceOp.MAX)
        e2e_val = cells_per_step * e2e_steps / float(wall.item()) / 1e9

    if rank == 0:
        peak, peak_src = B.read_peaks()
        value = cells_per_step * args.steps / (ms_total * 1e-3) / 1e9
        fused = nfused > 0
        payload = nbytes_local * (world - 1) // world
        hbm_bytes = (3 * 16 + 8 if fused else 6 * 16 + 8) * n**4 / world
        line = {
            "metric": B.METRIC, "value": value, "unit": B.UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": dict(B.workload_config(args), parallelism=f"x2/v2 slabs over {world} GPUs, 2 all-to-all re-shards per step"),
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": B.UNIT, "h2d_bytes_per_step": nbytes_local * world, "d2h_bytes_per_step": nbytes_local * world + 8 * world,
                    "steps": e2e_steps, "note": "every rank uploads its slab from pinned host memory, full Strang step, reads back ee and its slab"},
            "gpu_launches": int(launches), "exchanges_per_step": nex / args.steps, "exchange": sh.exchange,
            "all_to_all_ms": ms_a2a,
            "all_to_all_GBps_per_gpu": (nbytes_local * (world - 1) / world / (ms_a2a * 1e-3) / 1e9) if ms_a2a else None,
            "exchange_payload_bytes_per_gpu": nbytes_local * (world - 1) // world,
            "gpu_fused_passes_per_step": nfused / args.steps,
            "roofline": {"bound": "hbm", "kernel": "whole step per rank: %s" % ("3 fused passes (16 B/cell each) + 1 rho pass (8 B/cell)" if fused else "6 sweeps (16 B/cell) + 1 rho pass (8 B/cell)"),
                         "achieved": hbm_bytes / (ms_total / args.steps * 1e-3) / 1e9, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                         "frac": hbm_bytes / (ms_total / args.steps * 1e-3) / 1e9 / peak, "traffic": None,
                         "bytes_per_step_per_gpu": hbm_bytes},
            "nvlink": {"bytes_out_per_gpu_per_step": 2 * payload, "peak_GBps": 770.0, "peak_source": "B200_PROFILING.md peer copy",
                       "bound_ms": 2 * payload / 770e9 * 1e3, "frac_of_step": (2 * payload / 770e9 * 1e3) / (ms_total / args.steps),
                       "note": "two re-shards per Strang step; their stores ride inside the v2 / x2 passes (peer memory), no separate collective"},
            "last_ee": ee,
        }
        print(json.dumps(line))
    sh.close()
    dist.destroy_process_group()
    return 0
