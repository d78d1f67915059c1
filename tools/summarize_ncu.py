#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries kept under
profiles/ (tracked).  Runs here (no GPU needed).

    python tools/summarize_ncu.py launches gpurun_out/<tag>/launches.csv  > profiles/rN_launches_<tag>.txt
    python tools/summarize_ncu.py full     gpurun_out/<tag>/prof.ncu-rep  > profiles/rN_ncu_full_<tag>.txt
    python tools/summarize_ncu.py stalls   gpurun_out/<tag>/prof.ncu-rep [kernel-id]   # per-line stall samples
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
    "launch__block_size", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_lg_throttle",
    "smsp__pcsamp_warps_issue_stalled_mio_throttle", "smsp__pcsamp_warps_issue_stalled_barrier",
    "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_selected",
    "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_dispatch_stall",
    "smsp__pcsamp_warps_issue_stalled_no_instructions", "smsp__pcsamp_warps_issue_stalled_branch_resolving",
    "smsp__pcsamp_warps_issue_stalled_drain", "smsp__pcsamp_warps_issue_stalled_membar",
    "smsp__pcsamp_warps_issue_stalled_imc_miss", "smsp__pcsamp_warps_issue_stalled_sleeping",
]


def clean_csv(text):
    lines = [ln for ln in text.splitlines() if ln.startswith('"')]
    return list(csv.reader(io.StringIO("\n".join(lines))))


def launches(path):
    rows = clean_csv(open(path).read())
    hdr, rows = rows[0], rows[1:]
    ik, iv, ig, ib = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows:
        name = r[ik].split("(")[0]
        a = agg.setdefault(name, [0, 0.0, r[ig], r[ib]])
        a[0] += 1
        a[1] += float(r[iv].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {len(rows)} launches, {tot/1e6:.3f} ms of device time (ncu: cold-cache, serialised -- compare SHARES)")
    print(f"{'kernel':58s} {'n':>4s} {'total us':>10s} {'avg us':>9s} {'share':>6s}  grid / block of first launch")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k[:58]:58s} {a[0]:4d} {a[1]/1e3:10.1f} {a[1]/a[0]/1e3:9.2f} {a[1]/tot*100:5.1f}%  {a[2]} / {a[3]}")


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = clean_csv(out)
    return rows[0], rows[1], rows[2:]


def full(rep):
    hdr, units, rows = raw_page(rep)
    ik = hdr.index("Kernel Name")
    print(f"# {rep}: ncu --set full, one column per captured launch (values under a profiler are not bench numbers)")
    names = [f"{r[hdr.index('ID')]}:{r[ik].split('(')[0]}" for r in rows]
    print("metric [unit]".ljust(66) + "  ".join(n[:28].rjust(28) for n in names))
    for key in KEYS:
        if key not in hdr:
            continue
        i = hdr.index(key)
        print(f"{key} [{units[i]}]".ljust(66)[:66] + "  ".join(r[i].rjust(28) for r in rows))
    # derived: traffic and achieved bandwidth
    try:
        ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")

        def tobytes(v, u):
            m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            return float(v.replace(",", "")) * m.get(u, 1)

        def tons(v, u):
            m = {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}
            return float(v.replace(",", "")) * m.get(u, 1)

        vals = []
        for r in rows:
            tr = tobytes(r[ir], units[ir]) + tobytes(r[iw], units[iw])
            ns = tons(r[it], units[it])
            vals.append(f"{tr/1e9:.3f} GB, {tr/ns:.0f} GB/s")
        print("derived: dram traffic, dram GB/s under ncu".ljust(66) + "  ".join(v.rjust(28) for v in vals))
    except ValueError:
        pass


def stalls(rep, kid=None):
    cmd = ["ncu", "-i", rep, "--page", "source", "--csv"]
    if kid is not None:
        cmd += ["--kernel-id", f":::{kid}"]
    sys.stdout.write(subprocess.run(cmd, capture_output=True, text=True).stdout)


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        launches(sys.argv[2])
    elif mode == "full":
        full(sys.argv[2])
    elif mode == "stalls":
        stalls(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        sys.exit(__doc__)
