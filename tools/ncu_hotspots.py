"""Per-region / per-instruction warp-stall samples of one kernel from an ncu --set full --import-source on report:
    python tools/ncu_hotspots.py gpurun_out/<tag>/prof.ncu-rep > profiles/rN_ncu_hot_<tag>.txt   (runs without a GPU)"""
import csv, io, subprocess, sys
rep = sys.argv[1]
extra = sys.argv[2:]  # e.g. -k regex:k_sweep_fused -c 1
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + extra, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO("\n".join(l for l in txt.splitlines() if l.startswith('"')))))
hdr = rows[1]
ia, isrc, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples")
stall_cols = {h: k for k, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
data = []
for r in rows[2:]:
    try:
        data.append((int(r[ia], 16), r[isrc].strip(), int(r[isamp]), r))
    except Exception:
        pass
base = data[0][0]
tot = sum(d[2] for d in data)
print(f"# {rep}: {len(data)} SASS instructions, {tot} warp-stall samples")
# regions: split at backward branch targets = loops; simpler: bucket by 64 instructions and label with dominant opcodes
print("\n# cumulative sample share along the code (buckets of 48 instructions), with the opcode mix of each bucket")
B = 48
for k in range(0, len(data), B):
    chunk = data[k:k + B]
    s = sum(d[2] for d in chunk)
    if s < tot * 0.004:
        continue
    ops = {}
    for d in chunk:
        op = d[1].split()[0] if not d[1].startswith("@") else d[1].split()[1]
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0) + d[2]
    top = sorted(ops.items(), key=lambda x: -x[1])[:5]
    st = {h: sum(int(d[3][c] or 0) for d in chunk) for h, c in stall_cols.items()}
    tst = sorted(st.items(), key=lambda x: -x[1])[:4]
    print(f"  +{(chunk[0][0]-base):#07x}  {100*s/tot:5.1f}%  ops: " + ", ".join(f"{o} {100*v/s:.0f}%" for o, v in top) + "  | stalls: " + ", ".join(f"{h[6:]} {100*v/max(s,1):.0f}%" for h, v in tst))
print("\n# top 25 instructions by samples")
for d in sorted(data, key=lambda x: -x[2])[:25]:
    st = sorted(((h, int(d[3][c] or 0)) for h, c in stall_cols.items()), key=lambda x: -x[1])[:2]
    print(f"  +{(d[0]-base):#07x} {100*d[2]/tot:5.2f}%  {d[1][:60]:60s} " + ", ".join(f"{h[6:]} {v}" for h, v in st))
