#!/bin/bash
# per-kind / per-direction kernel timings and full-step rates on one GPU:
# one bench.py line per interpolation kind into gpurun_out/<tag>/kinds.jsonl
TAG=${1:-kinds}; OUT=gpurun_out/$TAG; mkdir -p $OUT; : > $OUT/kinds.jsonl
for cfg in "lagrange 3" "lagrange 5" "lagrange 7" "lagrange 9" "lagrange 11" "hermite 5" "hermite 9" "bspline_lu 3" "bspline_lu 5" "bspline_fft 7" "bspline_fft 11"; do
  set -- $cfg
  timeout 300 python bench.py --steps 5 --warmup 3 --interp $1 --order $2 --no-cpu 2>/dev/null | tail -1 >> $OUT/kinds.jsonl
done
python - <<PY
import json
print(f"{'kind':12s} {'order':>5s} {'ms/step':>8s} {'Gcell/s':>8s} | fused v1v2  x1x2 | strided v2    v1    x2 | contig x1   (ms per launch)")
for ln in open("$OUT/kinds.jsonl"):
    d = json.loads(ln); k = d["roofline"]["all_kernels"]; c = d["config"]
    f = lambda n: f"{k[n]['ms']:6.3f}"
    print(f"{c['interp']:12s} {c['order']:5d} {d['ms_per_step']:8.3f} {d['value']:8.1f} | {f('k_sweep_fused/v1v2')} {f('k_sweep_fused/x1x2')} | {f('k_sweep_strided/v2')} {f('k_sweep_strided/v1')} {f('k_sweep_strided/x2')} | {f('k_sweep_contig/x1')}")
PY
