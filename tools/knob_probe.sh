#!/bin/bash
# env-only probe of the pair-fused pass's tile knobs (128^4, order from $1, default 7): one line per setting
O=${1:-7}
mkdir -p gpurun_out/knobs
for cfg in "" "SLB_FUSED_CC_THREADS=128" "SLB_FUSED_CC_THREADS=32" "SLB_FUSED_THREADS=128" "SLB_FUSED_G=16"; do
env $cfg timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu --no-configs --order $O 2>>gpurun_out/knobs/err.log | tail -1 > gpurun_out/knobs/tmp.json
python - <<P
import json
d=json.load(open("gpurun_out/knobs/tmp.json"))
print("[$cfg]", round(d["ms_per_step"],4), {k: round(v["ms"],4) for k,v in d["roofline"]["all_kernels"].items() if "fused" in k})
P
done
