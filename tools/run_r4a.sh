mkdir -p gpurun_out/r5f
for m in 32 16; do
for cfg in "bspline_lu 9" "bspline_lu 7" "bspline_lu 5" "bspline_lu 3"; do
set -- $cfg
SLB_SEG_M128=$m timeout 300 python bench.py --steps 5 --warmup 3 --interp $1 --order $2 --no-cpu --no-configs 2>>gpurun_out/r5f/err.log | tail -1 > gpurun_out/r5f/tmp.json
python - <<P
import json
d = json.load(open("gpurun_out/r5f/tmp.json")); k = d["roofline"]["all_kernels"]; c = d["config"]
print("M128=$m", c["interp"], c["order"], "ms/step %.3f" % d["ms_per_step"], {n.split("/")[1]: round(v["ms"], 3) for n, v in k.items() if "fused" not in n})
P
done
done
