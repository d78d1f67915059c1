mkdir -p gpurun_out/r4n
timeout 900 python -m pytest tests/test_gpu_sweep.py -q -x > gpurun_out/r4n/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r4n/pytest.log
for cfg in "32768 4" "18000 4" "45000 2"; do
set -- $cfg
SLB_CONTIG_TILE_BYTES=$1 SLB_CONTIG_TILE_CTAS=$2 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-configs 2>gpurun_out/r4n/bench.err | tail -1 > gpurun_out/r4n/bench_$1_$2.json
python - <<P
import json
try:
    d=json.load(open("gpurun_out/r4n/bench_$1_$2.json"))
    print("$cfg", d["ms_per_step"], {k: round(v["ms"],4) for k,v in d["roofline"]["all_kernels"].items() if "x1" in k})
except Exception as e: print("$cfg ERR", e); print(open("gpurun_out/r4n/bench.err").read()[-600:])
P
done
for o in 3 5 9 11; do
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-configs --order $o 2>>gpurun_out/r4n/bench_o.err | tail -1 > gpurun_out/r4n/bench_o$o.json
python - <<P
import json
d=json.load(open("gpurun_out/r4n/bench_o$o.json"))
print("order", $o, d["ms_per_step"], {k: round(v["ms"],4) for k,v in d["roofline"]["all_kernels"].items() if "x1" in k})
P
done
