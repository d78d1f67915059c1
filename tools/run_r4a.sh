mkdir -p gpurun_out/r5i
timeout 600 python -m pytest tests/test_gpu_sweep.py tests/test_gpu_driver.py -q -x > gpurun_out/r5i/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r5i/pytest.log
for o in 3 5 7 9 11; do
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-configs --order $o 2>>gpurun_out/r5i/bench_o.err | tail -1 > gpurun_out/r5i/bench_o$o.json
python - <<P
import json
d=json.load(open("gpurun_out/r5i/bench_o$o.json"))
print("order", $o, d["ms_per_step"], {k: round(v["ms"],4) for k,v in d["roofline"]["all_kernels"].items() if "x1" in k})
P
done
SLB_CONTIG_TILE_BYTES=26000 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-configs 2>>gpurun_out/r5i/bench_o.err | tail -1 > gpurun_out/r5i/bench_lt24.json
python - <<P
import json
d=json.load(open("gpurun_out/r5i/bench_lt24.json"))
print("LT16->? bytes 26000", d["ms_per_step"], {k: round(v["ms"],4) for k,v in d["roofline"]["all_kernels"].items() if "x1" in k})
P
