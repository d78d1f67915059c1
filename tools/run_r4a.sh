mkdir -p gpurun_out/r5g
timeout 600 python -m pytest tests/test_gpu_driver.py tests/test_gpu_sweep.py -q -x > gpurun_out/r5g/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r5g/pytest.log
timeout 300 python tools/bench_c1.py 32 40 48 > gpurun_out/r5g/c1.json 2> gpurun_out/r5g/c1.err; echo "c1 rc=$?"; python - <<P
import json
d=json.load(open("gpurun_out/r5g/c1.json"))
for k,v in d.items():
    print(k, v["ms_per_step"], v["ms_per_step_graph"], v["ms_per_step_program"], v["program_equals_stepwise_bitwise"], v["program"]["block0_ns_per_op_kind_wait_run"][:7] if v.get("program") else v)
P
tail -3 gpurun_out/r5g/c1.err
for o in 3 5 7; do
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-configs --order $o 2>>gpurun_out/r5g/bench_o.err | tail -1 > gpurun_out/r5g/bench_o$o.json
python - <<P
import json
d=json.load(open("gpurun_out/r5g/bench_o$o.json"))
print("order", $o, d["ms_per_step"], {k: round(v["ms"],4) for k,v in d["roofline"]["all_kernels"].items() if "x1" in k})
P
done
