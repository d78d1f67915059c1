#!/bin/bash
# One GPU-box session: parity tests, bench (both arms), ncu launch list and full captures of the
# sweep kernels.  Run as:  gpurun --timeout 1500 -- 'bash tools/gpu_eval.sh [tag] [what...]'
# Outputs under gpurun_out/<tag>/ ; summaries are copied into profiles/ by tools/summarize_ncu.py.
TAG=${1:-eval}
shift
WHAT=${*:-tests bench ref launches full}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,memory.total --format=csv > "$OUT/smi.csv" 2>&1
for w in $WHAT; do
  case $w in
    tests)
      timeout 1200 python -m pytest tests -m gpu -q > "$OUT/pytest_gpu.log" 2>&1
      echo "pytest rc=$?"; tail -3 "$OUT/pytest_gpu.log" ;;
    bench)
      timeout 600 python bench.py --steps 10 --warmup 3 > "$OUT/bench.json" 2> "$OUT/bench.err"
      echo "bench rc=$?"; cat "$OUT/bench.json" ;;
    ref)
      timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"
      echo "ref rc=$?"; cat "$OUT/bench_ref.json" ;;
    launches)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
        --log-file "$OUT/launches.csv" python bench.py --steps 2 --warmup 3 --no-cpu > "$OUT/launches.log" 2>&1
      echo "launches rc=$?" ;;
    full)
      # skip the warm-up step's launches; 1 capture each of the strided and contiguous sweep kernels
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep_fused\|k_charge_partial -s 4 -c 5 \
        -f -o "$OUT/prof_sweep" python tools/prof_step.py --steps 1 > "$OUT/full.log" 2>&1
      echo "full rc=$?" ;;
    bspline)
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sweep|k_bspline' -s 12 -c 12 \
        -f -o "$OUT/prof_bspline" python tools/prof_step.py --steps 1 --interp bspline_fft --order 11 > "$OUT/full_bspline.log" 2>&1
      echo "bspline rc=$?"
      timeout 600 python bench.py --steps 5 --warmup 3 --interp bspline_fft --order 11 --no-cpu > "$OUT/bench_bspline.json" 2> "$OUT/bench_bspline.err"
      cat "$OUT/bench_bspline.json" ;;
    bsp)
      for cfg in "bspline_fft 11" "bspline_lu 5" "bspline_lu 3"; do
        set -- $cfg
        timeout 300 python bench.py --steps 5 --warmup 3 --interp $1 --order $2 --no-cpu --no-configs 2>>"$OUT/bench_bsp.err" | tail -1 >> "$OUT/bench_bsp.jsonl"
      done
      python - <<PY
import json
for ln in open("$OUT/bench_bsp.jsonl"):
    d = json.loads(ln); k = d["roofline"]["all_kernels"]; c = d["config"]
    print(c["interp"], c["order"], "ms/step %.3f" % d["ms_per_step"], "Gcell/s %.1f" % d["value"], {n: round(v["ms"], 3) for n, v in k.items()})
PY
      ;;
    bsptest)
      timeout 300 python -m pytest tests/test_gpu_sweep.py tests/test_gpu_driver.py -q -x -k "bspline or spline" > "$OUT/pytest_bsp.log" 2>&1
      echo "bsptest rc=$?"; tail -5 "$OUT/pytest_bsp.log" ;;
    bspab)
      for split in ${BSPAB_MODES:-1 0}; do
        for cfg in "bspline_fft 11" "bspline_lu 5" "bspline_lu 3"; do
          set -- $cfg
          SLB_BSPLINE_RF=$split timeout 300 python bench.py --steps 5 --warmup 3 --interp $1 --order $2 --no-cpu --no-configs 2>>"$OUT/bench_bspab.err" | tail -1 > "$OUT/tmp.json"
          python - <<PY
import json
d = json.load(open("$OUT/tmp.json")); k = d["roofline"]["all_kernels"]; c = d["config"]
print("rf=$split", c["interp"], c["order"], "ms/step %.3f" % d["ms_per_step"], "Gcell/s %.1f" % d["value"], {n.split("/")[1]: round(v["ms"], 3) for n, v in k.items() if "fused" not in n})
PY
          cat "$OUT/tmp.json" >> "$OUT/bench_bspab_rf$split.jsonl"
        done
      done ;;
    ncusplit)
      timeout 400 ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-k_bspline_fused} -s 3 -c 1 \
        -f -o "$OUT/prof_split" python tools/prof_step.py --steps 1 --interp bspline_fft --order 11 > "$OUT/ncusplit.log" 2>&1
      echo "ncusplit rc=$?"; tail -3 "$OUT/ncusplit.log"; ls -la "$OUT" ;;
    configs)
      timeout 300 python tools/bench_configs.py > "$OUT/bench_configs.json" 2> "$OUT/bench_configs.err"
      echo "configs rc=$?"; cat "$OUT/bench_configs.json"; tail -3 "$OUT/bench_configs.err" ;;
    points)
      timeout 300 python tools/bench_points.py > "$OUT/bench_points.json" 2> "$OUT/bench_points.err"
      echo "points rc=$?"; cat "$OUT/bench_points.json"; tail -3 "$OUT/bench_points.err" ;;
    smoke)
      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1
      echo "smoke rc=$?"; tail -2 "$OUT/smoke.log" ;;
    *)
      echo "unknown item $w" ;;
  esac
done
ls -la "$OUT"
