#!/bin/bash
# compute-sanitizer over tools/sanitize_cases.py on the GPU box:  gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
# Logs land in gpurun_out/sanitize_<tool>.log; copy the summaries to profiles/.
OUT=gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_cases.py > $OUT/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize cases done" $OUT/sanitize_$tool.log | tail -3
done
