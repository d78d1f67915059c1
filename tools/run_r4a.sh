mkdir -p gpurun_out/r4d
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_program -s 1 -c 1 -f -o gpurun_out/r4d/prof_program python tools/prof_program.py 50 > gpurun_out/r4d/ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r4d/ncu.log
