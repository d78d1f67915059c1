mkdir -p gpurun_out/r5h
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bspline_seg -s 6 -c 6 -f -o gpurun_out/r5h/prof_bspline_seg python tools/prof_step.py --steps 1 --interp bspline_fft --order 11 > gpurun_out/r5h/ncu_bsp.log 2>&1; echo "ncu bsp rc=$?"; tail -2 gpurun_out/r5h/ncu_bsp.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep_contig -s 2 -c 1 -f -o gpurun_out/r5h/prof_contig_tile python tools/prof_sweep_x1.py 7 > gpurun_out/r5h/ncu_tile.log 2>&1; echo "ncu tile rc=$?"; tail -2 gpurun_out/r5h/ncu_tile.log
