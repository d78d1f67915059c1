#!/bin/bash
# Multi-GPU session: sharded parity (vs the single-GPU driver) and the scaling bench.
# gpurun --gpus N --timeout 900 -- 'bash tools/gpu_multi.sh <tag> <N> [check] [bench] [nccl]'
TAG=$1; N=$2; shift 2
OUT=gpurun_out/$TAG; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for w in "$@"; do
  case $w in
    tests) timeout 600 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_pair.py -x -q 2>&1 | tail -3 ;;
    check)
      timeout 300 $TR tools/check_sharded.py 32 p2p 2>&1 | grep -E "rank|Error|error" | tee $OUT/check_p2p.log
      timeout 300 $TR tools/check_sharded.py 32 nccl 2>&1 | grep -E "rank|Error|error" | tee $OUT/check_nccl.log ;;
    bspline)
      timeout 300 $TR tools/check_sharded.py 32 p2p bspline_fft 11 2>&1 | grep -E "rank|Error|error" | tee $OUT/check_bspline_p2p.log
      timeout 300 $TR tools/check_sharded.py 32 nccl bspline_lu 5 2>&1 | grep -E "rank|Error|error" | tee $OUT/check_bspline_nccl.log
      timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --interp bspline_fft --order 11 2>&1 | grep -E "^\{|Error|error" | tee $OUT/bench_${N}gpu_bspline11.json ;;
    c5)   # BASELINE config 5: 2D2V 256^4 (34 GB), Lagrange 9, strong scaling on 8 GPUs
      timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 --size 256 --order 9 --no-e2e 2>&1 | grep -E "^\{|Error|error|Killed" | tee $OUT/bench_${N}gpu_c5_256_L9.json ;;
    c4)   # BASELINE config 4: 2D2V 128^4, B-spline FFT 11
      timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --interp bspline_fft --order 11 2>&1 | grep -E "^\{|Error|error" | tee $OUT/bench_${N}gpu_c4_bspline11.json ;;
    bench)
      timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 2>&1 | grep -E "^\{|Error|error" | tee $OUT/bench_${N}gpu_p2p.json ;;
    nccl)
      timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --exchange nccl 2>&1 | grep -E "^\{|Error|error" | tee $OUT/bench_${N}gpu_nccl.json ;;
    unfused)
      SLB_FUSE=0 timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 2>&1 | grep -E "^\{|Error|error" | tee $OUT/bench_${N}gpu_p2p_unfused.json ;;
  esac
done
