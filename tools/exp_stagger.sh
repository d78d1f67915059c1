mkdir -p gpurun_out/s4i
for st in 0 3000 6000 12000; do
  SLB_BSPLINE_STAGGER=$st timeout 200 python bench.py --steps 3 --warmup 3 --interp bspline_fft --order 11 --no-cpu 2>/dev/null | tail -1 > gpurun_out/s4i/st$st.json
  python -c "
import json; d=json.load(open('gpurun_out/s4i/st$st.json')); print('stagger=$st', {k.split('/')[1]:round(v['ms'],3) for k,v in d['roofline']['all_kernels'].items() if 'fused' not in k}, round(d['ms_per_step'],3))"
done
