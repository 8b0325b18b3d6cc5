#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_fused_2cta_ab.log; : > $L
for mode in 0 1; do
  export PNB_FUSED_2CTA=$mode
  for args in "" "--normals" "--save" "--normals --save" "--bwd" "--jadj"; do
    echo "== 2cta=$mode $args" >> $L
    timeout 120 python tools/bench_fused.py $args 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['kernel'], round(d['ms'],4), round(d['tflops'],1), [round(x,3) for x in d['all_ms']])" >> $L
  done
  PNB_FUSED_PROF=1 timeout 120 python tools/bench_fused.py 2>&1 | grep -m1 "cycles/CTA" >> $L
  PNB_FUSED_PROF=1 timeout 120 python tools/bench_fused.py --save 2>&1 | grep -m1 "cycles/CTA" >> $L
done
unset PNB_FUSED_2CTA
cat $L
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_tests9.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r2_tests9.log
timeout 900 python bench.py --steps 30 --warmup 3 > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err
echo "bench rc=$?"; python -c "
import json
b=json.load(open('gpurun_out/r2_bench4.json')); print(b['value'], b['ms_per_step'], b['e2e']['value']); print(b['c4'].get('ms_per_step')); r=b['render']; print(r.get('value'), r.get('ms_per_step'), r.get('gpu_launches'), r.get('error'))
ro=b['roofline']; print(ro['kernel'], ro['frac'], ro['mlp_stage']['frac'], ro['whole_step']['frac'], {k:(round(v['frac'],3), round(v['kernel_ms_per_step'],3)) for k,v in ro['programs'].items()}, {k:(round(v['frac'],3), round(v['kernel_ms_per_step'],3)) for k,v in ro['other_kernels'].items()})"
PNB_FUSED_2CTA=0 timeout 900 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench4_1cta.json 2> /dev/null
python -c "
import json
b=json.load(open('gpurun_out/r2_bench4_1cta.json')); print('1cta:', b['value'], b['ms_per_step']); r=b['render']; print(r.get('value'), r.get('ms_per_step'))"
