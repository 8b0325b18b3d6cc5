#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_fused_bias_ab.log; : > $L
for lib in prevfused new; do
  if [ $lib = prevfused ]; then export PNB_LIB_PATH=$PWD/panonerf_b200/libpanonerf_b200_prevfused.so; else unset PNB_LIB_PATH; fi
  for args in "" "--normals" "--save" "--normals --save"; do
    echo "== $lib $args" >> $L
    timeout 120 python tools/bench_fused.py $args 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['kernel'], round(d['ms'],4), round(d['tflops'],1), [round(x,3) for x in d['all_ms']])" >> $L
  done
done
unset PNB_LIB_PATH
cat $L
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_tests7.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r2_tests7.log
timeout 900 python bench.py --steps 30 --warmup 3 > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err
echo "bench rc=$?"; python -c "
import json
b=json.load(open('gpurun_out/r2_bench3.json')); print(b['value'], b['ms_per_step'], b['e2e']['value']); print(b['c4']); r=b['render']; print(r.get('value'), r.get('ms_per_step'), r.get('gpu_launches'), r.get('error'))
ro=b['roofline']; print(ro['kernel'], ro['frac'], ro['mlp_stage']['frac'], ro['whole_step']['frac'], {k:(round(v['frac'],3), round(v['kernel_ms_per_step'],3)) for k,v in ro['programs'].items()}, {k:(round(v['frac'],3), round(v['kernel_ms_per_step'],3)) for k,v in ro['other_kernels'].items()})"
