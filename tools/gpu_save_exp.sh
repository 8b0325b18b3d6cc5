#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_fused_save_modes.log; : > $L
for mode in ${MODES:-t w}; do
  export PNB_FUSED_SAVE=$mode
  timeout 300 python -m pytest tests/test_fused_gpu.py -q -x 2>&1 | tail -1 >> $L
  for args in "--save" "--normals --save" "--bwd" "--jadj"; do
    echo "== save=$mode $args" >> $L
    timeout 120 python tools/bench_fused.py $args 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['kernel'], round(d['ms'],4), round(d['tflops'],1))" >> $L
  done
  PNB_FUSED_PROF=1 timeout 120 python tools/bench_fused.py --save 2>&1 | grep -m1 "cycles/CTA" >> $L
done
cat $L
