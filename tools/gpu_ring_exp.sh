#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_fused_ring_modes.log; : > $L
for mode in 2 0 1; do
  if [ $mode = 2 ]; then unset PNB_LIB_PATH; else export PNB_LIB_PATH=$PWD/panonerf_b200/libpanonerf_b200_ring$mode.so; fi
  for args in "" "--save" "--normals --save" "--bwd" "--jadj"; do
    echo "== ring=$mode $args" >> $L
    timeout 120 python tools/bench_fused.py $args 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['kernel'], round(d['ms'],4), round(d['tflops'],1))" >> $L
  done
  PNB_FUSED_PROF=1 timeout 120 python tools/bench_fused.py --save 2>&1 | grep -m1 "cycles/CTA" >> $L
done
cat $L
