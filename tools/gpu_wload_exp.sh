#!/bin/bash
# A/B: weight tiles through the TMA unit (default) vs LDGSTS (PNB_FUSED_DEBUG=4)
mkdir -p gpurun_out
L=gpurun_out/fused_wload_ab.log; : > $L
for dbg in 0 4 0 4; do
  export PNB_FUSED_DEBUG=$dbg
  if [ $dbg = 4 ]; then timeout 300 python -m pytest tests/test_fused_gpu.py -q -x 2>&1 | tail -1 >> $L; fi
  for args in "" "--normals" "--save" "--normals --save" "--bwd" "--jadj"; do
    echo "== debug=$dbg $args" >> $L
    timeout 120 python tools/bench_fused.py $args 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['kernel'], round(d['ms'],4), round(d['tflops'],1))" >> $L
  done
  PNB_FUSED_PROF=1 timeout 120 python tools/bench_fused.py --save 2>&1 | grep -m1 "cycles/CTA" >> $L
done
cat $L
