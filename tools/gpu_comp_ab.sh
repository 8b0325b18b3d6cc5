#!/bin/bash
# A/B of the blocked compositing layouts (K samples per lane x L lanes per ray) + bench.py C1 extra smoke
mkdir -p gpurun_out
L=gpurun_out/composite_ab.log; : > $L
for v in "" 1; do
  if [ -n "$v" ]; then export PNB_COMPOSITE_K8=1; else unset PNB_COMPOSITE_K8; fi
  echo "== K8=$v" >> $L
  for i in 1 2 3; do timeout 120 python tools/bench_micro.py 2>&1 | grep composite_fwd | cut -c1-130 >> $L; done
  timeout 200 python -m pytest tests/test_kernels_gpu.py -q -x -m gpu -k "composite" 2>&1 | tail -1 >> $L
done
cat $L
