#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fused_gpu.py -q -x > gpurun_out/r2_tests_fused6.log 2>&1
echo "fused pytest rc=$?"; tail -12 gpurun_out/r2_tests_fused6.log
L=gpurun_out/r2_fused_bias_exp.log; : > $L
for args in "" "--normals" "--save" "--normals --save" "--bwd" "--jadj"; do
  echo "== $args" >> $L
  timeout 120 python tools/bench_fused.py $args 2>&1 | tail -1 | cut -c1-220 >> $L
done
cat $L
timeout 600 python tools/debug_c4.py graphed eager 2>&1 | tail -4
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_tests6.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/r2_tests6.log
