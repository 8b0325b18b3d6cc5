#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_c4_debug3.log; : > $L
for f in "graphed" "graphed eager" "graphed eager profile" "graphed eager profile sampler" "eager" "profile"; do
  echo "=== $f" >> $L
  timeout 300 python tools/debug_c4.py $f 2>&1 | grep -v Warning | tail -8 >> $L
done
grep -E "===|ok|FAILED|Error" $L
