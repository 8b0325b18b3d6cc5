#!/bin/bash
# ncu --set full on the stand-alone memory-bound kernels (tools/bench_micro.py) - stall reasons and DRAM bytes
mkdir -p gpurun_out
timeout 300 python tools/bench_micro.py 22 > gpurun_out/r02_micro_plain.log 2>&1 || { echo "plain failed"; exit 1; }
for k in sample_cast_kernel composite_fwd_kernel resample_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 5 --launch-count 1 -f -o gpurun_out/r02_prof_micro_$k python tools/bench_micro.py 22 > gpurun_out/r02_ncu_micro_$k.log 2>&1
  echo "ncu $k rc $?"
done
ls -la gpurun_out/r02_prof_micro_*.ncu-rep
