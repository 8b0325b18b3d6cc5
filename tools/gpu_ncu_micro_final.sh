#!/bin/bash
# ncu --set full on the stand-alone memory-bound kernels in their final round-2 state (tools/bench_micro.py 22)
mkdir -p gpurun_out
timeout 300 python tools/bench_micro.py 22 > gpurun_out/r02_micro_plain.log 2>&1 || { echo "plain failed"; exit 1; }
for k in ipe_fwd_tile_kernel ipe_jvp_tile_kernel ipe_vjp_kernel sample_cast_kernel composite_fwd_blocked_kernel resample_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 3 --launch-count 2 -f -o gpurun_out/r02f_micro_$k python tools/bench_micro.py 22 > gpurun_out/r02f_ncu_micro_$k.log 2>&1
  echo "ncu $k rc $?"
done
ls -la gpurun_out/r02f_micro_*.ncu-rep | awk '{print $5, $9}'
