#!/bin/bash
mkdir -p gpurun_out
lscpu | grep -E "Model name|^CPU\(s\)" > gpurun_out/r2_host3.txt
timeout 300 python tools/debug_c4.py 4096 > gpurun_out/r2_c4_debug.log 2>&1; echo "c4 debug rc=$?"; tail -25 gpurun_out/r2_c4_debug.log
timeout 300 python tools/debug_c4.py 4096 --with-graphed > gpurun_out/r2_c4_debug2.log 2>&1; echo "c4 debug2 rc=$?"; tail -12 gpurun_out/r2_c4_debug2.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_tests3.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r2_tests3.log
python tools/ipe_repro.py --seeds 40 > gpurun_out/r2_ipe_repro3.log 2>&1; tail -1 gpurun_out/r2_ipe_repro3.log
for mode in ipe noipe; do
  for chunk in 0 32768; do
    if [ $mode = noipe ]; then export PNB_NO_FUSED_IPE=1; else unset PNB_NO_FUSED_IPE; fi
    timeout 600 python bench.py --workload render --steps 3 --warmup 1 --render-chunk $chunk > gpurun_out/r2_render_${mode}_${chunk}.json 2> gpurun_out/r2_render_${mode}_${chunk}.err
    echo "render $mode chunk=$chunk rc=$?"; python -c "
import json,sys
d=json.load(open('gpurun_out/r2_render_${mode}_${chunk}.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks'])"
  done
done
