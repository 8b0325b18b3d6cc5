#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_fused_ipe_exp.log
: > $L
for args in "" "--ipe" "--normals" "--normals --ipe"; do
  echo "== $args" >> $L
  python tools/bench_fused.py $args >> $L 2>&1
  PNB_FUSED_PROF=1 python tools/bench_fused.py $args 2>&1 | grep -m2 "cycles/CTA" >> $L
done
echo "== --ipe handshake-only (debug 4)" >> $L
PNB_FUSED_DEBUG=4 python tools/bench_fused.py --ipe >> $L 2>&1
PNB_FUSED_DEBUG=4 PNB_FUSED_PROF=1 python tools/bench_fused.py --ipe 2>&1 | grep -m2 "cycles/CTA" >> $L
cat $L
timeout 900 python bench.py --steps 30 --warmup 3 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err
echo "bench rc=$?"; python -c "
import json
b=json.load(open('gpurun_out/r2_bench2.json')); print(b['value'], b['ms_per_step'], b['c4'], b['render'].get('value'), b['render'].get('error'))"
