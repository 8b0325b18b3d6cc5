mkdir -p gpurun_out
for rep in 1 4 16; do
for fl in "" "--jadj"; do
  echo "== rep=$rep $fl"
  PNB_FUSED_REPLICAS=$rep PNB_FUSED_PROF=1 timeout 120 python tools/bench_fused.py $fl 2>&1 | grep -E "cycles/CTA|kernel" | tail -2 | cut -c1-250
done
done
for fl in "--normals" "--save" "--normals --save" "--bwd"; do
  echo "== $fl"
  PNB_FUSED_PROF=1 timeout 120 python tools/bench_fused.py $fl 2>&1 | grep -E "cycles/CTA|kernel" | tail -2 | cut -c1-250
done
echo "== debug 2"; PNB_FUSED_DEBUG=2 PNB_FUSED_PROF=1 timeout 120 python tools/bench_fused.py 2>&1 | grep -E "cycles/CTA|kernel" | tail -2 | cut -c1-250
