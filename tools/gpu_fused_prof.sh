mkdir -p gpurun_out
for fl in "" "--normals" "--save" "--normals --save" "--bwd" "--jadj"; do
  echo "== $fl"
  PNB_FUSED_PROF=1 timeout 120 python tools/bench_fused.py $fl 2>&1 | grep -E "cycles/CTA|kernel" | tail -2 | cut -c1-250
done
