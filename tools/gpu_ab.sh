python -c "import torch"
timeout 100 python -m pytest tests/test_fused_gpu.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -1
for mode in direct tma; do
for fl in "--save" "--normals --save" "--bwd" "--jadj"; do
  echo "== $mode $fl"; PNB_FUSED_SAVE=$mode PNB_FUSED_PROF=1 timeout 120 python tools/bench_fused.py $fl 2>&1 | grep -E "cycles/CTA|kernel" | tail -2 | cut -c1-200
done
done
