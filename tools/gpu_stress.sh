python -c "import torch"
export PNB_LIB_PATH=$PWD/panonerf_b200/build/libmode2.so
for w in fwd fwdsave fwdj bwd jadj; do timeout 40 python tools/stress_fused.py $w 4096 500 2>&1 | tail -1; done
for i in 1 2; do
timeout 100 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | grep -E "^E  |passed|failed|FAILED" | head -5
echo "rc ${PIPESTATUS[0]}"
done
for fl in "" "--normals" "--save" "--normals --save" "--bwd" "--jadj"; do
  echo "== mode2 $fl"; timeout 120 python tools/bench_fused.py $fl 2>&1 | tail -1 | cut -c1-120
done
unset PNB_LIB_PATH
for fl in "" "--normals" "--save" "--normals --save" "--bwd" "--jadj"; do
  echo "== mode0 $fl"; timeout 120 python tools/bench_fused.py $fl 2>&1 | tail -1 | cut -c1-120
done
