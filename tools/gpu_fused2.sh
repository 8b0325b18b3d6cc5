mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_fused_gpu.py -x -q -m gpu -p no:cacheprovider > gpurun_out/fused_tests.log 2>&1; echo "fused tests rc $?"; tail -25 gpurun_out/fused_tests.log
for fl in "" "--normals" "--save" "--normals --save" "--bwd" "--jadj"; do
  echo "== $fl"
  timeout 120 python tools/bench_fused.py $fl 2>&1 | tail -1 | cut -c1-260
done
for dbg in 1 2; do echo "== debug $dbg"; PNB_FUSED_DEBUG=$dbg timeout 120 python tools/bench_fused.py 2>&1 | tail -1 | cut -c1-200; done
