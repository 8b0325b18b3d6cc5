mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_fused_gpu.py -x -q -m gpu > gpurun_out/fused_tests.log 2>&1; echo "fused tests rc $?"; tail -5 gpurun_out/fused_tests.log
for dbg in 0 1 2 3; do
  for fl in "" "--normals"; do
    echo "debug=$dbg $fl"
    PNB_FUSED_DEBUG=$dbg timeout 120 python tools/bench_fused.py $fl 2>&1 | tail -1 | cut -c1-200
  done
done
