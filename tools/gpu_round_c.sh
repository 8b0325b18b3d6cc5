mkdir -p gpurun_out
timeout 240 python -m pytest tests -x -q -m gpu > gpurun_out/tests.log 2>&1
echo "pytest exit $?"; grep -E "^E  |FAILED|passed|failed|Error" gpurun_out/tests.log | head -10
for fl in "--normals" "--normals --save" "--bwd"; do timeout 100 python tools/bench_fused.py $fl 2>&1 | tail -1 | cut -c1-130; done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; tail -2 gpurun_out/bench.err; tail -1 gpurun_out/bench.log | cut -c1-260
timeout 300 python bench.py --workload render --steps 2 --warmup 1 > gpurun_out/render1.log 2>&1; echo "render rc $?"; tail -1 gpurun_out/render1.log | cut -c1-200
