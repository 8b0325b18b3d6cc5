mkdir -p gpurun_out
timeout 240 python -m pytest tests -x -q -m gpu > gpurun_out/tests.log 2>&1
echo "pytest exit $?"; grep -E "^E  |FAILED|passed|failed|Error" gpurun_out/tests.log | head -10
for fl in "--save" "--normals --save" "--bwd" "--jadj"; do timeout 100 python tools/bench_fused.py $fl 2>&1 | tail -1 | cut -c1-130; done
for w in fwdsave fwdj bwd jadj; do timeout 40 python tools/stress_fused.py $w 4096 300 2>&1 | tail -1; timeout 40 python tools/stress_fused.py $w 770 300 2>&1 | tail -1; done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; tail -2 gpurun_out/bench.err; tail -1 gpurun_out/bench.log | cut -c1-260
