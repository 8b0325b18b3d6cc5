mkdir -p gpurun_out
timeout 200 python -m pytest tests -x -q -m gpu > gpurun_out/tests.log 2>&1
echo "pytest exit $?"; grep -E "^E  |FAILED|passed|failed|Error" gpurun_out/tests.log | head -10
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err; tail -1 gpurun_out/bench.log | cut -c1-900
