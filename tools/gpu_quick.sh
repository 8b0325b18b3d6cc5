mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/tests.log 2>&1; echo "pytest exit $?"; grep -E "^E  |FAILED|passed|failed" gpurun_out/tests.log | head -12
timeout 200 python tools/bench_micro.py 2>&1 | tee gpurun_out/micro.log | cut -c1-175
