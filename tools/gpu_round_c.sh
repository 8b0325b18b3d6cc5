mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/tests.log 2>&1
echo "pytest exit $?"; grep -E "^E  |FAILED|passed|failed|Error" gpurun_out/tests.log | head -10
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; tail -2 gpurun_out/bench.err; tail -1 gpurun_out/bench.log | cut -c1-260
timeout 300 python bench.py --workload render --steps 2 --warmup 1 > gpurun_out/render1.log 2>&1; echo "render rc $?"; tail -1 gpurun_out/render1.log | cut -c1-200
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu.log 2>&1
echo "list rc $?"
python tools/summarize_launches.py gpurun_out/launches.csv 2>/dev/null | head -22
