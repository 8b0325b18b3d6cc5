# launch list of one eager training step (after the plain command exits 0) + the IPE kernel tests
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_guard_bands_gpu.py -x -q -m gpu 2>&1 | tail -3
CMD="python bench.py --steps 2 --warmup 3 --preheat 0 --no-cpu-baseline --no-extras --no-graph"
timeout 300 $CMD > gpurun_out/plain_step.log 2>&1 || { echo "plain step run failed"; tail -5 gpurun_out/plain_step.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/step_launches.csv $CMD > gpurun_out/ncu_step.log 2>&1
echo "step list rc $?"
python tools/summarize_launches.py gpurun_out/step_launches.csv > gpurun_out/step_launches.txt 2>&1; head -70 gpurun_out/step_launches.txt
