mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/tests.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/tests.log
grep -E "^E  |FAILED|passed|failed" gpurun_out/tests.log | head -40
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-2500
timeout 600 python bench.py --workload render --steps 1 --warmup 1 > gpurun_out/render1.log 2>&1; echo "render rc $?"; tail -1 gpurun_out/render1.log | cut -c1-900
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "list rc $?"
python tools/summarize_launches.py gpurun_out/launches.csv | head -50
