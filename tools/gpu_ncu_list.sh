# Per-launch device times of one (eager) training step: the launch list the round evidence is built from.
mkdir -p gpurun_out
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu.log 2>&1
echo "list rc $?"
python tools/summarize_launches.py gpurun_out/launches.csv 2>/dev/null | head -12
