mkdir -p gpurun_out
timeout 200 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -1
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-240
timeout 200 python bench.py --workload render --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
