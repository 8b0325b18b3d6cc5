mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do timeout 300 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -1; done
