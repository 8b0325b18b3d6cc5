mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_fused_gpu.py tests/test_models_gpu.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -1
for fl in "" "--save" "--normals --save" "--bwd" "--jadj"; do timeout 100 python tools/bench_fused.py $fl 2>&1 | tail -1 | cut -c1-130; done
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-240
