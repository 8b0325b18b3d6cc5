# IPE gradient tile kernels: parity + micro-benchmark (tile vs per-feature), then the full GPU suite and the bench.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_guard_bands_gpu.py -x -q -m gpu 2>&1 | tail -5
echo "== tile"; timeout 200 python tools/bench_micro.py 2>&1 | cut -c1-170 | tee gpurun_out/micro_tile.log
echo "== per-feature"; PNB_IPE_SLOW=1 timeout 200 python tools/bench_micro.py 2>&1 | grep "vjp\|jvp" | cut -c1-170 | tee gpurun_out/micro_slow.log
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/tests.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/tests.log
timeout 300 python bench.py > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "train rc $?"; tail -1 gpurun_out/bench_train.json | cut -c1-400
