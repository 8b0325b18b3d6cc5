mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_fused_gpu.py -x -q -m gpu > gpurun_out/fused_tests.log 2>&1; echo "fused tests rc $?"; tail -15 gpurun_out/fused_tests.log
timeout 120 python tools/bench_fused.py > gpurun_out/fused_bench.log 2>&1; echo "bench rc $?"; tail -2 gpurun_out/fused_bench.log
timeout 120 python tools/bench_fused.py --normals > gpurun_out/fused_bench_n.log 2>&1; echo "bench rc $?"; tail -2 gpurun_out/fused_bench_n.log
timeout 120 python tools/bench_fused.py --normals --save > gpurun_out/fused_bench_ns.log 2>&1; echo "bench rc $?"; tail -2 gpurun_out/fused_bench_ns.log
