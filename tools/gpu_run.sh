mkdir -p gpurun_out
python -m pytest tests/test_gemm_gpu.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -15
ITERS=10 python tools/bench_gemm.py 2>&1 | tail -5
PNB_NO_FAST_EPILOGUE=1 ITERS=5 python tools/bench_gemm.py 2>&1 | tail -4
