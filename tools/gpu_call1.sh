#!/bin/bash
# round 2, call 1: IPE outlier hunt, full GPU suite, parity measurements, baseline bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2_gpu.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/r2_gpu.txt
python tools/ipe_repro.py --seeds 80 --threads 0 4 > gpurun_out/r2_ipe_repro.log 2>&1
echo "ipe_repro rc=$?"
python -m pytest tests -m gpu -q -x --deselect tests/test_kernels_gpu.py::test_resample_seeded > gpurun_out/r2_tests1.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2_tests1.log
python tools/parity_report.py > gpurun_out/r2_parity_report.log 2>&1
echo "parity rc=$?"; tail -3 gpurun_out/r2_parity_report.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench0.json 2> gpurun_out/r2_bench0.err
echo "bench rc=$?"; cat gpurun_out/r2_bench0.json | head -c 1500
