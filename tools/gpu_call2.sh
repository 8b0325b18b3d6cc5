#!/bin/bash
# round 2, call 2: in-kernel IPE tests, full suite, parity vs float64 truth, new bench line
mkdir -p gpurun_out
lscpu | grep -E "Model name|^CPU\(s\)" > gpurun_out/r2_host2.txt
timeout 600 python -m pytest tests/test_fused_gpu.py -q -x > gpurun_out/r2_tests_fused.log 2>&1
echo "fused pytest rc=$?"; tail -5 gpurun_out/r2_tests_fused.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_tests2.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/r2_tests2.log
python tools/ipe_repro.py --seeds 40 > gpurun_out/r2_ipe_repro2.log 2>&1; tail -1 gpurun_out/r2_ipe_repro2.log
python tools/parity_report.py > gpurun_out/r2_parity_report2.log 2>&1
echo "parity rc=$?"
timeout 900 python bench.py --steps 30 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
echo "bench rc=$?"; head -c 600 gpurun_out/r2_bench1.json; tail -3 gpurun_out/r2_bench1.err
