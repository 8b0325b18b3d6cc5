#!/bin/bash
mkdir -p gpurun_out
lscpu | grep -E "Model name|^CPU\(s\)|Hypervisor|L3" > gpurun_out/r2_host_micro.txt
python tools/ipe_repro.py --seeds 60 > gpurun_out/r2_ipe_repro_micro.log 2>&1; tail -1 gpurun_out/r2_ipe_repro_micro.log | cut -c1-300
for i in 1 2 3; do timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k ipe > gpurun_out/r2_tests_ipe_$i.log 2>&1; echo "ipe test run $i rc=$?"; done
ls gpurun_out/ipe_cpu_oracle_outlier_* 2>/dev/null
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_config_size_gpu.py tests/test_guard_bands_gpu.py tests/test_fused_gpu.py tests/test_image_gpu.py tests/test_models_gpu.py -q > gpurun_out/r2_tests_micro.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2_tests_micro.log | cut -c1-300
timeout 300 python tools/bench_micro.py 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ', d['kernel'][:40], round(d['ms'],4), round(d['frac_of_hbm_roofline'],3))"
