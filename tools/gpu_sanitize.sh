#!/bin/bash
# one compute-sanitizer tool per call (B200_PROFILING.md): $1 = memcheck | racecheck | initcheck
mkdir -p gpurun_out
TOOL=$1
timeout 1500 compute-sanitizer --tool $TOOL --error-exitcode 0 --print-limit 20 \
  python -m pytest tests/test_kernels_gpu.py tests/test_image_gpu.py -q -x -k "ipe or resample or composite or sample_cast or raygen or metrics or writers or tonemap or normals or env_cast or activations" \
  > gpurun_out/r2_sanitizer_$TOOL.log 2>&1
echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY|Error|hazard" gpurun_out/r2_sanitizer_$TOOL.log | head -20
