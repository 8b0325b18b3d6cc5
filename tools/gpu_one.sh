mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "empty" > gpurun_out/tests_one.log 2>&1; echo "pytest exit $?"; grep -E "^E  |FAILED|passed|failed|Error|error" gpurun_out/tests_one.log | head -20; tail -30 gpurun_out/tests_one.log | grep -E "^\S+\.py:[0-9]+" | head
