mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_resample_grad_gpu.py -x -q -m gpu > gpurun_out/tests_one.log 2>&1; echo "pytest exit $?"; grep -E "^E  |FAILED|passed|failed|Error|error" gpurun_out/tests_one.log | head -20
