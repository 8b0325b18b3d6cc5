mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_resample_grad_gpu.py tests/test_models_gpu.py -x -q -m gpu -k "resample or ipe_var or cast_rays or fence or rg" > gpurun_out/tests_rg.log 2>&1; echo "pytest exit $?"; grep -E "^E  |FAILED|passed|failed|Error" gpurun_out/tests_rg.log | head -30
