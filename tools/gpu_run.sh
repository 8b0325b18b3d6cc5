mkdir -p gpurun_out
python -m pytest tests/test_models_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "tensor_core" 2>&1 | grep -E "^E  |passed|failed" | head
