python -c "import torch"
for i in 1 2 3; do timeout 40 python tools/stress_fused.py bwd 4096 500 2>&1 | tail -1; echo "rc ${PIPESTATUS[0]}"; done
for i in 1 2; do
timeout 100 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | grep -E "^E  |passed|failed|FAILED" | head -5
echo "rc ${PIPESTATUS[0]}"
done
