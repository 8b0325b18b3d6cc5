python -c "import torch" 
for i in 1 2; do
timeout 200 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | grep -E "^E  |passed|failed|FAILED" | head -5
echo "rc ${PIPESTATUS[0]}"
done
timeout 200 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
