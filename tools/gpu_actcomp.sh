mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/tests.log 2>&1; echo "pytest exit $?"; grep -E "^E  |FAILED|passed|failed" gpurun_out/tests.log | head -12
timeout 100 python __graft_entry__.py --smoke 2>&1 | tail -1 | cut -c1-200
for rep in 1 2; do
echo "== fused act+composite"; timeout 300 python bench.py --steps 100 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['clocks'])"
echo "== unfused"; PNB_UNFUSED_ACT=1 timeout 300 python bench.py --steps 100 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['clocks'])"
done
echo "== render fused"; timeout 300 python bench.py --workload render --steps 3 --warmup 1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d.get('gpu_launches'))"
echo "== render unfused"; PNB_UNFUSED_ACT=1 timeout 300 python bench.py --workload render --steps 3 --warmup 1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d.get('gpu_launches'))"
