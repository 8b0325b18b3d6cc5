mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/tests.log 2>&1; echo "pytest exit $?"; grep -E "^E  |FAILED|passed|failed" gpurun_out/tests.log | head -12
for fl in "" "--normals" "--save" "--normals --save" "--bwd" "--jadj"; do timeout 120 python tools/bench_fused.py $fl 2>&1 | tail -1 | cut -c1-140; done
for rep in 1 2; do timeout 300 python bench.py --steps 100 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['clocks'])"; done
