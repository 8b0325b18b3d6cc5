mkdir -p gpurun_out
timeout 400 python bench.py --steps 60 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "train rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_train.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['final_loss'], d['clocks'])
print({k:(d[k].get('ms_per_step'), d[k].get('error')) for k in ('c4','c1','render')}, d['cpu_baseline']['value'])
PY
tail -2 gpurun_out/bench_train.err
timeout 200 python bench.py --workload render --steps 1 --warmup 1 2>/dev/null | cut -c1-120
