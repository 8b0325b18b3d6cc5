mkdir -p gpurun_out
( time timeout 600 python bench.py > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err ) 2>&1 | grep real; echo "train rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_train.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'])
print(json.dumps(d.get('micro_kernels'), indent=0)[:1500])
print({k: d[k].get('ms_per_step') for k in ('c4','c1','render')})
PY
tail -3 gpurun_out/bench_train.err
