#!/bin/bash
# 8 GPUs, end of round 2: the driver's launch line at N = 8 and N = 1 on the same box
mkdir -p gpurun_out
n=8
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 100 --warmup 3 > gpurun_out/r02_bench_train_n$n.raw 2> gpurun_out/r02_bench_train_n$n.err
echo "bench n$n rc=$?"
grep '^{' gpurun_out/r02_bench_train_n$n.raw | tail -1 > gpurun_out/r02_bench_train_n$n.json
python -c "
import json
b=json.load(open('gpurun_out/r02_bench_train_n$n.json')); print($n, b['value'], b['ms_per_step'], b['config'].get('update_in_graph'), b['e2e']['value'], b['clocks']); print('  c4', b['c4'].get('value'), b['c4'].get('ms_per_step'), b['c4'].get('error')); print('  render', b['render'].get('value'), b['render'].get('ms_per_step'), b['render'].get('error'))"
timeout 300 python bench.py --steps 100 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02_bench_train_n1_samebox8.json 2>/dev/null
python -c "
import json
b=json.load(open('gpurun_out/r02_bench_train_n1_samebox8.json')); print('n1 same box:', b['value'], b['ms_per_step'], b['clocks'])"
