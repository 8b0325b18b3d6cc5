#!/bin/bash
# 2 GPUs, end of round 2: the driver's launch line for N = 2 (both arms) and N = 1 on the same box
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 60 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
echo "bench n2 rc=$?"; python -c "
import json
b=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1]); print(b['value'], b['ms_per_step'], b['config'].get('update_in_graph'), b['e2e']['value']); print(b['c4'].get('value'), b['c4'].get('ms_per_step')); print(b['render'].get('value'), b['render'].get('ms_per_step'), b['render'].get('error')); print(sorted(b.keys()))"
tail -3 gpurun_out/r2_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
timeout 300 python bench.py --steps 60 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_n1_samebox.json 2>/dev/null
python -c "
import json
b=json.loads(open('gpurun_out/r2_bench_n1_samebox.json').read().strip().splitlines()[-1]); print('n1 same box:', b['value'], b['ms_per_step'])"
