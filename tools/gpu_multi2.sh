#!/bin/bash
# 2 GPUs: data-parallel gradient equivalence, NCCL all-reduce + Adam inside the CUDA graph, scaling lines
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dropin_gpu.py tests/test_guard_bands_gpu.py -q > gpurun_out/r2_tests_multi.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r2_tests_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
echo "bench n2 rc=$?"; python -c "
import json
b=json.load(open('gpurun_out/r2_bench_n2.json')); print(b['value'], b['ms_per_step'], b['config'].get('update_in_graph'), b['e2e']['value']); print(b['c4']); print(b['render'].get('value'), b['render'].get('ms_per_step'), b['render'].get('error'))"
tail -3 gpurun_out/r2_bench_n2.err
PNB_GRAPH_NCCL=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 40 --warmup 3 --no-extras > gpurun_out/r2_bench_n2_nccl_outside.json 2> gpurun_out/r2_bench_n2_nccl_outside.err
python -c "
import json
b=json.load(open('gpurun_out/r2_bench_n2_nccl_outside.json')); print('nccl outside graph:', b['value'], b['ms_per_step'], b['config'].get('update_in_graph'))"
timeout 600 python bench.py --steps 40 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_n1_samebox.json 2>/dev/null
python -c "
import json
b=json.load(open('gpurun_out/r2_bench_n1_samebox.json')); print('n1 same box:', b['value'], b['ms_per_step'])"
