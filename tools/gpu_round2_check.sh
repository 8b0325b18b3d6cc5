#!/bin/bash
# Round-2 regression call: full GPU suite, memory-bound micro-benchmarks, fused-kernel micro-benchmarks, bench line.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_tests_full.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r2_tests_full.log
timeout 300 python tools/bench_micro.py > gpurun_out/r02_micro_memory_bound_kernels.jsonl 2>gpurun_out/r2_micro.err
python -c "
import json
for l in open('gpurun_out/r02_micro_memory_bound_kernels.jsonl'):
    d=json.loads(l); print(d['kernel'], round(d['ms'],4), round(d['frac_of_hbm_roofline'],3))"
: > gpurun_out/r02_fused_mlp_microbench.jsonl
for args in "" "--normals" "--save" "--normals --save" "--bwd" "--jadj"; do
  timeout 120 python tools/bench_fused.py $args 2>/dev/null | tail -1 >> gpurun_out/r02_fused_mlp_microbench.jsonl
done
timeout 900 python bench.py --steps 150 --warmup 3 > gpurun_out/r02_bench_train_n1.json 2> gpurun_out/r2_bench_final.err
echo "bench rc=$?"; python -c "
import json
b=json.load(open('gpurun_out/r02_bench_train_n1.json')); print(b['value'], b['ms_per_step'], b['e2e']['value'], b['clocks']); print(b['c4']); r=b['render']; print(r.get('value'), r.get('ms_per_step'), r.get('gpu_launches'), r.get('error'))
ro=b['roofline']; print(ro['kernel'], ro['frac'], ro['mlp_stage']['frac'], ro['whole_step']['frac'], {k:(round(v['frac'],3), round(v['kernel_ms_per_step'],3)) for k,v in ro['programs'].items()}, {k:(round(v['frac'],3), round(v['kernel_ms_per_step'],3)) for k,v in ro['other_kernels'].items()}); print(b['cpu_baseline'])"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null; cat gpurun_out/r02_bench_reference_arm.json | cut -c1-300
