#!/bin/bash
# Round-2 regression call: smoke, full GPU suite, micro-benchmarks, bench line.
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py --smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke.log | cut -c1-300
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_tests_full.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r2_tests_full.log | cut -c1-300
ls gpurun_out/ipe_cpu_oracle_outlier_* 2>/dev/null
timeout 300 python tools/bench_micro.py > gpurun_out/r02_micro_memory_bound_kernels.jsonl 2>gpurun_out/r2_micro.err
python -c "
import json
for l in open('gpurun_out/r02_micro_memory_bound_kernels.jsonl'):
    d=json.loads(l); print(d['kernel'][:40], round(d['ms'],4), round(d['frac_of_hbm_roofline'],3))"
timeout 900 python bench.py --steps 150 --warmup 3 > gpurun_out/r02_bench_train_n1.json 2> gpurun_out/r2_bench_final.err
echo "bench rc=$?"; python -c "
import json
b=json.load(open('gpurun_out/r02_bench_train_n1.json')); print(b['value'], b['ms_per_step'], b['e2e']['value'], b['clocks']); print(b['c4'].get('ms_per_step'), b['c4'].get('value')); r=b['render']; print(r.get('value'), r.get('ms_per_step'), r.get('gpu_launches'), r['e2e']['value'], r.get('error'))
ro=b['roofline']; print(ro['kernel'], ro['frac'], ro['mlp_stage']['frac'], ro['whole_step']['frac'], {k:(round(v['frac'],3), round(v['kernel_ms_per_step'],3)) for k,v in ro['programs'].items()}, {k:(round(v['frac'],3), round(v['kernel_ms_per_step'],3)) for k,v in ro['other_kernels'].items()}); print(r['roofline']['frac'], r['roofline']['whole_step']['frac'])"
