#!/bin/bash
# BASELINE.json configs[4]: stress render 2048 x 1024 at 256 + 256 samples (and the N = 128 variant of C3)
mkdir -p gpurun_out
timeout 900 python bench.py --workload render --render-hw 1024 2048 --num-samples 256 --steps 2 --warmup 1 > gpurun_out/r02_bench_render_c5.json 2> gpurun_out/r02_bench_render_c5.err
echo "c5 rc=$?"; tail -2 gpurun_out/r02_bench_render_c5.err | cut -c1-300
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_render_c5.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['config']['rays_per_forward'], d['step_tflops'], d['roofline']['frac'], d['roofline']['whole_step']['frac'])"
timeout 600 python bench.py --workload render --num-samples 128 --steps 2 --warmup 1 > gpurun_out/r02_bench_render_n128.json 2>/dev/null
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_render_n128.json')); print(d['value'], d['ms_per_step'], d['step_tflops'], d['roofline']['whole_step']['frac'])"
timeout 600 python bench.py --workload render --steps 3 --warmup 1 > gpurun_out/r02_bench_render_n1.json 2>/dev/null
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_render_n1.json')); print(d['value'], d['ms_per_step'], d['step_tflops'], d['roofline']['whole_step']['frac'], d['clocks'])"
