# Round evidence: bench lines (train / render / reference arm), per-launch device times of one training step and one
# render chunk, micro-benchmarks of the memory-bound stages, fused-kernel micro-benchmarks.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total,driver_version --format=csv > gpurun_out/gpu.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "train rc $?"; tail -1 gpurun_out/bench_train.json | cut -c1-400
timeout 600 python bench.py --workload render --steps 2 --warmup 1 > gpurun_out/bench_render.json 2> gpurun_out/bench_render.err; echo "render rc $?"; tail -1 gpurun_out/bench_render.json | cut -c1-400
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc $?"; tail -1 gpurun_out/bench_reference.json | cut -c1-300
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "train list rc $?"
timeout 300 python bench.py --workload render --render-hw 128 256 --steps 1 --warmup 1 > gpurun_out/plain_r.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/launches_render.csv \
    python bench.py --workload render --render-hw 128 256 --steps 1 --warmup 1 > gpurun_out/ncu_r.log 2>&1
echo "render list rc $?"
timeout 120 python tools/bench_micro.py > gpurun_out/micro.log 2>&1; echo "micro rc $?"
for fl in "" "--normals" "--save" "--normals --save" "--bwd" "--jadj"; do timeout 120 python tools/bench_fused.py $fl 2>&1 | tail -1; done > gpurun_out/fused_micro.log
cat gpurun_out/fused_micro.log | cut -c1-150
