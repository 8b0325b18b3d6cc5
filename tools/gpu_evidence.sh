#!/bin/bash
# Round-end evidence (B200_PROFILING.md recipe): what the driver runs (tests, smoke, bench arms), then the ncu launch
# lists of one training step and one 128x256 render, and one --set full capture of the step's tensor-core kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv,noheader
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/tests.log 2>&1
echo "pytest exit $?"; grep -E "^E  |FAILED|passed|failed" gpurun_out/tests.log | head -8
timeout 120 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log | cut -c1-200
timeout 400 python bench.py > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "train rc $?"; tail -1 gpurun_out/bench_train.json | cut -c1-200
timeout 300 python bench.py --workload render --steps 3 --warmup 1 > gpurun_out/bench_render.json 2> gpurun_out/bench_render.err; echo "render rc $?"; tail -1 gpurun_out/bench_render.json | cut -c1-200
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc $?"; tail -1 gpurun_out/bench_reference.json | cut -c1-200
timeout 120 python tools/bench_micro.py > gpurun_out/micro.log 2>&1; echo "micro rc $?"
for fl in "" "--normals" "--save" "--normals --save" "--bwd" "--jadj"; do timeout 120 python tools/bench_fused.py $fl 2>&1 | tail -1; done > gpurun_out/fused_micro.log
cut -c1-130 gpurun_out/fused_micro.log
CMD="python bench.py --steps 2 --warmup 3 --preheat 0 --no-cpu-baseline --no-extras --no-graph"
timeout 300 $CMD > gpurun_out/plain_step.log 2>&1 || { echo "plain step run failed"; tail -5 gpurun_out/plain_step.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/r02_train_step_launches.csv $CMD > gpurun_out/ncu_step.log 2>&1
echo "step list rc $?"
python tools/summarize_launches.py gpurun_out/r02_train_step_launches.csv > gpurun_out/r02_train_step_launches.txt 2>&1; head -12 gpurun_out/r02_train_step_launches.txt
RCMD="python bench.py --workload render --render-hw 128 256 --steps 1 --warmup 1"
timeout 300 $RCMD > gpurun_out/plain_render.log 2>&1 || { echo "plain render failed"; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/r02_render_launches.csv $RCMD > gpurun_out/ncu_render.log 2>&1
echo "render list rc $?"
python tools/summarize_launches.py gpurun_out/r02_render_launches.csv > gpurun_out/r02_render_launches.txt 2>&1; head -12 gpurun_out/r02_render_launches.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mlp_fused_kernel|wgrad_batch_kernel" --launch-skip 30 --launch-count 10 -f -o gpurun_out/r02_prof_step $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc $?"; ls -la gpurun_out/r02_prof_step.ncu-rep
