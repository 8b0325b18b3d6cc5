#!/bin/bash
# Round-2 ncu evidence (B200_PROFILING.md recipe): launch lists of one training step and one panorama-chunk render, and
# one --set full capture of the tensor-core kernels of the step (DRAM traffic, tensor-pipe activity).
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --preheat 0 --no-cpu-baseline --no-extras --no-graph"
timeout 300 $CMD > gpurun_out/r02_plain_step.log 2>&1 || { echo "plain step run failed"; tail -5 gpurun_out/r02_plain_step.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/r02_train_step_launches.csv $CMD > gpurun_out/r02_ncu_step.log 2>&1
echo "step list rc $?"
python tools/summarize_launches.py gpurun_out/r02_train_step_launches.csv > gpurun_out/r02_train_step_launches.txt 2>&1; head -14 gpurun_out/r02_train_step_launches.txt
RCMD="python bench.py --workload render --render-hw 128 256 --steps 1 --warmup 1"
timeout 300 $RCMD > gpurun_out/r02_plain_render.log 2>&1 || { echo "plain render failed"; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/r02_render_launches.csv $RCMD > gpurun_out/r02_ncu_render.log 2>&1
echo "render list rc $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mlp_fused_kernel|wgrad_batch_kernel" --launch-skip 30 --launch-count 10 -f -o gpurun_out/r02_prof_step $CMD > gpurun_out/r02_ncu_full.log 2>&1
echo "full rc $?"; ls -la gpurun_out/r02_prof_step.ncu-rep
