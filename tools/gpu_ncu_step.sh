mkdir -p gpurun_out
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mlp_fused_kernel|wgrad_batch_kernel" --launch-skip 30 --launch-count 10 -f -o gpurun_out/prof_step python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_step.log 2>&1
echo "ncu rc $?"; tail -2 gpurun_out/ncu_step.log | cut -c1-200
