mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"composite_fwd_kernel|sample_cast_kernel|resample_kernel" --launch-skip 9 --launch-count 3 -f -o gpurun_out/prof_micro python tools/bench_micro.py 22 > gpurun_out/ncu_micro.log 2>&1
echo "ncu rc $?"; tail -2 gpurun_out/ncu_micro.log | cut -c1-200
