mkdir -p gpurun_out
timeout 300 python tools/bench_micro.py 22 > gpurun_out/micro_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:resample_kernel --launch-skip 3 --launch-count 1 -f -o gpurun_out/r02g_resample python tools/bench_micro.py 22 > gpurun_out/ncu_one.log 2>&1
echo "rc $?"
