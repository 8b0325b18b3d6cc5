mkdir -p gpurun_out
timeout 120 python tools/bench_fused.py 303104 $FUSED_FLAGS > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_fused_kernel -s 3 -c 1 -f -o gpurun_out/prof_fused python tools/bench_fused.py 303104 $FUSED_FLAGS > gpurun_out/ncu_fused.log 2>&1; echo "ncu rc $?"; tail -3 gpurun_out/ncu_fused.log
