mkdir -p gpurun_out
ITERS=3 python tools/bench_gemm.py > gpurun_out/gemm_plain.log 2>&1 && \
ITERS=1 ncu --set full --clock-control none --import-source on -k regex:linear_tc_kernel -s 4 -c 2 -f -o gpurun_out/prof_linear \
   python tools/bench_gemm.py > gpurun_out/gemm_ncu.log 2>&1
echo rc $?; cat gpurun_out/gemm_plain.log
