# Round evidence: (1) per-launch device times of one training step, (2) full ncu capture of the dominant kernel.
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "list rc $?"
ITERS=3 python tools/bench_gemm.py > gpurun_out/gemm_plain.log 2>&1 && \
ITERS=1 ncu --set full --clock-control none --import-source on -k regex:linear_tc_fast -s 4 -c 2 -f -o gpurun_out/prof_linear_fast \
   python tools/bench_gemm.py > gpurun_out/gemm_ncu.log 2>&1
echo "full rc $?"; cat gpurun_out/gemm_plain.log
