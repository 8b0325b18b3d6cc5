mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_wgrad_tests.log 2>&1
echo "pytest exit $?"; grep -E "^E  |FAILED|passed|failed" gpurun_out/r2_wgrad_tests.log | head -8
echo "== micro (time split)"
timeout 120 python tools/bench_wgrad.py 2>&1 | tail -2 | cut -c1-120
timeout 120 python tools/bench_wgrad.py 819200 2>&1 | tail -2 | cut -c1-120
for rep in 1 2; do
  for lib in base new; do
    if [ $lib = base ]; then export PNB_LIB_PATH=$PWD/tools/bin/libpnb_base.so; else unset PNB_LIB_PATH; fi
    timeout 300 python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_wgrad_ab_${lib}_$rep.json 2> gpurun_out/r2_wgrad_ab_${lib}_$rep.err
    echo "$lib $rep rc $?"
    python - <<PY
import json
d=json.loads(open("gpurun_out/r2_wgrad_ab_${lib}_$rep.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("${lib}", d["ms_per_step"], "fused", r["kernel_ms_per_step"], "wgrad", r["other_kernels"]["wgrad_batch_kernel"]["kernel_ms_per_step"], d["clocks"], d["final_loss"])
PY
  done
done
