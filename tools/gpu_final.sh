# End-of-round verification + evidence refresh (what the driver runs, plus the micro-benchmarks).
mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/tests.log 2>&1
echo "pytest exit $?"; grep -E "^E  |FAILED|passed|failed" gpurun_out/tests.log | head -8
timeout 120 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log | cut -c1-200
timeout 300 python bench.py > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "train rc $?"; tail -1 gpurun_out/bench_train.json | cut -c1-250
timeout 300 python bench.py --workload render --steps 2 --warmup 1 > gpurun_out/bench_render.json 2> gpurun_out/bench_render.err; echo "render rc $?"; tail -1 gpurun_out/bench_render.json | cut -c1-250
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc $?"; tail -1 gpurun_out/bench_reference.json | cut -c1-200
timeout 120 python tools/bench_micro.py > gpurun_out/micro.log 2>&1; echo "micro rc $?"; cut -c1-120 gpurun_out/micro.log
for fl in "" "--normals" "--save" "--normals --save" "--bwd" "--jadj"; do timeout 120 python tools/bench_fused.py $fl 2>&1 | tail -1; done > gpurun_out/fused_micro.log
cut -c1-130 gpurun_out/fused_micro.log
