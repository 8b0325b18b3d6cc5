mkdir -p gpurun_out
python bench.py --workload render --steps 1 --warmup 1 > gpurun_out/render1.log 2>&1; echo "render rc $?"; tail -1 gpurun_out/render1.log | cut -c1-700
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/train2.log 2>&1; echo "train2 rc $?"; tail -1 gpurun_out/train2.log | cut -c1-900
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload render --steps 1 --warmup 1 > gpurun_out/render2.log 2>&1; echo "render2 rc $?"; tail -1 gpurun_out/render2.log | cut -c1-500
