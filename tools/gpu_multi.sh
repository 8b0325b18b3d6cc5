# N-GPU weak-scaling (train) and strong-scaling (render) lines: bash tools/gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/train_n$N.log 2>&1; echo "train$N rc $?"; grep '^{' gpurun_out/train_n$N.log | tail -1 | cut -c1-330
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload render --steps 2 --warmup 1 > gpurun_out/render_n$N.log 2>&1; echo "render$N rc $?"; grep '^{' gpurun_out/render_n$N.log | tail -1 | cut -c1-330
