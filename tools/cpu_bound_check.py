"""Is the training step launch-bound?  Host time to ENQUEUE a step vs device time to execute it."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
torch.manual_seed(4)
system = bench.make_system(dev)
opt = system.configure_optimizers()
packed_h, gt_h = bench.host_batch(0, bench.RAYS_PER_GPU)
rays_d, gt_d = bench.unpack_rays(packed_h.to(dev)), gt_h.to(dev)
def step():
    opt.zero_grad()
    loss = system.training_step((rays_d, gt_d))
    loss.backward()
    opt.step()
for _ in range(5):
    step()
torch.cuda.synchronize()
n = 10
t0 = time.perf_counter()
for _ in range(n):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3*(t1-t0)/n:.2f} ms/step, total {1e3*(t2-t0)/n:.2f} ms/step")
