"""Reproduce the accumulated (C4) training step at a small size with a full traceback."""
import os, sys, traceback
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from panonerf_b200.systems.base_system import AccumulatedTrainStep, GraphedTrainStep

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
system = bench.make_system(dev)
opt = system.configure_optimizers()
n, k = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 4
packed_h, gt_h = bench.host_batch(0, n)
rays = bench.unpack_rays(packed_h.to(dev))
gt = gt_h.to(dev)
if "--with-graphed" in sys.argv:
    g = GraphedTrainStep(system, opt, bench.unpack_rays(packed_h[:1024].to(dev)), gt[:1024].clone())
    g()
try:
    step = AccumulatedTrainStep(system, opt, rays, gt, micro_batches=k)
    for _ in range(3):
        print("loss", float(step(rays, gt)))
    # equivalence with one eager step on the whole batch (deterministic sampling)
    print("ok")
except Exception:
    traceback.print_exc()
