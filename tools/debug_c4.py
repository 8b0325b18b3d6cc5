"""Reproduce the accumulated (C4) training step after the same history bench.py has, with a full traceback."""
import os, sys, traceback
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from panonerf_b200 import field
from panonerf_b200.systems.base_system import AccumulatedTrainStep, GraphedTrainStep

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
system = bench.make_system(dev)
opt = system.configure_optimizers()
n, k = 4096, 4
packed_h, gt_h = bench.host_batch(0, n)
rays = bench.unpack_rays(packed_h.to(dev))
gt = gt_h.to(dev)
flags = set(sys.argv[1:])
small = bench.unpack_rays(packed_h[:1024].to(dev))
if "graphed" in flags:
    g = GraphedTrainStep(system, opt, small, gt[:1024].clone())
    for _ in range(3):
        g()
if "eager" in flags:
    for _ in range(2):
        opt.zero_grad()
        loss = system.training_step((small, gt[:1024].contiguous()))
        loss.backward()
        opt.step()
if "profile" in flags:
    field.PROFILE = {}
    opt.zero_grad()
    loss = system.training_step((small, gt[:1024].contiguous()))
    loss.backward()
    opt.step()
    torch.cuda.synchronize()
    prof = field.PROFILE
    field.PROFILE = None
    print("profiled families", list(prof))
if "sampler" in flags:
    s = bench.ClockSampler(0)
    s.start()
    torch.cuda.synchronize()
    print(s.result())
try:
    step = AccumulatedTrainStep(system, opt, rays, gt, micro_batches=k)
    for _ in range(2):
        print("loss", float(step(rays, gt)))
    print("ok", sorted(flags))
except Exception:
    traceback.print_exc()
    print("FAILED", sorted(flags))
