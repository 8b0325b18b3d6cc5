"""Time the UNMODIFIED reference (imported from /root/reference through oracle/ref_harness.py) on the host cores of the
build container: the verbatim PanoNeRFSystem.training_step (forward, backward, torch.optim.Adam) on synthetic rays.
The upstream tree is not present on the GPU box, so this number is recorded once in BASELINE.md; bench.py's CPU arm
times the oracle port of the same algorithm on the GPU box's own host.

    python tools/time_reference_here.py [rays]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402
from oracle import panonerf_oracle as O  # noqa: E402


def main():
    n_rays = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    threads = os.cpu_count()
    torch.set_num_threads(threads)
    ns = rh.load()
    c2w = np.eye(4, dtype=np.float32)
    c2w[:3, 3] = [0.1, 0.2, 0.3]
    ds = rh.make_pano_dataset(ns, 256, 512, [c2w])
    rays = ns.Rays(*[torch.from_numpy(np.asarray(getattr(ds.rays, k)[0])).float().reshape(-1, np.asarray(getattr(ds.rays, k)[0]).shape[-1])
                     for k in ns.Rays._fields])
    g = torch.Generator().manual_seed(0)
    perm = torch.randperm(256 * 512, generator=g)[:n_rays]
    rays = ns.Rays(*[x[perm].contiguous() for x in rays])
    gt = torch.rand(n_rays, 3, generator=g) * 2
    env = ns.Rays(*[x.float() for x in ds.generate_lit_rays(num=10)])
    torch.manual_seed(4)
    model = ns.pano_mip_nerf.PanoMipNeRF(num_samples=64, rgb_activation="softplus", rgb_padding=0.0,
                                         mlp_num_density_channels=5, num_env_samples=10)
    model.mlp.load_state_dict(O.synth_state_dict(seed=4, width=256, c_density=5))
    sysm = ns.panonerf_system.PanoNeRFSystem.__new__(ns.panonerf_system.PanoNeRFSystem)
    torch.nn.Module.__init__(sysm)
    sysm._hp = rh._AttrDict({"train.surface_start_step": 0, "train.surface": True, "loss.ort_loss": 0.1,
                             "loss.coarse_loss_mult": 0.1, "loss.surface_loss": 1, "loss.chrom_loss": 0.1})
    sysm.mip_nerf, sysm.env_rays, sysm.train_randomized, sysm.white_bkgd = model, env, True, False
    opt = torch.optim.Adam(model.mlp.parameters(), lr=2e-4)

    def step():
        opt.zero_grad()
        loss = sysm.training_step((rays, gt, None, None, None), 0)
        loss.backward()
        opt.step()
        return float(loss)

    step()
    times = []
    for _ in range(2):
        t0 = time.perf_counter()
        loss = step()
        times.append(time.perf_counter() - t0)
    dt = min(times)
    cpu = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][:1]
    print(json.dumps({"impl": "reference (unmodified upstream tree, verbatim training_step + torch.optim.Adam)",
                      "rays": n_rays, "seconds_per_step": dt, "rays_per_s": n_rays / dt, "threads": threads,
                      "cpu": cpu[0] if cpu else None, "torch": torch.__version__, "loss": loss}))


if __name__ == "__main__":
    main()
