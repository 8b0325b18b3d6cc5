"""BASELINE.json full sizes on the GPU, checked through size-independent properties (the CPU oracle cannot reach these
sizes in seconds): configs[1] - the 8192-ray PanoNeRF training step, bf16 fused path against the fp32 parity path
(which tests/test_models_gpu.py pins to the reference's golden vectors) - and configs[2] - the 1024 x 512 panorama
render (524 288 rays) - plus row-sharding invariance of the render (what the multi-GPU run relies on)."""
import math

import numpy as np
import pytest
import torch

from util import O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _system(precision):
    from panonerf_b200 import _lib
    if not _lib.lib().pnb_tc_available():
        pytest.skip("not an sm_100 device")
    from panonerf_b200.systems.base_system import default_hparams
    from panonerf_b200.systems.panonerf_system import PanoNeRFSystem
    from panonerf_b200.datasets.pano_datasets import generate_lit_rays, pixel_radius
    hp = default_hparams("panonerf", precision=precision)
    hp["train.randomized"] = False
    system = PanoNeRFSystem(hp).to(DEV)
    system.mip_nerf.mlp.load_state_dict(O.synth_state_dict(seed=4, width=256, c_density=5))
    c2w = np.eye(4, dtype=np.float32)
    c2w[:3, 3] = [0.1, 0.2, 0.3]
    dev = torch.device(DEV, 0)
    system.env_rays = generate_lit_rays(pixel_radius(256, 512, c2w, dev), num=10, device=dev)
    return system, c2w


def test_config1_training_step_8192_rays_bf16_against_fp32_path():
    from panonerf_b200.datasets.pano_datasets import generate_rays
    out = {}
    for prec in ("fp32", "bf16"):
        system, c2w = _system(prec)
        rays = generate_rays(256, 512, c2w, 0.0, 10.0, DEV)
        perm = torch.randperm(256 * 512, generator=torch.Generator().manual_seed(0))[:8192].to(DEV)
        rays = type(rays)(*[x[perm].contiguous() for x in rays])
        gt = (torch.rand(8192, 3, generator=torch.Generator().manual_seed(0)) * 2).to(DEV)
        loss = system.training_step((rays, gt))
        loss.backward()
        out[prec] = (float(loss), {k: p.grad.detach().double().flatten() for k, p in system.mip_nerf.mlp.named_parameters()})
    (l32, g32), (l16, g16) = out["fp32"], out["bf16"]
    assert math.isfinite(l16) and abs(l16 - l32) <= 3e-2 * abs(l32), (l16, l32)      # stated bf16 bound on the loss
    for k in g32:
        assert torch.isfinite(g16[k]).all(), k
        if k.startswith(("extra_layer", "view_layers", "color_layer")):               # first-order terms only
            cos = float((g16[k] @ g32[k]) / (g16[k].norm() * g32[k].norm() + 1e-30))
            assert cos > 0.98, (k, cos)


def test_config2_panorama_render_properties_and_row_sharding():
    from panonerf_b200.datasets.pano_datasets import generate_rays
    system, c2w = _system("bf16")
    H, W = 512, 1024

    def render(row0, nrows):
        rays = generate_rays(H, W, c2w, 0.0, 10.0, DEV, row0=row0, nrows=nrows)
        rays = type(rays)(*[x.view(1, nrows, W, -1) for x in rays])
        return system.render_image((rays, torch.empty(1, nrows, W, 3, device=DEV)))

    c_rgb, f_rgb, c_dep, f_dep, nor, alb, _, sf, sd = render(0, H)
    assert f_rgb.shape == (1, 3, H, W) and f_dep.shape == (1, 1, H, W)
    for name, x in (("coarse rgb", c_rgb), ("fine rgb", f_rgb), ("surface rgb", sf), ("shading", sd), ("albedo", alb)):
        assert torch.isfinite(x).all() and float(x.min()) >= 0.0, name              # softplus / sigmoid radiance
    for dep in (c_dep, f_dep):
        assert float(dep.min()) >= 0.0 and float(dep.max()) <= 10.0                 # clamp to [t_0, t_N] (mip.py:475-476)
    assert float(alb.min()) >= 0.03 - 1e-6 and float(alb.max()) <= 0.80 + 1e-6        # sigmoid * 0.77 + 0.03
    n = nor.permute(0, 2, 3, 1).reshape(-1, 3).norm(dim=-1)
    assert float(((n - 1).abs() < 1e-3).float().mean()) > 0.999                       # unit normals
    # a rank's row block is bit-identical to the same rows of the full render (ray-sharded multi-GPU render)
    part = render(192, 64)
    for full, blk in ((f_rgb, part[1]), (f_dep, part[3]), (nor, part[4]), (sf, part[7])):
        assert torch.equal(full[:, :, 192:256], blk)
