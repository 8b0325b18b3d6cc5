"""Build-container only: the oracle against the UNMODIFIED reference imported from /root/reference (skipped where the
upstream tree is absent - there the golden vectors of tests/golden/ pin the oracle instead)."""
import numpy as np
import pytest
import torch

from oracle import ref_harness as rh
from util import O

pytestmark = pytest.mark.skipif(not rh.available(), reason="upstream tree not present")


@pytest.mark.parametrize("pano", [False, True])
def test_forward_matches_reference_modules(pano):
    ns = rh.load()
    c2w = np.eye(4, dtype=np.float32)
    c2w[:3, 3] = [0.1, 0.2, 0.3]
    ds = rh.make_pano_dataset(ns, 8, 16, [c2w])
    rays = ns.Rays(*[torch.from_numpy(np.asarray(getattr(ds.rays, k)[0])).float().reshape(-1, np.asarray(getattr(ds.rays, k)[0]).shape[-1])[:20]
                     for k in ns.Rays._fields])
    env = ns.Rays(*[x.float() for x in ds.generate_lit_rays(num=10)])
    torch.manual_seed(4)
    cls = ns.pano_mip_nerf.PanoMipNeRF if pano else ns.mip_nerf.MipNeRF
    model = cls(num_samples=12, rgb_activation="softplus", rgb_padding=0.0, mlp_net_width=32,
                mlp_num_density_channels=5 if pano else 1)
    sd = {k: v.detach().clone() for k, v in model.mlp.state_dict().items()}
    for randomized in (False, True):
        torch.manual_seed(9)
        if pano:
            ref = model(rays=rays, env_rays=env, randomized=randomized, white_bkgd=True, enable_surf=True, use_ort_loss=True)
            torch.manual_seed(9)
            got, _ = O.panonerf_forward(sd, O.Rays(*rays), O.Rays(*env), dict(num_samples=12), randomized=randomized,
                                        white_bkgd=True)
        else:
            ref = model(rays=rays, randomized=randomized, white_bkgd=True, use_ort_loss=True)
            torch.manual_seed(9)
            got, _ = O.mipnerf_forward(sd, O.Rays(*rays), dict(num_samples=12), randomized=randomized, white_bkgd=True,
                                       use_ort_loss=True)
        for lvl in range(2):
            for a, b in zip(ref[lvl], got[lvl]):
                if a is None:
                    assert b is None
                else:
                    assert torch.allclose(a, b, rtol=1e-5, atol=2e-6)


def test_jacrev_normals_equal_autograd_normals():
    sd = O.synth_state_dict(seed=1, width=32, c_density=5)
    g = torch.Generator().manual_seed(0)
    mean, cov, vd = torch.randn(3, 5, 3, generator=g), torch.rand(3, 5, 3, generator=g) * 1e-3, torch.randn(3, 3, generator=g)
    cfg = {**O.DEFAULT_CFG}
    a = O._density_normals(sd, mean, cov, vd, cfg, False)
    b = O._density_normals(sd, mean, cov, vd, {**cfg, "normals_impl": "jacrev"}, False)
    assert torch.allclose(a, b, rtol=1e-4, atol=1e-6)


def test_metric_and_image_restatements_match_reference():
    """oracle.calc_psnr / calc_ws_psnr / solid_angle_refinement / png_pixels against utils/metrics.py,
    utils/surface_rendering.py and utils/vis.py of the upstream tree."""
    import importlib
    import sys
    ns = rh.load()
    sys.path.insert(0, rh.REF_ROOT)
    try:
        metrics = importlib.import_module("utils.metrics")
    finally:
        sys.path.remove(rh.REF_ROOT)
    g = torch.Generator().manual_seed(0)
    a, b = torch.rand(3, 16, 32, generator=g), torch.rand(3, 16, 32, generator=g)
    assert torch.equal(O.solid_angle_refinement(16, 32), ns.surface_rendering.solid_angle_refinement(h=16, w=32))
    assert float(O.calc_psnr(a, b)) == float(metrics.calc_psnr(a, b))
    assert float(O.calc_ws_psnr(a, b)) == float(metrics.calc_ws_psnr(a, b))
    img = a[None]
    ref = (img[0].permute(1, 2, 0).data.cpu().numpy() * 255).astype(np.uint8)        # utils/vis.py:29-35
    assert np.array_equal(O.png_pixels(img), ref)
