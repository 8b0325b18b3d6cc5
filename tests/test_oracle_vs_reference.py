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


def test_variant_restatements_match_reference():
    """SURVEY 8f rank 4 functions (dead upstream, restated in the oracle) against the upstream tree on fresh inputs."""
    import importlib
    import sys
    ns = rh.load()
    sys.path.insert(0, rh.REF_ROOT)
    try:
        vr = importlib.import_module("utils.vector_rotation")
    finally:
        sys.path.remove(rh.REF_ROOT)
    sr, mip = ns.surface_rendering, ns.mip
    g = torch.Generator().manual_seed(0)
    b, d = 37, 10
    nz = lambda *s: torch.nn.functional.normalize(torch.randn(*s, generator=g), dim=-1)
    alb, nrm, rough, l, v = torch.rand(b, 3, generator=g), nz(b, 3), torch.rand(b, 1, generator=g) + 0.05, nz(b, d, 3), nz(b, 3)
    env, om = torch.rand(b, d, 3, generator=g) * 3, torch.full((1, d, 1), 4 * np.pi / d)
    for fr, fo in ((sr.microfeast_brdf, O.microfacet_terms), (sr.blinn_phong_brdf, O.blinn_phong_terms)):
        for x, y in zip(fr(alb, nrm, rough, l, v), fo(alb, nrm, rough, l, v)):
            assert torch.allclose(x, y, rtol=1e-6, atol=1e-7)
    assert torch.equal(sr.microfeast_brdf(alb, nrm, rough, l, v)[1], O.microfacet_terms(alb, nrm, rough, l, v, masked=True)[1])
    for x, y in zip(sr.surface_rendering(env, alb, nrm, rough, l, v, om), O.rough_shade(env, alb, nrm, rough, l, v, om)):
        assert torch.equal(x, y)
    tv = nz(20, 3)
    tv[3], tv[4] = torch.tensor([0.0, -1.0, 0.0]), torch.tensor([0.0, 1.0, 0.0])
    assert torch.equal(vr.RotToTarget().rot2t(tv.clone()), O.rot_to_target(tv))
    pts = torch.randn(b, 3, generator=g)
    e32 = O.Rays(*[x.float() for x in O.fibonacci_env_rays(d, 0.0035)])
    torch.manual_seed(5)
    tr, (mr, cr), dr = mip.sample_each_points_hemisp(pts.view(-1, 1, 3), l, 7, e32.near, e32.far, e32.radii, True)
    torch.manual_seed(5)
    to, (mo, co), do = O.env_samples_hemisp(pts, l, e32, 7, True)
    assert torch.equal(tr, to) and torch.equal(mr, mo) and torch.equal(cr, co) and torch.equal(dr, do)
    rgb, den = torch.rand(b * d, 7, 3, generator=g), -torch.log(torch.rand(b * d, 7, 1, generator=g))
    for x, y in zip(mip.volumetric_lighting_composing(rgb, den, tr, dr, True), O.composite_lighting(rgb, den, to, do, True)):
        assert torch.equal(x, y)
