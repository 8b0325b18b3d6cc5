"""GPU parity for the variants the upstream hot path does not call today: specular BRDFs
(utils/surface_rendering.py:6-101), the microfacet branch of surface_rendering (:147-151), RotToTarget.rot2t
(utils/vector_rotation.py:50-89), sample_each_points_hemisp (models/mip.py:197-237) and
volumetric_lighting_composing (models/mip.py:486-527).  Forward values against the reference's own outputs
(tests/golden/variants.npz, made by make_golden.py), gradients against autograd through the oracle.  fp32, 1e-5
relative unless noted."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from util import O, T, assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def g():
    return load_golden("variants.npz")


def _leaf(x):
    return x.clone().requires_grad_()


@pytest.mark.parametrize("name", ["mf", "bp"])
def test_specular_brdfs_forward_and_gradients(g, name):
    from panonerf_b200.utils import surface_rendering as sr
    fn = sr.microfeast_brdf if name == "mf" else sr.blinn_phong_brdf
    f = lambda k: T(g[k]).to(DEV)
    alb, nrm, rough = _leaf(f("albedo")), _leaf(f("normal")), _leaf(f("roughness"))
    dif, spec, nol = fn(alb, nrm, rough, f("l"), f("v"))
    assert dif.shape == (48, 10, 3) and spec.shape == (48, 10, 1) and nol.shape == (48, 10, 1)
    assert_close(dif.cpu(), T(g[f"{name}_diffuse_brdf"]), 1e-6, "diffuse_brdf")
    assert_close(nol.cpu(), T(g[f"{name}_nol"]), 1e-5, "NoL", floor=1e-3)
    assert_close(spec.cpu(), T(g[f"{name}_spec"]), 2e-5, "spec", floor=1e-4)
    # gradients: autograd through the oracle (masked division: upstream's own gradient is NaN wherever a light is
    # below the horizon, see oracle.microfacet_terms)
    gen = torch.Generator().manual_seed(3)
    gs, gn, gd = torch.rand(48, 10, 1, generator=gen), torch.rand(48, 10, 1, generator=gen), torch.rand(48, 10, 3, generator=gen)
    ((spec * gs.to(DEV)).sum() + (nol * gn.to(DEV)).sum() + (dif * gd.to(DEV)).sum()).backward()
    ca, cn, cr = _leaf(T(g["albedo"])), _leaf(T(g["normal"])), _leaf(T(g["roughness"]))
    ofn = (lambda *a: O.microfacet_terms(*a, masked=True)) if name == "mf" else O.blinn_phong_terms
    odif, ospec, onol = ofn(ca, cn, cr, T(g["l"]), T(g["v"]))
    ((ospec * gs).sum() + (onol * gn).sum() + (odif * gd).sum()).backward()
    assert torch.isfinite(cn.grad).all() and torch.isfinite(cr.grad).all()
    assert_close(alb.grad.cpu(), ca.grad, 1e-5, "d albedo")
    assert_close(nrm.grad.cpu(), cn.grad, 5e-5, "d normal")
    assert_close(rough.grad.cpu(), cr.grad, 5e-5, "d roughness")


def test_microfacet_surface_rendering(g):
    from panonerf_b200.utils.surface_rendering import surface_rendering
    f = lambda k: T(g[k]).to(DEV)
    env, alb, nrm, rough = _leaf(f("env")), _leaf(f("albedo")), _leaf(f("normal")), _leaf(f("roughness"))
    rgb, dif, spc, shading = surface_rendering(env, alb, nrm, rough, f("l"), f("v"), f("omega"), output_sd=True)
    assert shading is None                                   # as upstream (utils/surface_rendering.py:163-165)
    assert_close(rgb.cpu(), T(g["sr_rgb"]), 1e-5, "rgb")
    assert_close(dif.cpu(), T(g["sr_diffuse"]), 1e-5, "diffuse")
    assert_close(spc.cpu(), T(g["sr_specular"]), 2e-5, "specular")
    gen = torch.Generator().manual_seed(4)
    g1, g2, g3 = (torch.rand(48, 3, generator=gen) for _ in range(3))
    ((rgb * g1.to(DEV)).sum() + (dif * g2.to(DEV)).sum() + (spc * g3.to(DEV)).sum()).backward()
    ce, ca, cn, cr = _leaf(T(g["env"])), _leaf(T(g["albedo"])), _leaf(T(g["normal"])), _leaf(T(g["roughness"]))
    orgb, odif, ospc = O.rough_shade(ce, ca, cn, cr, T(g["l"]), T(g["v"]), T(g["omega"]), masked=True)
    ((orgb * g1).sum() + (odif * g2).sum() + (ospc * g3).sum()).backward()
    for name, a, b in (("env", env, ce), ("albedo", alb, ca), ("normal", nrm, cn), ("roughness", rough, cr)):
        assert_close(a.grad.cpu(), b.grad, 5e-5, "d " + name)
    # the live Lambertian call is untouched by the new branch
    rgb_l, dif_l, spec_l, sh_l = surface_rendering(f("env"), f("albedo"), f("normal"), None, f("l")[:1].expand(48, -1, -1),
                                                   f("v"), f("omega"), output_sd=True)
    ref = O.lambert_shade(T(g["env"]), T(g["albedo"]), T(g["normal"]), T(g["l"])[:1].expand(48, -1, -1), T(g["omega"]))
    assert_close(rgb_l.cpu(), ref[0], 1e-5, "lambert rgb")
    assert float(spec_l.abs().max()) == 0.0


def test_rot_to_target(g):
    from panonerf_b200.utils.vector_rotation import RotToTarget
    tv = _leaf(T(g["tvec"]).to(DEV))
    rot = RotToTarget().rot2t(tv)
    assert rot.shape == (48, 3, 3)
    assert torch.allclose(rot.cpu(), T(g["rot"]), rtol=0, atol=2e-6)
    assert torch.equal(rot[1].cpu(), torch.diag(torch.tensor([1.0, -1.0, 1.0])))      # antipodal target
    assert torch.equal(rot[2].cpu(), torch.eye(3))                                   # target == +y
    # it is a rotation onto the target: R (0,1,0) == tvec
    ok = torch.ones(48, dtype=torch.bool)
    ok[1] = False
    assert torch.allclose(rot[:, :, 1].detach().cpu()[ok], T(g["tvec"])[ok], atol=2e-6)
    gr = torch.rand(48, 3, 3, generator=torch.Generator().manual_seed(5))
    keep = torch.ones(48, dtype=torch.bool)
    keep[1:3] = False                                        # acos'(+-1) is infinite there (upstream: inf/nan as well)
    (rot * gr.to(DEV))[keep.to(DEV)].sum().backward()
    ct = _leaf(T(g["tvec"]))
    (O.rot_to_target(ct) * gr)[keep].sum().backward()
    assert_close(tv.grad.cpu()[keep], ct.grad[keep], 1e-4, "d tvec")


def test_hemisphere_env_sampling(g):
    from panonerf_b200.models.mip import sample_each_points_hemisp
    f = lambda k: T(g[k]).to(DEV)
    pts = _leaf(f("points"))
    t, (m, c), dirs = sample_each_points_hemisp(pts.view(-1, 1, 3), f("l"), 8, f("env_near"), f("env_far"),
                                                f("env_radii"), True, t_rand=f("hs_t_rand"))
    assert_close(t.cpu(), T(g["hs_t"]), 1e-6, "t")
    assert_close(m.cpu(), T(g["hs_mean"]), 1e-5, "means")
    assert_close(c.cpu(), T(g["hs_cov"]), 1e-5, "covs")
    assert torch.equal(dirs.cpu(), T(g["hs_dirs"]))
    gm = torch.rand(480, 8, 3, generator=torch.Generator().manual_seed(6))
    (m * gm.to(DEV)).sum().backward()
    assert_close(pts.grad.cpu(), gm.view(48, 80, 3).sum(1), 1e-5, "d points")
    t0, (m0, c0), _ = sample_each_points_hemisp(f("points").view(-1, 1, 3), f("l"), 8, f("env_near"), f("env_far"),
                                                f("env_radii"), False)
    assert_close(t0.cpu(), T(g["hs_t_det"]), 1e-6, "t det")
    assert_close(m0.cpu(), T(g["hs_mean_det"]), 1e-5, "means det")
    assert_close(c0.cpu(), T(g["hs_cov_det"]), 1e-5, "covs det")


def test_attenuated_compositing(g):
    from panonerf_b200.models.mip import volumetric_lighting_composing, volumetric_rendering
    f = lambda k: T(g[k]).to(DEV)
    rgb, den = _leaf(f("vl_rgb")), _leaf(f("vl_density"))
    comp, dist, acc, w, tm = volumetric_lighting_composing(rgb, den, f("hs_t"), f("hs_dirs"), True, output_t=True)
    assert_close(comp.cpu(), T(g["vl_comp"]), 1e-5, "comp")
    assert_close(dist.cpu(), T(g["vl_dist"]), 1e-5, "dist")
    assert_close(acc.cpu(), T(g["vl_acc"]), 1e-5, "acc")
    assert_close(w.cpu(), T(g["vl_weights"]), 1e-5, "weights", floor=2e-2)    # (floors as in test_composite_fwd_bwd)
    assert tm.shape == w.shape
    ((comp * f("vl_g_comp")).sum() + (dist * f("vl_g_dist")).sum() + (acc * f("vl_g_acc")).sum()
     + (w * f("vl_g_w")).sum()).backward()
    assert_close(rgb.grad.cpu(), T(g["vl_d_rgb"]), 1e-5, "d rgb", floor=2e-2)
    assert_close(den.grad.cpu(), T(g["vl_d_density"]), 5e-5, "d density")         # (floor: mean magnitude)
    # N = 8 runs the blocked kernel; a 300-sample ray runs the generic one: both against the oracle
    gen = torch.Generator().manual_seed(7)
    n = 300
    t = torch.sort(torch.rand(64, n + 1, generator=gen) * 6, dim=-1).values
    c_in, d_in = torch.rand(64, n, 3, generator=gen), -torch.log(torch.rand(64, n, 1, generator=gen)) * 0.2
    dirs = torch.randn(64, 3, generator=gen)
    ref = O.composite_lighting(c_in, d_in, t, dirs, False)
    got = volumetric_lighting_composing(c_in.to(DEV), d_in.to(DEV), t.to(DEV), dirs.to(DEV), False)
    for a, b, k in zip(got, ref, ("comp", "dist", "acc", "w")):
        assert_close(a.cpu(), b, 2e-5, k, floor=2e-2 if k == "w" else 1e-3)
    # the un-attenuated path is unchanged by the flag word
    plain = volumetric_rendering(c_in.to(DEV), d_in.to(DEV), t.to(DEV), dirs.to(DEV), True)
    ref_p = O.composite(c_in, d_in, t, dirs, True)
    assert_close(plain[0].cpu(), ref_p[0], 2e-5, "plain comp")
