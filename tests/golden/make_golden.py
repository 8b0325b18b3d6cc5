"""Generate the golden vectors in this directory by EXECUTING THE UNMODIFIED REFERENCE (imported from
/root/reference, which exists only in the build container).  Re-run with

    python tests/golden/make_golden.py

The outputs (`*.npz`) are committed; the GPU box and CI only read them.  Inputs are synthetic:
an equirect grid from the reference's own `PanoDataset._generate_rays`, near=0, far=10, weights from the
reference's own initialiser under `torch.manual_seed(4)` (small nets, stored in the fixture) or from
`oracle.panonerf_oracle.synth_state_dict` (full-width net, regenerated from its seed and checksummed).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh          # noqa: E402
from oracle import panonerf_oracle as O       # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def npy(x):
    return None if x is None else x.detach().cpu().numpy()


def camera(rot_seed=None, trans=(0.1, 0.2, 0.3)):
    c2w = np.eye(4, dtype=np.float32)
    if rot_seed is not None:
        from scipy.spatial.transform import Rotation
        c2w[:3, :3] = Rotation.random(random_state=rot_seed).as_matrix().astype(np.float32)
    c2w[:3, 3] = trans
    return c2w


def ref_rays(ns, h, w, c2w):
    ds = rh.make_pano_dataset(ns, h, w, [c2w])
    rays = ns.Rays(*[torch.from_numpy(np.asarray(getattr(ds.rays, k)[0])).float()
                     .reshape(-1, np.asarray(getattr(ds.rays, k)[0]).shape[-1]) for k in ns.Rays._fields])
    return ds, rays


def ops_fixture(ns):
    out = {}
    mip, sr = ns.mip, ns.surface_rendering
    # --- ray generation (rotation exercised) + env rays
    h, w = 8, 16
    c2w = camera(rot_seed=0)
    ds, rays = ref_rays(ns, h, w, c2w)
    out["raygen_c2w"] = c2w
    out["raygen_hw"] = np.array([h, w])
    for k in ns.Rays._fields:
        out[f"raygen_{k}"] = npy(getattr(rays, k))
    env = ds.generate_lit_rays(num=10)
    out["env_radius"] = np.array(float(ds.radii))
    for k in ns.Rays._fields:
        out[f"env_{k}"] = npy(getattr(env, k).float())
    # --- sample_along_rays deterministic + randomized (torch.rand stream reproduced by seed)
    b, n = 24, 16
    r = ns.Rays(*[x[:b] for x in rays])
    t, (mean, cov) = mip.sample_along_rays(r.origins, r.directions, r.radii, n, r.near, r.far, False, False, "cone")
    out.update(sample_t=npy(t), sample_mean=npy(mean), sample_cov=npy(cov))
    torch.manual_seed(11)
    t_rand = torch.rand(b, n + 1)
    torch.manual_seed(11)
    t2, (mean2, cov2) = mip.sample_along_rays(r.origins, r.directions, r.radii, n, r.near, r.far, True, False, "cone")
    out.update(sample_t_rand=npy(t_rand), sample_t_r=npy(t2), sample_mean_r=npy(mean2), sample_cov_r=npy(cov2))
    # --- IPE / pos_enc
    g = torch.Generator().manual_seed(1)
    m_in = (torch.rand(5, 7, 3, generator=g) * 10 - 5)
    c_in = torch.rand(5, 7, 3, generator=g) * torch.tensor([1e-6, 1e-3, 1.0])
    out.update(ipe_mean=npy(m_in), ipe_cov=npy(c_in), ipe_out=npy(mip.integrated_pos_enc((m_in, c_in), 0, 16)))
    out.update(posenc_out=npy(mip.pos_enc(r.viewdirs, 0, 4, True)))
    # --- volumetric_rendering + autograd grads
    rgb = torch.rand(b, n, 3, generator=g, requires_grad=True)
    den = (-torch.log(torch.rand(b, n, 1, generator=g))).requires_grad_(True)
    den.data[3] = 0.0                                             # an empty ray: acc == 0 -> nan_to_num path
    den.data[4] *= 50.0                                           # an opaque ray
    comp, dist, acc, wts = mip.volumetric_rendering(rgb, den, t2, r.directions, white_bkgd=True)
    gc, gd, ga, gw = (torch.rand(b, 3, generator=g), torch.rand(b, generator=g), torch.rand(b, generator=g),
                      torch.rand(b, n, generator=g))
    (comp * gc).sum().add((dist * gd).sum()).add((acc * ga).sum()).add((wts * gw).sum()).backward()
    out.update(vr_rgb=npy(rgb), vr_density=npy(den), vr_t=npy(t2), vr_dirs=npy(r.directions), vr_comp=npy(comp),
               vr_dist=npy(dist), vr_acc=npy(acc), vr_weights=npy(wts), vr_g_comp=npy(gc), vr_g_dist=npy(gd),
               vr_g_acc=npy(ga), vr_g_w=npy(gw), vr_d_rgb=npy(rgb.grad), vr_d_density=npy(den.grad))
    # --- sorted_piecewise_constant_pdf / resample (incl. all-zero weights and a spike)
    wt = wts.detach().clone()
    wt[5] = 0.0
    wt[6] = 0.0
    wt[6, 9] = 1.0
    new_t, (rm, rc) = mip.resample_along_rays(r.origins, r.directions, r.radii, t2, wt.clone(), False, "cone", True, 0.01)
    out.update(rs_w=npy(wt), rs_t=npy(t2), rs_new_t=npy(new_t), rs_mean=npy(rm), rs_cov=npy(rc))
    wb = O.blur_weights(wt, 0.01)
    # indices: recompute exactly as mip.py:253-283 does (the function itself does not return them)
    _, inds, cdf = O.pdf_sample(t2, wb.clone(), n + 1, False, return_aux=True)
    assert torch.equal(O.pdf_sample(t2, wb.clone(), n + 1, False), new_t)
    assert torch.equal(mip.sorted_piecewise_constant_pdf(t2, wb.clone(), n + 1, False), new_t)
    out.update(rs_inds=npy(inds), rs_cdf=npy(cdf))
    torch.manual_seed(12)
    new_t_r = mip.sorted_piecewise_constant_pdf(t2, wb.clone(), n + 1, True)
    torch.manual_seed(12)
    s = 1 / (n + 1)
    u_r = (torch.arange(n + 1) * s)[None] + torch.empty(b, n + 1).uniform_(to=(s - O.F32_EPS))
    u_r = torch.clamp_max(u_r, 1.0 - O.F32_EPS)
    out.update(rs_u_r=npy(u_r), rs_new_t_r=npy(new_t_r))
    # --- surface_rendering + hdr_to_ldr
    e = torch.rand(b, 10, 3, generator=g) * 3
    alb = torch.rand(b, 3, generator=g)
    nrm = torch.nn.functional.normalize(torch.randn(b, 3, generator=g), dim=-1)
    l = env.directions.float()[None].expand(b, -1, -1).contiguous()
    srgb, dif, _, shd = sr.surface_rendering(e, alb, nrm, None, l, r.viewdirs, env.lossmult.float(), output_sd=True)
    out.update(sr_env=npy(e), sr_albedo=npy(alb), sr_normal=npy(nrm), sr_l=npy(l), sr_rgb=npy(srgb),
               sr_diffuse=npy(dif), sr_shading=npy(shd))
    x = torch.rand(b, 3, generator=g) * 4
    out.update(tm_in=npy(x), tm_out=npy(sr.hdr_to_ldr(x)), tm_out_u8=npy(sr.hdr_to_ldr(x, dtype="uint8")))
    np.savez_compressed(os.path.join(HERE, "ops.npz"), **out)
    print("ops.npz", len(out), "arrays")


def model_fixture(ns, name, pano, width, b, n, h, w, sd_from_seed=None, store_grads=True, gslice=64,
                  stop_resample_grad=True):
    out = {}
    c2w = camera(rot_seed=1)
    ds, rays = ref_rays(ns, h, w, c2w)
    perm = torch.randperm(h * w, generator=torch.Generator().manual_seed(0))[:b]
    r = ns.Rays(*[x[perm].contiguous() for x in rays])
    env = ds.generate_lit_rays(num=10)
    env32 = ns.Rays(*[x.float() for x in env])
    gt = torch.rand(b, 3, generator=torch.Generator().manual_seed(0)) * 2
    torch.manual_seed(4)
    cls = ns.pano_mip_nerf.PanoMipNeRF if pano else ns.mip_nerf.MipNeRF
    model = cls(num_samples=n, rgb_activation="softplus", rgb_padding=0.0, mlp_net_width=width,
                mlp_num_density_channels=5 if pano else 1, num_env_samples=10,
                stop_resample_grad=stop_resample_grad)
    if sd_from_seed is not None:
        sd = O.synth_state_dict(seed=sd_from_seed, width=width, c_density=5 if pano else 1)
        model.mlp.load_state_dict(sd)
        out["sd_seed"] = np.array(sd_from_seed)
        out["sd_checksum"] = np.array([float(sum(v.double().sum() for v in sd.values())),
                                       float(sum((v.double() ** 2).sum() for v in sd.values()))])
    else:
        for k, v in model.mlp.state_dict().items():
            out["sd/" + k] = npy(v)
    out.update(c2w=c2w, hw=np.array([h, w]), perm=npy(perm), gt=npy(gt), n=np.array(n), width=np.array(width),
               env_radius=np.array(float(ds.radii)), stop_resample_grad=np.array(int(stop_resample_grad)))
    if pano:
        res = model(rays=r, env_rays=env32, randomized=False, white_bkgd=False, enable_surf=True, use_ort_loss=True)
        loss = O.panonerf_loss(res, r, gt)
        # cross-check the loss restatement against the reference's verbatim training_step
        sysm = ns.panonerf_system.PanoNeRFSystem.__new__(ns.panonerf_system.PanoNeRFSystem)
        torch.nn.Module.__init__(sysm)
        sysm._hp = rh._AttrDict({"train.surface_start_step": 0, "train.surface": True, "loss.ort_loss": 0.1,
                                 "loss.coarse_loss_mult": 0.1, "loss.surface_loss": 1, "loss.chrom_loss": 0.1})
        sysm.mip_nerf, sysm.env_rays, sysm.train_randomized, sysm.white_bkgd = model, env32, False, False
        ref_loss = sysm.training_step((r, gt, None, None, None), 0)
        assert abs(float(ref_loss) - float(loss)) < 1e-6, (float(ref_loss), float(loss))
        names = ["comp_rgb", "distance", "ort_loss", "normal", "albedo", "roughness", "surface_rgb", "diffuse",
                 "shading"]
    else:
        res = model(rays=r, randomized=False, white_bkgd=False, use_ort_loss=True)
        loss = O.mipnerf_loss(res, r, gt, ort_mult=0.1)
        names = ["comp_rgb", "distance", "ort_loss", "normal"]
    for lvl in range(2):
        for nm, v in zip(names, res[lvl]):
            if v is not None:
                out[f"out/{lvl}/{nm}"] = npy(v)
    out["loss"] = npy(loss)
    loss.backward()
    for k, p in model.mlp.named_parameters():
        g = p.grad
        out["gnorm/" + k] = np.array(float(g.norm()))
        if store_grads:
            out["grad/" + k] = npy(g)
        else:
            out["gslice/" + k] = npy(g.reshape(-1)[:: max(1, g.numel() // gslice)][:gslice])
    if not store_grads:
        # float64 evaluation of the oracle on the same inputs: the yardstick for the gradient bounds.  Sums over ~1e6
        # samples with heavy cancellation (high IPE frequencies, 1/|grad sigma| in the normals) make every fp32
        # gradient - the reference's own included - deviate from the exact value; the tests require ours to be as
        # close to this float64 value as the reference's fp32 gradient is.
        for tag, ort_on in (("g64", True),) + ((("noort/g64", False),) if not pano else ()):
            sd64 = {k: v.detach().double().clone().requires_grad_() for k, v in model.mlp.state_dict().items()}
            r64 = O.Rays(*[x.double() for x in r])
            if pano:
                e64 = O.Rays(*[x.double() for x in env32])
                res64, _ = O.panonerf_forward(sd64, r64, e64, dict(num_samples=n), train=True)
                l64 = O.panonerf_loss(res64, r64, gt.double())
            else:
                res64, _ = O.mipnerf_forward(sd64, r64, dict(num_samples=n), use_ort_loss=ort_on, train=True)
                l64 = O.mipnerf_loss(res64, r64, gt.double(), ort_mult=0.1 if ort_on else 0.0)
            l64.backward()
            out[tag + "/loss"] = np.array(float(l64))
            for k, v in sd64.items():
                g = v.grad
                out[tag + "/gnorm/" + k] = np.array(float(g.norm()))
                out[tag + "/gslice/" + k] = npy(g.reshape(-1)[:: max(1, g.numel() // gslice)][:gslice])
            print(name, tag, "float64 loss", float(l64))
    if not pano and not store_grads:
        # first-order variant (configs/mipnerf.yaml: ort_loss 0 -> no normals, no second-order terms): the gradient is a
        # plain sum over samples, so the fp32 parity path can be held to a 1e-5-class bound
        model.mlp.zero_grad()
        res = model(rays=r, randomized=False, white_bkgd=False, use_ort_loss=False)
        loss0 = O.mipnerf_loss(res, r, gt, ort_mult=0.0)
        out["noort/loss"] = npy(loss0)
        loss0.backward()
        for k, p in model.mlp.named_parameters():
            g = p.grad
            out["noort/gnorm/" + k] = np.array(float(g.norm()))
            out["noort/gslice/" + k] = npy(g.reshape(-1)[:: max(1, g.numel() // gslice)][:gslice])
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, "loss", float(loss), "arrays", len(out))


def resample_fixture(ns):
    """Config-size resampling vectors (4096 rays x N = 64 / 128 / 256) from the reference's resample_along_rays:
    searchsorted indices in full (bit-exact contract), the new fence-posts as a SHA-256 of their bytes plus every
    16th row.  Inputs are regenerated from the seed by the tests (torch's CPU generator is host-independent)."""
    import hashlib
    mip = ns.mip
    out = {}
    for n in (64, 128, 256):
        t, w, o, d, rad = O.resample_case(n)
        new_t, (rm, rc) = mip.resample_along_rays(o, d, rad, t, w.clone(), False, "cone", True, 0.01)
        wb = O.blur_weights(w, 0.01)
        new_t2, inds, cdf = O.pdf_sample(t, wb.clone(), n + 1, False, return_aux=True)
        assert torch.equal(new_t2, new_t), "oracle restatement differs from the reference"
        assert torch.equal(mip.sorted_piecewise_constant_pdf(t, wb.clone(), n + 1, False), new_t)
        assert int(inds.max()) <= n and int(inds.min()) >= 1
        out[f"inds/{n}"] = inds.numpy().astype(np.uint16 if n > 255 else np.uint8)
        out[f"new_t_sha256/{n}"] = np.frombuffer(hashlib.sha256(new_t.numpy().tobytes()).digest(), dtype=np.uint8)
        out[f"new_t_rows16/{n}"] = npy(new_t[::16])
        out[f"mean_rows64/{n}"] = npy(rm[::64])
        out[f"cov_rows64/{n}"] = npy(rc[::64])
        out[f"cdf_sha256/{n}"] = np.frombuffer(hashlib.sha256(cdf.numpy().tobytes()).digest(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "resample_large.npz"), **out)
    print("resample_large.npz", len(out), "arrays")


def variants_fixture(ns):
    """Functions the upstream hot path does not call today: specular BRDFs, the microfacet
    branch of surface_rendering, RotToTarget.rot2t, sample_each_points_hemisp, volumetric_lighting_composing."""
    import importlib
    sys.path.insert(0, rh.REF_ROOT)
    try:
        vr = importlib.import_module("utils.vector_rotation")
    finally:
        sys.path.remove(rh.REF_ROOT)
    out = {}
    mip, sr = ns.mip, ns.surface_rendering
    g = torch.Generator().manual_seed(21)
    b, d, ne = 48, 10, 8
    nrm = torch.nn.functional.normalize(torch.randn(b, 3, generator=g), dim=-1)
    alb = torch.rand(b, 3, generator=g)
    rough = torch.rand(b, 1, generator=g) + 0.05
    v = torch.nn.functional.normalize(torch.randn(b, 3, generator=g), dim=-1)
    env = torch.rand(b, d, 3, generator=g) * 3
    tv = nrm.clone()
    tv[1] = torch.tensor([0.0, -1.0, 0.0])
    tv[2] = torch.tensor([0.0, 1.0, 0.0])
    rot = vr.RotToTarget().rot2t(tv.clone())
    envr = O.fibonacci_env_rays(d, 0.0035)
    e32 = ns.Rays(*[x.float() for x in envr])
    hemi = e32.directions.clone()
    hemi[:, 1] = hemi[:, 1].abs()                                  # upper hemisphere around +y
    l = torch.einsum("bij,dj->bdi", rot, hemi).contiguous()        # per-ray light directions
    om = e32.lossmult
    out.update(normal=npy(nrm), albedo=npy(alb), roughness=npy(rough), v=npy(v), env=npy(env), tvec=npy(tv),
               rot=npy(rot), l=npy(l), omega=npy(om))
    for name, fn in (("mf", sr.microfeast_brdf), ("bp", sr.blinn_phong_brdf)):
        dif, spec, nol = fn(alb, nrm, rough, l, v)
        out.update({f"{name}_diffuse_brdf": npy(dif), f"{name}_spec": npy(spec), f"{name}_nol": npy(nol)})
    rgb, dif, spc = sr.surface_rendering(env, alb, nrm, rough, l, v, om)
    out.update(sr_rgb=npy(rgb), sr_diffuse=npy(dif), sr_specular=npy(spc))
    pts = torch.randn(b, 3, generator=g)
    torch.manual_seed(13)
    t_rand = torch.rand(1, ne + 1)
    torch.manual_seed(13)
    t, (mean, cov), dirs = mip.sample_each_points_hemisp(pts.view(-1, 1, 3), l, ne, e32.near, e32.far, e32.radii, True)
    t0, (mean0, cov0), _ = mip.sample_each_points_hemisp(pts.view(-1, 1, 3), l, ne, e32.near, e32.far, e32.radii, False)
    out.update(points=npy(pts), env_near=npy(e32.near), env_far=npy(e32.far), env_radii=npy(e32.radii),
               hs_t_rand=npy(t_rand), hs_t=npy(t), hs_mean=npy(mean), hs_cov=npy(cov), hs_dirs=npy(dirs),
               hs_t_det=npy(t0.contiguous()), hs_mean_det=npy(mean0), hs_cov_det=npy(cov0))
    crgb = torch.rand(b * d, ne, 3, generator=g, requires_grad=True)
    cden = (-torch.log(torch.rand(b * d, ne, 1, generator=g))).requires_grad_(True)
    comp, dist, acc, w = mip.volumetric_lighting_composing(crgb, cden, t, dirs, True)
    gc, gd, ga, gw = (torch.rand(b * d, 3, generator=g), torch.rand(b * d, generator=g), torch.rand(b * d, generator=g),
                      torch.rand(b * d, ne, generator=g))
    (comp * gc).sum().add((dist * gd).sum()).add((acc * ga).sum()).add((w * gw).sum()).backward()
    out.update(vl_rgb=npy(crgb), vl_density=npy(cden), vl_comp=npy(comp), vl_dist=npy(dist), vl_acc=npy(acc),
               vl_weights=npy(w), vl_g_comp=npy(gc), vl_g_dist=npy(gd), vl_g_acc=npy(ga), vl_g_w=npy(gw),
               vl_d_rgb=npy(crgb.grad), vl_d_density=npy(cden.grad))
    np.savez_compressed(os.path.join(HERE, "variants.npz"), **out)
    print("variants.npz", len(out), "arrays")


def main():
    assert rh.available(), "reference tree not found"
    torch.set_num_threads(8)
    ns = rh.load()
    which = set(sys.argv[1:]) or {"ops", "small", "resample", "c1", "c2s", "variants", "rg"}
    if "rg" in which:
        # stop_resample_grad=False (models/mip.py:336-350): the fine level's loss reaches the coarse weights
        model_fixture(ns, "mipnerf_w64_rg.npz", False, 64, 24, 16, 8, 16, stop_resample_grad=False)
        model_fixture(ns, "panonerf_w64_rg.npz", True, 64, 24, 16, 8, 16, stop_resample_grad=False)
    if "variants" in which:
        variants_fixture(ns)
    if "ops" in which:
        ops_fixture(ns)
    if "small" in which:
        model_fixture(ns, "mipnerf_w64.npz", False, 64, 24, 16, 8, 16)
        model_fixture(ns, "panonerf_w64.npz", True, 64, 24, 16, 8, 16)
        model_fixture(ns, "mipnerf_w256.npz", False, 256, 16, 64, 16, 32, sd_from_seed=4, store_grads=False)
        model_fixture(ns, "panonerf_w256.npz", True, 256, 16, 64, 16, 32, sd_from_seed=4, store_grads=False)
    if "resample" in which:
        resample_fixture(ns)
    # BASELINE.json config sizes: C1 = MipNeRF, 4096 rays of a 64 x 128 grid, 128 + 128 samples;
    # C2 subset = PanoMipNeRF, 2048 of the 8192 rays of a 256 x 512 grid, 64 + 64 samples, surface + ort + chroma on.
    if "c1" in which:
        model_fixture(ns, "mipnerf_c1.npz", False, 256, 4096, 128, 64, 128, sd_from_seed=4, store_grads=False,
                      gslice=2048)
    if "c2s" in which:
        model_fixture(ns, "panonerf_c2s.npz", True, 256, 2048, 64, 256, 512, sd_from_seed=4, store_grads=False,
                      gslice=2048)


if __name__ == "__main__":
    main()
