"""GPU parity: every non-GEMM kernel of the C ABI against the golden vectors of the reference and against the
oracle on seeded inputs.  fp32 tolerance 1e-5 relative (scale-aware) unless noted; indices bit-exact."""
import numpy as np
import pytest
import torch

from util import O, T, assert_close, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rays(g, n=24):
    return O.Rays(*[T(g[f"raygen_{k}"])[:n].to(DEV) for k in O.Rays._fields])


def test_raygen_matches_reference(golden_ops):
    from panonerf_b200.datasets.pano_datasets import generate_rays, generate_lit_rays, pixel_radius
    g = golden_ops
    h, w = [int(v) for v in g["raygen_hw"]]
    rays = generate_rays(h, w, g["raygen_c2w"], 0.0, 10.0, DEV)
    for k in O.Rays._fields:
        assert_close(getattr(rays, k).cpu(), T(g[f"raygen_{k}"]), 2e-6, k)
    # row-block generation (ray sharding) is bit-identical to the full image
    part = generate_rays(h, w, g["raygen_c2w"], 0.0, 10.0, DEV, row0=3, nrows=2)
    for k in O.Rays._fields:
        assert torch.equal(getattr(part, k), getattr(rays, k)[3 * w:5 * w]), k
    rad = pixel_radius(h, w, g["raygen_c2w"], DEV)
    assert abs(rad - float(g["env_radius"])) < 2e-6 * rad
    env = generate_lit_rays(float(g["env_radius"]), num=10, device=DEV)
    for k in O.Rays._fields:
        assert torch.equal(getattr(env, k).float().cpu(), T(g[f"env_{k}"])), k


def test_raygen_large_panorama_properties():
    from panonerf_b200.datasets.pano_datasets import generate_rays
    c2w = np.eye(4, dtype=np.float32)
    c2w[:3, 3] = [0.1, 0.2, 0.3]
    h, w = 512, 1024
    rays = generate_rays(h, w, c2w, 0.0, 10.0, DEV)
    ref = O.equirect_rays(h, w, c2w, 0.0, 10.0)
    assert_close(rays.directions.cpu(), ref.directions, 1e-5, "directions", floor=1.0)
    assert_close(rays.viewdirs.cpu(), ref.viewdirs, 1e-5, "viewdirs", floor=1.0)
    # the per-column radius is a difference of nearly equal fp32 vectors: 1 ulp of a direction is 2e-5 of it
    assert_close(rays.radii.cpu(), ref.radii, 1e-4, "radii")
    assert torch.allclose(rays.viewdirs.norm(dim=-1), torch.ones(h * w, device=DEV), atol=1e-6)


def test_sample_cast(golden_ops):
    from panonerf_b200.models import mip
    g = golden_ops
    r = _rays(g)
    t, (m, c) = mip.sample_along_rays(r.origins, r.directions, r.radii, 16, r.near, r.far, False, False, "cone")
    assert torch.equal(t.cpu(), T(g["sample_t"]))
    assert_close(m.cpu(), T(g["sample_mean"]), 1e-6, "mean")
    assert_close(c.cpu(), T(g["sample_cov"]), 1e-5, "cov")
    t, (m, c) = mip.sample_along_rays(r.origins, r.directions, r.radii, 16, r.near, r.far, True, False, "cone",
                                      t_rand=T(g["sample_t_rand"]).to(DEV))
    assert_close(t.cpu(), T(g["sample_t_r"]), 1e-6, "t_rand")
    assert_close(m.cpu(), T(g["sample_mean_r"]), 1e-6, "mean_r")
    assert_close(c.cpu(), T(g["sample_cov_r"]), 1e-5, "cov_r")
    with pytest.raises(NotImplementedError):
        mip.sample_along_rays(r.origins, r.directions, r.radii, 16, r.near, r.far, False, False, "cylinder")
    with pytest.raises(AssertionError):
        mip.cast_rays(t, r.origins, r.directions, r.radii, "sphere")


def test_randomized_stream_equals_reference(golden_ops):
    """With the same seed the model-side draw consumes torch's CUDA generator exactly like upstream would."""
    from panonerf_b200.models import mip
    r = _rays(golden_ops)
    torch.manual_seed(3)
    t1, _ = mip.sample_along_rays(r.origins, r.directions, r.radii, 16, r.near, r.far, True, False, "cone")
    torch.manual_seed(3)
    t_rand = torch.rand(24, 17, device=DEV)
    cpu = lambda x: x.cpu()
    t2, _ = O.sample_along_rays(cpu(r.origins), cpu(r.directions), cpu(r.radii), 16, cpu(r.near), cpu(r.far), True,
                                t_rand=t_rand.cpu())
    assert_close(t1.cpu(), t2, 1e-6, "t")


def test_disparity_sampling(golden_ops):
    from panonerf_b200.models import mip
    r = _rays(golden_ops)
    near = torch.full_like(r.near, 0.5)
    t, (m, c) = mip.sample_along_rays(r.origins, r.directions, r.radii, 32, near, r.far, False, True, "cone")
    cpu = lambda x: x.cpu()
    t2, (m2, c2) = O.sample_along_rays(cpu(r.origins), cpu(r.directions), cpu(r.radii), 32, cpu(near), cpu(r.far),
                                       False, disparity=True)
    assert_close(t.cpu(), t2, 1e-6, "t")
    assert_close(c.cpu(), c2, 1e-5, "cov")


def test_ipe_and_posenc(golden_ops):
    from panonerf_b200.models import mip
    g = golden_ops
    enc = mip.integrated_pos_enc((T(g["ipe_mean"]).to(DEV), T(g["ipe_cov"]).to(DEV)), 0, 16)
    assert float((enc.cpu() - T(g["ipe_out"])).abs().max()) < 2e-6
    pe = mip.pos_enc(_rays(g).viewdirs, 0, 4, True)
    assert float((pe.cpu() - T(g["posenc_out"])).abs().max()) < 1e-6
    # large seeded case + the vjp / jvp pair against autograd
    gen = torch.Generator().manual_seed(0)
    mean = (torch.rand(4096, 3, generator=gen) * 10 - 5)
    cov = torch.rand(4096, 3, generator=gen) * torch.tensor([1e-6, 1e-4, 1e-2])
    mean_r = mean.clone().requires_grad_()
    ref = O.ipe(mean_r, cov, 0, 16)
    from panonerf_b200 import ops
    out = torch.empty(4096, 96, device=DEV)
    ops.ipe_into(mean.to(DEV), cov.to(DEV), 0, 16, out)
    # The parity pin is the float64 value of the reference's formula on the reference's own fp32 arguments
    # (oracle.ipe_exact) - the number any correct fp32 sin / exp approximates, independent of the host's libm: strict,
    # every element.  Round 1 tolerated "1-2 outliers of 1e-4 in about 1 of 5 suite runs"; round 2 root-caused them
    # (profiles/r02_ipe_outlier_hunt.md): on some hosts the CPU's fp32 torch.sin / exp - i.e. the ORACLE - is off by
    # ~1e-4 on one or two elements while the GPU kernel (both its SFU and its double-precision path) agrees with float64.
    # So: GPU vs float64 strict everywhere; GPU vs CPU oracle strict wherever the CPU oracle itself is within 2e-6 of
    # float64; elements where the HOST misbehaves are written to gpurun_out/ with full diagnostics and reported.
    exact = O.ipe_exact(mean, cov, 0, 16)
    assert float((out.cpu().double() - exact).abs().max()) < 2e-6, "GPU IPE vs float64 evaluation"
    cpu_err = (ref.detach().double() - exact).abs()
    host_bad = cpu_err >= 2e-6
    if bool(host_bad.any()):
        import json
        import os
        import subprocess
        import warnings
        idx = [int(i) for i in torch.nonzero(host_bad.flatten()).flatten()[:16]]
        scales = torch.tensor([2.0 ** i for i in range(16)])
        y = (mean[..., None, :] * scales[:, None]).flatten(-2)
        arg = torch.cat([y, y + 0.5 * torch.tensor(np.pi)], -1).flatten()
        yv = torch.cat([(cov[..., None, :] * scales[:, None] ** 2).flatten(-2)] * 2, -1).flatten()
        again = O.ipe(mean, cov, 0, 16).flatten()
        recs = [{"flat_index": i, "arg_fp32": float(arg[i]), "exp_arg_fp32": float(-0.5 * yv[i]),
                 "cpu_oracle": float(ref.detach().flatten()[i]), "cpu_oracle_recomputed": float(again[i]),
                 "float64": float(exact.flatten()[i]), "gpu": float(out.cpu().flatten()[i]),
                 "sin_fp32_vectorised": float(torch.sin(arg)[i]), "sin_fp32_single": float(torch.sin(arg[i:i + 1])),
                 "exp_fp32_vectorised": float(torch.exp(-0.5 * yv)[i]), "exp_fp32_single": float(torch.exp(-0.5 * yv[i:i + 1])),
                 "sin_float64": float(torch.sin(arg[i].double())), "exp_float64": float(torch.exp(-0.5 * yv[i].double()))}
                for i in idx]
        info = {"n_host_outliers": int(host_bad.sum()), "max_cpu_err": float(cpu_err.max()), "records": recs,
                "threads": torch.get_num_threads(), "cpu_capability": torch.backends.cpu.get_cpu_capability(),
                "lscpu": subprocess.run("lscpu | grep -E 'Model name|Flags|^CPU\\(s\\)|Hypervisor'", shell=True,
                                        capture_output=True, text=True).stdout[:3000]}
        os.makedirs("gpurun_out", exist_ok=True)
        with open(os.path.join("gpurun_out", f"ipe_cpu_oracle_outlier_{os.getpid()}.json"), "w") as fh:
            json.dump(info, fh, indent=1)
        warnings.warn("host fp32 sin/exp deviates from float64 (CPU oracle, not the GPU kernel): " + json.dumps(info)[:1500])
        assert int(host_bad.sum()) <= 8, "the host's fp32 sin/exp is broadly wrong: " + json.dumps(info)[:1500]
    gpu_vs_cpu = (out.cpu() - ref.detach()).abs()
    assert float(gpu_vs_cpu[~host_bad].max()) < 2e-6, "GPU IPE vs CPU oracle"
    gvec = torch.randn(4096, 96, generator=gen)
    (gref,) = torch.autograd.grad((ref * gvec).sum(), mean_r)
    gout = ops.ipe_vjp(mean.to(DEV), cov.to(DEV), 0, 16, gvec.to(DEV))
    assert_close(gout.cpu(), gref, 2e-5, "ipe_vjp")
    v = torch.randn(4096, 3, generator=gen)
    jv = torch.empty(4096, 96, device=DEV)
    ops.ipe_jvp_into(mean.to(DEV), cov.to(DEV), 0, 16, v.to(DEV), jv)
    # <J v, g> == <v, J^T g>
    lhs = float((jv.cpu().double() * gvec.double()).sum())
    rhs = float((gref.double() * v.double()).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(abs(rhs), 1.0)
    # bf16 output variant writes into a strided buffer
    buf = torch.zeros(4096, 352, device=DEV, dtype=torch.bfloat16)
    ops.ipe_into(mean.to(DEV), cov.to(DEV), 0, 16, buf[:, 256:])
    assert float((buf[:, 256:].float().cpu() - ref.detach()).abs().max()) < 4e-3
    assert float(buf[:, :256].abs().max()) == 0.0


def test_composite_fwd_bwd(golden_ops):
    from panonerf_b200 import ops
    g = golden_ops
    rgb = T(g["vr_rgb"]).to(DEV).requires_grad_()
    den = T(g["vr_density"])[..., 0].contiguous().to(DEV).requires_grad_()
    comp, dist, acc, w = ops.composite(rgb, den, T(g["vr_t"]).to(DEV), T(g["vr_dirs"]).to(DEV), True)
    assert_close(comp.cpu(), T(g["vr_comp"]), 1e-5, "comp")
    assert_close(dist.cpu(), T(g["vr_dist"]), 1e-5, "dist")
    assert_close(acc.cpu(), T(g["vr_acc"]), 1e-5, "acc", floor=1e-3)
    # alpha = 1 - exp(-x) cancels for small x: 1 ulp of expf is 6e-8 ABSOLUTE on a weight in [0,1]
    assert_close(w.cpu(), T(g["vr_weights"]), 1e-5, "weights", floor=2e-2)
    ((comp * T(g["vr_g_comp"]).to(DEV)).sum() + (dist * T(g["vr_g_dist"]).to(DEV)).sum() +
     (acc * T(g["vr_g_acc"]).to(DEV)).sum() + (w * T(g["vr_g_w"]).to(DEV)).sum()).backward()
    assert_close(rgb.grad.cpu(), T(g["vr_d_rgb"]), 1e-5, "d_rgb", floor=2e-2)   # = weights * g_comp
    ref = T(g["vr_d_density"])[..., 0]
    ok = ~torch.isnan(ref).any(dim=1)           # the empty ray: upstream returns NaN, the kernel a finite gradient
    assert torch.isfinite(den.grad).all()
    assert_close(den.grad.cpu()[ok], ref[ok], 2e-5, "d_density", floor=1e-3)


def test_composite_seeded_large():
    from panonerf_b200 import ops
    gen = torch.Generator().manual_seed(5)
    R, N = 2048, 128
    rgb = torch.rand(R, N, 3, generator=gen)
    den = -torch.log(torch.rand(R, N, generator=gen))
    t = torch.sort(torch.rand(R, N + 1, generator=gen) * 10, dim=-1).values
    dirs = torch.randn(R, 3, generator=gen)
    rgb_r, den_r = rgb.clone().requires_grad_(), den.clone().requires_grad_()
    ref = O.composite(rgb_r, den_r[..., None], t, dirs, False)
    rgb_g, den_g = rgb.to(DEV).requires_grad_(), den.to(DEV).requires_grad_()
    out = ops.composite(rgb_g, den_g, t.to(DEV), dirs.to(DEV), False)
    gs = [torch.rand(x.shape, generator=gen) for x in ref]
    sum((a * b).sum() for a, b in zip(ref, gs)).backward()
    sum((a * b.to(DEV)).sum() for a, b in zip(out, gs)).backward()
    for a, b, nm in zip(out, ref, ("comp", "dist", "acc", "w")):
        assert_close(a.cpu(), b.detach(), 1e-5, nm, floor=2e-2)
    assert_close(rgb_g.grad.cpu(), rgb_r.grad, 2e-5, "d_rgb", floor=2e-2)
    assert_close(den_g.grad.cpu(), den_r.grad, 5e-5, "d_density", floor=1e-2)
    # size-independent property: weights are a sub-stochastic partition of unity, acc = sum(w) in [0,1]
    assert float(out[2].max()) <= 1.0 + 1e-6 and float(out[3].min()) >= 0.0


def test_resample_bit_exact_indices(golden_ops):
    from panonerf_b200.models import mip
    g = golden_ops
    r = _rays(g)
    new_t, (m, c), inds = mip.resample_along_rays(r.origins, r.directions, r.radii, T(g["rs_t"]).to(DEV),
                                                  T(g["rs_w"]).to(DEV), False, "cone", True, 0.01, return_inds=True)
    assert torch.equal(inds.cpu(), T(g["rs_inds"])), "searchsorted indices must be bit-exact"
    assert_close(new_t.cpu(), T(g["rs_new_t"]), 1e-5, "new_t")
    assert_close(m.cpu(), T(g["rs_mean"]), 1e-5, "mean")
    assert_close(c.cpu(), T(g["rs_cov"]), 1e-4, "cov", floor=1e-6)
    assert bool((new_t[:, 1:] >= new_t[:, :-1]).all()), "resampled fence-posts must be sorted"
    # randomized: same u -> same samples; plain sorted_piecewise_constant_pdf on pre-blurred weights
    wb = O.blur_weights(T(g["rs_w"]), 0.01).to(DEV)
    out = mip.sorted_piecewise_constant_pdf(T(g["rs_t"]).to(DEV), wb, 17, True, u=T(g["rs_u_r"]).to(DEV))
    assert_close(out.cpu(), T(g["rs_new_t_r"]), 1e-5, "new_t_r")
    # stop_grad=False: same values, and the weights receive a gradient (tests/test_resample_grad_gpu.py checks it)
    wg = T(g["rs_w"]).to(DEV).requires_grad_()
    nt2, (m2, c2) = mip.resample_along_rays(r.origins, r.directions, r.radii, T(g["rs_t"]).to(DEV), wg, False, "cone",
                                            False, 0.01)
    assert torch.equal(nt2.detach(), new_t) and nt2.requires_grad and m2.requires_grad and c2.requires_grad
    (nt2.sum() + m2.sum()).backward()
    assert wg.grad is not None and bool(torch.isfinite(wg.grad).all()) and float(wg.grad.abs().max()) > 0


@pytest.mark.parametrize("n", [64, 128, 256])
def test_resample_seeded(n):
    from panonerf_b200 import ops
    gen = torch.Generator().manual_seed(n)
    R = 4096
    w = torch.rand(R, n, generator=gen) ** 4
    w[0] = 0.0
    w[1] = 0.0
    w[1, n // 2] = 1.0
    t = torch.sort(torch.rand(R, n + 1, generator=gen) * 10, dim=-1).values
    ref, inds_ref, _ = O.pdf_sample(t, O.blur_weights(w, 0.01), n + 1, False, return_aux=True)
    out, inds = ops.resample(t.to(DEV), w.to(DEV), 0.01, return_inds=True)
    # bit-exact: the kernel reproduces torch.sum's CPU reduction order (ATen vectorized_inner_sum) and the fp64 cumsum
    assert torch.equal(inds.cpu(), inds_ref), f"{int((inds.cpu() != inds_ref).sum())} index mismatches"
    assert torch.equal(out.cpu(), ref), "resampled fence-posts differ from the oracle's"
    assert bool((out[:, 1:] >= out[:, :-1]).all())
    assert float(out.min()) >= float(t.min()) and float(out.max()) <= float(t.max())


def test_activations_and_density_grad():
    from panonerf_b200 import ops, _lib
    gen = torch.Generator().manual_seed(1)
    M, C = 5000, 5
    raw_rgb = torch.randn(M, 3, generator=gen) * 8
    raw_den = torch.randn(M, C, generator=gen) * 8
    raw_rgb[0, 0], raw_den[0, 0] = 25.0, 30.0            # softplus threshold branch
    rr, rd = raw_rgb.clone().requires_grad_(), raw_den.clone().requires_grad_()
    rgb = O.softplus(rr) * (1 + 2 * 0.001) - 0.001
    den = O.softplus(rd[..., :1] - 1.0)
    alb = torch.sigmoid(rd[..., 1:-1]) * 0.77 + 0.03
    g1, g2, g3 = torch.randn(M, 3, generator=gen), torch.randn(M, generator=gen), torch.randn(M, 3, generator=gen)
    ((rgb * g1).sum() + (den[:, 0] * g2).sum() + (alb * g3).sum()).backward()
    a, b = raw_rgb.to(DEV).requires_grad_(), raw_den.to(DEV).requires_grad_()
    o1, o2, o3 = ops.activations(a, b, -1.0, 0.001, True)
    assert_close(o1.cpu(), rgb.detach(), 1e-5, "rgb", floor=1e-3)
    assert_close(o2.cpu(), den.detach()[:, 0], 1e-5, "density", floor=1e-3)
    assert_close(o3.cpu(), alb.detach(), 1e-5, "albedo")
    ((o1 * g1.to(DEV)).sum() + (o2 * g2.to(DEV)).sum() + (o3 * g3.to(DEV)).sum()).backward()
    assert_close(a.grad.cpu(), rr.grad, 1e-5, "d_raw_rgb", floor=1e-3)
    assert_close(b.grad.cpu(), rd.grad, 2e-5, "d_raw_den", floor=1e-3)


def test_normals_aggregate_fwd_bwd():
    from panonerf_b200 import ops
    import torch.nn.functional as F
    gen = torch.Generator().manual_seed(2)
    R, N = 300, 64
    n_raw = torch.randn(R, N, 3, generator=gen) * torch.rand(R, N, 1, generator=gen) * 10
    w = torch.rand(R, N, generator=gen) ** 3
    dirs = torch.randn(R, 3, generator=gen)
    albs = torch.rand(R, N, 3, generator=gen)
    a, b, c = n_raw.clone().requires_grad_(), w.clone().requires_grad_(), albs.clone().requires_grad_()
    nh = F.normalize(a, dim=-1)
    nw = b[..., None] / torch.sum(b, -1).view(-1, 1, 1)
    normal = F.normalize(torch.sum(nw * nh, dim=1), dim=-1)
    ort = torch.sum(nw * torch.relu(torch.bmm(nh, dirs.view(-1, 3, 1))) ** 2, dim=1)[:, 0]
    alb = torch.sum(nw * c, dim=1)
    g1, g2, g3 = torch.randn(R, 3, generator=gen), torch.randn(R, generator=gen), torch.randn(R, 3, generator=gen)
    ((normal * g1).sum() + (ort * g2).sum() + (alb * g3).sum()).backward()
    x, y, z = n_raw.to(DEV).requires_grad_(), w.to(DEV).requires_grad_(), albs.to(DEV).requires_grad_()
    o1, o2, o3 = ops.normals_aggregate(x, y, dirs.to(DEV), z)
    assert_close(o1.cpu(), normal.detach(), 2e-5, "normal", floor=1e-2)
    assert_close(o2.cpu(), ort.detach(), 1e-5, "ort", floor=1e-3)
    assert_close(o3.cpu(), alb.detach(), 1e-5, "albedo")
    ((o1 * g1.to(DEV)).sum() + (o2 * g2.to(DEV)).sum() + (o3 * g3.to(DEV)).sum()).backward()
    assert_close(x.grad.cpu(), a.grad, 2e-4, "d_n_raw", floor=float(a.grad.abs().mean()))
    assert_close(y.grad.cpu(), b.grad, 5e-5, "d_weights", floor=float(b.grad.abs().mean()))
    assert_close(z.grad.cpu(), c.grad, 1e-5, "d_albedos", floor=1e-4)


def test_env_cast_and_shade(golden_ops):
    from panonerf_b200 import ops
    g = golden_ops
    gen = torch.Generator().manual_seed(4)
    r = O.Rays(*[T(g[f"raygen_{k}"])[:24] for k in O.Rays._fields])
    env = O.Rays(*[T(g[f"env_{k}"]) for k in O.Rays._fields])
    dist = (torch.rand(24, generator=gen) * 5).requires_grad_()
    pts = r.origins + r.directions * dist.view(-1, 1)
    t_ref, (m_ref, c_ref), _ = O.env_samples(pts, env, 10, False)
    gm = torch.randn(m_ref.shape, generator=gen)
    (m_ref * gm).sum().backward()
    d = dist.detach().to(DEV).requires_grad_()
    f = lambda x: x.to(DEV)
    t, m, c = ops.env_cast(f(r.origins), f(r.directions), d, f(env.directions), f(env.radii), f(env.near), f(env.far), 10)
    assert_close(t.cpu(), t_ref, 1e-6, "t")
    assert_close(m.cpu(), m_ref.detach(), 1e-6, "means")
    assert_close(c.cpu(), c_ref, 1e-5, "covs", floor=1e-6)
    (m * gm.to(DEV)).sum().backward()
    assert_close(d.grad.cpu(), dist.grad, 1e-5, "d_dist", floor=1e-2)
    # shading fwd vs the reference vectors, bwd vs autograd of the oracle
    e, a, n = T(g["sr_env"]).requires_grad_(), T(g["sr_albedo"]).requires_grad_(), T(g["sr_normal"]).requires_grad_()
    rgb_ref, _, shd_ref = O.lambert_shade(e, a, n, T(g["sr_l"]), env.lossmult)
    assert torch.allclose(rgb_ref, T(g["sr_rgb"]), rtol=1e-6)
    g1, g2 = torch.randn(24, 3, generator=gen), torch.randn(24, 3, generator=gen)
    ((rgb_ref * g1).sum() + (shd_ref * g2).sum()).backward()
    e2, a2, n2 = [T(g[k]).to(DEV).requires_grad_() for k in ("sr_env", "sr_albedo", "sr_normal")]
    rgb, shd = ops.shade(e2, a2, n2, f(env.directions), f(env.lossmult).reshape(-1).contiguous())
    assert_close(rgb.cpu(), T(g["sr_rgb"]), 1e-5, "surface_rgb")
    assert_close(shd.cpu(), T(g["sr_shading"]), 1e-5, "shading")
    ((rgb * g1.to(DEV)).sum() + (shd * g2.to(DEV)).sum()).backward()
    assert_close(e2.grad.cpu(), e.grad, 1e-5, "d_env", floor=1e-3)
    assert_close(a2.grad.cpu(), a.grad, 1e-5, "d_albedo", floor=1e-3)
    assert_close(n2.grad.cpu(), n.grad, 1e-5, "d_normal", floor=1e-3)


def test_tonemap_losses(golden_ops):
    from panonerf_b200 import ops
    from panonerf_b200.utils.surface_rendering import hdr_to_ldr
    import torch.nn.functional as F
    g = golden_ops
    x = T(g["tm_in"]).to(DEV)
    assert_close(hdr_to_ldr(x).cpu(), T(g["tm_out"]), 1e-5, "ldr", floor=1e-2)
    q = hdr_to_ldr(x, dtype="uint8").cpu()
    assert float((q - T(g["tm_out_u8"])).abs().max()) < 1e-6, "uint8 quantisation must agree"
    gen = torch.Generator().manual_seed(6)
    R = 1000
    pred = (torch.rand(R, 3, generator=gen) * 3 + 1e-3)
    gt = O.hdr_to_ldr(torch.rand(R, 3, generator=gen) * 2, quantize=True)
    mask = torch.ones(R, 1)
    p = pred.clone().requires_grad_()
    ref = (mask * (O.hdr_to_ldr(p) - gt) ** 2).sum() / mask.sum()
    ref.backward()
    p2 = pred.to(DEV).requires_grad_()
    out = ops.tonemap_mse(p2, gt.to(DEV), mask.reshape(-1).to(DEV), 1.0 / R)
    assert abs(float(out) - float(ref)) < 1e-6 * max(1.0, abs(float(ref)))
    out.backward()
    assert_close(p2.grad.cpu(), p.grad, 2e-5, "d_pred", floor=float(p.grad.abs().mean()))
    alb = torch.rand(R, 3, generator=gen).requires_grad_()
    ref = ((F.normalize(gt, dim=-1) - F.normalize(alb, dim=-1)) ** 2).mean()
    ref.backward()
    a2 = alb.detach().to(DEV).requires_grad_()
    out = ops.chroma_loss(gt.to(DEV), a2)
    assert abs(float(out) - float(ref)) < 1e-6
    out.backward()
    assert_close(a2.grad.cpu(), alb.grad, 2e-5, "d_albedo", floor=float(alb.grad.abs().mean()))


def test_adam_matches_torch():
    from panonerf_b200 import ops
    gen = torch.Generator().manual_seed(7)
    n = 100003
    p0 = torch.randn(n, generator=gen)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=2e-4)
    p = p0.to(DEV)
    m, v = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    for step in range(1, 4):
        g = torch.randn(n, generator=gen)
        ref.grad = g.clone()
        opt.step()
        ops.adam_step(p, g.to(DEV), m, v, 2e-4, step)
    assert_close(p.cpu(), ref.detach(), 1e-6, "adam")


@pytest.mark.parametrize("R,N,C,d_mod,white,want_alb", [(301, 64, 5, 0, False, True), (257, 10, 5, 10, False, False),
                                                        (100, 128, 1, 0, True, False), (33, 256, 5, 0, False, True),
                                                        (64, 7, 1, 0, False, False)])
def test_act_composite_fused_is_bit_identical(R, N, C, d_mod, white, want_alb):
    """pnb_act_composite_fwd/bwd (activations inside the compositing kernels) against the two-kernel sequence
    pnb_act_fwd + pnb_composite_fwd / pnb_composite_bwd + pnb_act_bwd: every output and every gradient bit for bit,
    and the forward against the oracle's compute_graph activations + volumetric_rendering."""
    from panonerf_b200 import ops
    g = torch.Generator().manual_seed(R * 1000 + N)
    raw_rgb = (torch.randn(R * N, 3, generator=g) * 2).to(DEV)
    raw_den = (torch.randn(R * N, C, generator=g) * 3).to(DEV)
    raw_den[::7, 0] = 25.0                                      # the softplus threshold branch
    t = torch.sort(torch.rand(R, N + 1, generator=g) * 6, dim=1).values.contiguous().to(DEV)
    nd = d_mod if d_mod else R
    dirs = torch.randn(nd, 3, generator=g).to(DEV)
    gs = [torch.randn(R, 3, generator=g).to(DEV), torch.randn(R, generator=g).to(DEV),
          torch.randn(R, generator=g).to(DEV), torch.randn(R, N, generator=g).to(DEV),
          torch.randn(R * N, 3, generator=g).to(DEV)]

    def run(fused):
        a, b = raw_rgb.clone().requires_grad_(), raw_den.clone().requires_grad_()
        if fused:
            outs = ops.act_composite(a, b, t, dirs, white, -1.0, 0.001, want_alb, d_mod=d_mod)
        else:
            rgb, den, alb = ops.activations(a, b, -1.0, 0.001, want_alb)
            outs = ops.composite(rgb.view(R, N, 3), den.view(R, N), t, dirs, white, d_mod=d_mod) + (alb,)
        loss = sum((o * gg).sum() for o, gg in zip(outs, gs) if o is not None)
        loss.backward()
        return [o.detach() for o in outs if o is not None], a.grad, b.grad

    of, ga_f, gb_f = run(True)
    ou, ga_u, gb_u = run(False)
    assert len(of) == len(ou) == (5 if want_alb else 4)
    for x, y in zip(of, ou):
        assert torch.equal(x, y)
    assert torch.equal(ga_f, ga_u) and torch.equal(gb_f, gb_u)
    # oracle: activations of compute_graph + volumetric_rendering
    rgb_o = torch.nn.functional.softplus(raw_rgb.cpu()) * (1 + 2 * 0.001) - 0.001
    den_o = torch.nn.functional.softplus(raw_den.cpu()[:, 0] - 1.0)
    d_o = dirs.cpu() if not d_mod else dirs.cpu().repeat((R + d_mod - 1) // d_mod, 1)[:R]
    comp, dist, acc, w = O.composite(rgb_o.view(R, N, 3), den_o.view(R, N, 1), t.cpu(), d_o, white)
    assert_close(of[0].cpu(), comp, 2e-5, "comp_rgb")
    assert_close(of[3].cpu(), w, 2e-5, "weights")
    if want_alb:
        assert_close(of[4].cpu(), torch.sigmoid(raw_den.cpu()[:, 1:4]) * 0.77 + 0.03, 1e-6, "albedo")


def test_no_cpu_fallback():
    from panonerf_b200 import ops
    with pytest.raises(RuntimeError):
        ops.pos_enc(torch.zeros(4, 3), 4)


def test_ipe_fast_path_matches_the_exact_path(monkeypatch):
    """The tiled forward IPE (fixed-point phase, SFU sin/cos, TwoSum-corrected second half) against the kernel that
    range-reduces every feature in double precision, on negative / large means and tiny / huge variances."""
    from panonerf_b200 import ops
    g = torch.Generator().manual_seed(5)
    M = 10007
    mean = (torch.rand(M, 3, generator=g) * 2 - 1) * torch.tensor([0.5, 8.0, 60.0])
    cov = torch.rand(M, 3, generator=g) * torch.tensor([1e-9, 1e-5, 3e-2])
    fast = torch.empty(M, 96, device=DEV)
    ops.ipe_into(mean.to(DEV), cov.to(DEV), 0, 16, fast)
    monkeypatch.setenv("PNB_IPE_SLOW", "1")
    slow = torch.empty(M, 96, device=DEV)
    ops.ipe_into(mean.to(DEV), cov.to(DEV), 0, 16, slow)
    assert float((fast - slow).abs().max()) < 2e-6
    ref = O.ipe(mean, cov, 0, 16)
    assert float((fast.cpu() - ref).abs().max()) < 2e-6


def test_ipe_jvp_tile_kernel(monkeypatch):
    """The tiled IPE jvp kernel (one thread per sample and component, 16-byte rows through shared memory): fp32 rows
    are bit-identical to the per-feature kernel it replaces; bf16 rows (SFU sin / cos / exp2) stay within one bf16 ulp;
    a ragged sample count (not a multiple of the 128-sample tile) and strided rows."""
    from panonerf_b200 import ops
    g = torch.Generator().manual_seed(11)
    M = 10007
    mean = ((torch.rand(M, 3, generator=g) * 2 - 1) * torch.tensor([0.5, 8.0, 60.0])).to(DEV)
    cov = (torch.rand(M, 3, generator=g) * torch.tensor([1e-9, 1e-5, 3e-2])).to(DEV)
    v = torch.randn(M, 3, generator=g).to(DEV)

    def run():
        jv = torch.empty(M, 96, device=DEV)
        ops.ipe_jvp_into(mean, cov, 0, 16, v, jv)
        jw = torch.zeros(M, 352, device=DEV, dtype=torch.bfloat16)
        ops.ipe_jvp_into(mean, cov, 0, 16, v, jw[:, 256:])
        assert float(jw[:, :256].abs().max()) == 0.0
        return jv, jw[:, 256:].float()

    jv, jw = run()
    monkeypatch.setenv("PNB_IPE_SLOW", "1")
    jv_ref, jw_ref = run()
    monkeypatch.delenv("PNB_IPE_SLOW")
    assert torch.equal(jv, jv_ref)
    # one bf16 ulp of the exact rows plus the SFU's ~5e-7 absolute sine error at the scale of the row (2^l v)
    d = (jw - jw_ref).abs()
    assert float((d - 2.0 ** -7 * jw_ref.abs() - 4e-6 * jw_ref.abs().amax(1, keepdim=True)).max()) <= 1e-6
    # <J v, g> == <v, J^T g> with J^T g from autograd through the oracle
    gvec = torch.randn(M, 96, generator=g)
    mean_r = mean.cpu().clone().requires_grad_()
    (gref,) = torch.autograd.grad((O.ipe(mean_r, cov.cpu(), 0, 16) * gvec).sum(), mean_r)
    lhs = float((jv.cpu().double() * gvec.double()).sum())
    rhs = float((gref.double() * v.cpu().double()).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(abs(rhs), 1.0)


def test_device_ray_feed_gathers_the_rays_of_its_pixels():
    """DeviceRayFeed: a sampled batch is exactly the K1 rays / GT colours of the drawn pixel ids,
    over two cameras, and a training step runs on it."""
    from panonerf_b200.datasets.pano_datasets import DeviceRayFeed, generate_rays
    h, w = 16, 32
    cams = []
    for k in range(2):
        c = np.eye(4, dtype=np.float32)
        c[:3, 3] = [0.1 * k, 0.2, 0.3]
        cams.append(c)
    imgs = [torch.rand(h, w, 3, generator=torch.Generator().manual_seed(k)) * 2 for k in range(2)]
    feed = DeviceRayFeed(imgs, cams, 0.0, 10.0, DEV, seed=3)
    assert len(feed) == 2 * h * w
    rays, gt, ids = feed.sample(300, return_ids=True)
    ref = [generate_rays(h, w, c, 0.0, 10.0, DEV) for c in cams]
    for f, name in enumerate(rays._fields):
        full = torch.cat([getattr(r, name) for r in ref], 0)
        assert torch.equal(getattr(rays, name), full[ids]), name
    full_gt = torch.cat([i.reshape(-1, 3) for i in imgs], 0).to(DEV)
    assert torch.equal(gt, full_gt[ids])
    assert int(ids.min()) >= 0 and int(ids.max()) < len(feed) and len(torch.unique(ids)) > 200


def test_empty_batches():
    """Zero rays: every stage returns empty tensors of the reference's shapes (PyTorch ops on empty tensors do), forward
    and backward, including the MLP (both precisions) and a whole model forward."""
    from panonerf_b200 import ops
    from panonerf_b200.models import mip
    from panonerf_b200.models.mip_nerf import MipNeRF
    from panonerf_b200.datasets.base_datasets import Rays
    z = lambda *s: torch.zeros(*s, device=DEV)
    t, (m, c) = mip.sample_along_rays(z(0, 3), z(0, 3), z(0, 1), 16, z(0, 1), z(0, 1), False, False, "cone")
    assert t.shape == (0, 17) and m.shape == (0, 16, 3) and c.shape == (0, 16, 3)
    enc = mip.integrated_pos_enc((m, c), 0, 16)
    assert enc.shape == (0, 16, 96)
    comp, dist, acc, w = ops.composite(z(0, 16, 3), z(0, 16), t, z(0, 3), False)
    assert comp.shape == (0, 3) and w.shape == (0, 16)
    nt, (m2, c2) = mip.resample_along_rays(z(0, 3), z(0, 3), z(0, 1), t, w, False, "cone", True, 0.01)
    assert nt.shape == (0, 17) and m2.shape == (0, 16, 3)
    for prec in ("fp32", "bf16"):
        model = MipNeRF(num_samples=16, rgb_activation="softplus", precision=prec).to(DEV)
        rays = Rays(z(0, 3), z(0, 3), z(0, 3), z(0, 1), z(0, 1), z(0, 1), z(0, 1), z(0, 1))
        out = model(rays=rays, randomized=False, white_bkgd=False, use_ort_loss=False)
        assert out[1][0].shape == (0, 3)
        (out[0][0].sum() + out[1][0].sum()).backward()
        assert all(p.grad is None or float(p.grad.abs().max()) == 0.0 for p in model.mlp.parameters())
