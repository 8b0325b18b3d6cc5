"""stop_resample_grad=False (models/mip.py:336-350): the kernels that carry the fine level's gradient back to the coarse
weights, each against autograd through the oracle (float64 where the function is smooth, fp32 where indices are
involved), and the module-level parity against the reference's own run (tests/golden/*_rg.npz) in test_models_gpu.py."""
import math

import numpy as np
import pytest
import torch

from util import O, assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ipe64(mean, cov):
    """float64 IPE of fp32-representable Gaussians with the reference's fp32 argument y + fl32(pi/2) (mip.py:428): the
    rounding of that sum is locally constant, so it enters as a constant shift."""
    scales = 2.0 ** torch.arange(16, dtype=torch.float64)
    y = (mean[..., None, :] * scales[:, None]).flatten(-2)
    y32 = y.detach().float()
    shift = ((y32 + 0.5 * torch.tensor(np.pi)).double() - y32.double())
    e = torch.exp(-0.5 * (cov[..., None, :] * scales[:, None] ** 2).flatten(-2))
    return torch.cat([e * torch.sin(y), e * torch.sin(y + shift)], -1)


def test_ipe_variance_gradient_and_normal_hessian_terms():
    from panonerf_b200 import ops
    g = torch.Generator().manual_seed(3)
    M = 3001
    mean = ((torch.rand(M, 3, generator=g) * 2 - 1) * torch.tensor([0.5, 4.0, 9.0]))
    cov = torch.rand(M, 3, generator=g) * torch.tensor([1e-8, 1e-5, 3e-3])
    d_enc = torch.randn(M, 96, generator=g)
    h = torch.randn(M, 96, generator=g)
    d_v = torch.randn(M, 3, generator=g)
    m64, c64 = mean.double().requires_grad_(), cov.double().requires_grad_()
    enc = _ipe64(m64, c64)
    (gc_ref,) = torch.autograd.grad((enc * d_enc.double()).sum(), c64, retain_graph=True)
    (v,) = torch.autograd.grad((enc * h.double()).sum(), m64, create_graph=True)
    hm_ref, hc_ref = torch.autograd.grad((v * d_v.double()).sum(), (m64, c64))
    md, cd = mean.to(DEV), cov.to(DEV)
    d_covs = torch.empty(M, 3, device=DEV)
    ops.ipe_cov_hess(md, cd, 0, 16, d_enc=d_enc.to(DEV), d_covs=d_covs)
    assert_close(d_covs.cpu().double(), gc_ref, 2e-5, "d_covs")
    # bf16 rows of d_enc (the tensor-core path hands fp32 rows; the kernel accepts both)
    d16 = d_enc.bfloat16()
    (gc16,) = torch.autograd.grad((_ipe64(m64, c64) * d16.double()).sum(), c64)
    ops.ipe_cov_hess(md, cd, 0, 16, d_enc=d16.to(DEV), d_covs=d_covs)
    assert_close(d_covs.cpu().double(), gc16, 2e-5, "d_covs (bf16 rows)")
    dm = torch.full((M, 3), 2.0, device=DEV)
    dc = torch.full((M, 3), -1.0, device=DEV)
    ops.ipe_cov_hess(md, cd, 0, 16, h_enc=h.to(DEV), d_v=d_v.to(DEV), d_means=dm, d_covs=dc, accumulate=True)
    assert_close((dm.cpu().double() - 2.0), hm_ref, 2e-5, "normal term -> means")
    assert_close((dc.cpu().double() + 1.0), hc_ref, 2e-5, "normal term -> covs")


@pytest.mark.parametrize("R,N", [(257, 64), (33, 17), (40, 256)])
def test_cast_rays_backward(R, N):
    from panonerf_b200 import _lib, ops
    g = torch.Generator().manual_seed(R + N)
    t = torch.sort(torch.rand(R, N + 1, generator=g) * 8 + 0.05, dim=1).values.contiguous()
    d = torch.randn(R, 3, generator=g)
    o = torch.randn(R, 3, generator=g)
    rad = torch.rand(R, 1, generator=g) * 0.01 + 0.001
    gm, gc = torch.randn(R, N, 3, generator=g), torch.randn(R, N, 3, generator=g) * 100
    t64 = t.double().requires_grad_()
    mean, cov = O.cast_cone(t64, o.double(), d.double(), rad.double())
    (ref,) = torch.autograd.grad((mean * gm.double()).sum() + (cov * gc.double()).sum(), t64)
    out = torch.full((R, N + 1), 0.5, device=DEV)
    import ctypes
    p = lambda x: ctypes.c_void_p(x.data_ptr())
    td, dd, rd, gmd, gcd = (x.to(DEV).contiguous() for x in (t, d, rad, gm, gc))
    with torch.cuda.device(0):
        _lib.check(_lib.lib().pnb_cast_rays_bwd(R, N, p(td), p(dd), p(rd), p(gmd), p(gcd), p(out), 1,
                                               ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "cast_rays_bwd")
    assert_close(out.cpu().double() - 0.5, ref, 5e-5, "d_t")


@pytest.mark.parametrize("R,N,C", [(129, 64, 5), (50, 20, 1), (20, 200, 5)])
def test_compositing_gradient_wrt_fence_posts(R, N, C):
    from panonerf_b200 import ops
    g = torch.Generator().manual_seed(R * 7 + N)
    raw_rgb = torch.randn(R * N, 3, generator=g)
    raw_den = torch.randn(R * N, C, generator=g) * 2
    t = torch.sort(torch.rand(R, N + 1, generator=g) * 6 + 0.1, dim=1).values.contiguous()
    dirs = torch.randn(R, 3, generator=g)
    gs = [torch.randn(R, 3, generator=g), torch.randn(R, generator=g), torch.randn(R, generator=g),
          torch.randn(R, N, generator=g)]
    t64 = t.double().requires_grad_()
    rgb_o = torch.nn.functional.softplus(raw_rgb.double()) * (1 + 2 * 0.001) - 0.001
    den_o = torch.nn.functional.softplus(raw_den.double()[:, 0] - 1.0)
    outs = O.composite(rgb_o.view(R, N, 3), den_o.view(R, N, 1), t64, dirs.double(), False)
    (ref,) = torch.autograd.grad(sum((o * gg.double()).sum() for o, gg in zip(outs, gs)), t64)
    td = t.to(DEV).requires_grad_()
    a, b = raw_rgb.to(DEV).requires_grad_(), raw_den.to(DEV).requires_grad_()
    got = ops.act_composite(a, b, td, dirs.to(DEV), False, -1.0, 0.001, False)
    sum((o * gg.to(DEV)).sum() for o, gg in zip(got[:4], gs)).backward()
    assert_close(td.grad.cpu().double(), ref, 5e-5, "d_t")
    # the other gradients are those of the path without a fence-post gradient, bit for bit
    a2, b2 = raw_rgb.to(DEV).requires_grad_(), raw_den.to(DEV).requires_grad_()
    got2 = ops.act_composite(a2, b2, t.to(DEV), dirs.to(DEV), False, -1.0, 0.001, False)
    sum((o * gg.to(DEV)).sum() for o, gg in zip(got2[:4], gs)).backward()
    assert torch.equal(a.grad, a2.grad) and torch.equal(b.grad, b2.grad)


@pytest.mark.parametrize("R,N,randomized,padding", [(300, 64, False, 0.01), (64, 128, True, 0.01), (100, 16, False, 0.01),
                                                    (77, 37, True, 0.0), (16, 256, False, 0.01), (50, 64, False, 1e-9)])
def test_resample_backward(R, N, randomized, padding):
    """pnb_resample_bwd against fp32 autograd through the oracle's blur-pool + PDF sampling (same indices: the forward
    is bit-exact), incl. near-empty rays (the 1e-5 padding branch of mip.py:253-257) and equal neighbours (ties of the
    blur-pool maxima)."""
    from panonerf_b200 import ops
    g = torch.Generator().manual_seed(R + 3 * N)
    w = torch.rand(R, N, generator=g) ** 4
    w[1] = 0.0                                   # empty ray: weight sum below eps when padding is tiny
    w[2, 5:9] = w[2, 5]                          # ties in the blur-pool
    w[3] = 0.0
    w[3, 0] = 1.0                                # all mass in front: the cdf reaches 1 early (ties of min(1, cumsum))
    t = torch.sort(torch.rand(R, N + 1, generator=g) * 6, dim=1).values.contiguous()
    u = None
    if randomized:
        s = 1 / (N + 1)
        u = (torch.arange(N + 1) * s)[None] + torch.rand(R, N + 1, generator=g) * (s - O.F32_EPS)
        u = torch.clamp_max(u, 1.0 - O.F32_EPS).contiguous()
    g_t = torch.randn(R, N + 1, generator=g)
    wr = w.clone().requires_grad_()
    new_t = O.pdf_sample(t, O.blur_weights(wr, padding), N + 1, randomized, u)
    (ref,) = torch.autograd.grad((new_t * g_t).sum(), wr)
    o = torch.zeros(R, 3)
    d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1)
    rad = torch.full((R, 1), 0.003)
    wd = w.to(DEV).requires_grad_()
    nt, means, covs = ops.resample_cast_grad(t.to(DEV), wd, padding, None if u is None else u.to(DEV), o.to(DEV),
                                             d.to(DEV), rad.to(DEV))
    assert torch.equal(nt.detach().cpu(), new_t.detach())
    (nt * g_t.to(DEV)).sum().backward()
    err = float((wd.grad.cpu() - ref).norm() / ref.norm())
    assert err <= 2e-4, err
    # through the Gaussians as well (cast_rays backward chained in front)
    gm, gc = torch.randn(R, N, 3, generator=g), torch.randn(R, N, 3, generator=g) * 10
    wr2 = w.clone().requires_grad_()
    nt2 = O.pdf_sample(t, O.blur_weights(wr2, padding), N + 1, randomized, u)
    m2, c2 = O.cast_cone(nt2, o, d, rad)
    (ref2,) = torch.autograd.grad((m2 * gm).sum() + (c2 * gc).sum() + (nt2 * g_t).sum(), wr2)
    wd2 = w.to(DEV).requires_grad_()
    nt, means, covs = ops.resample_cast_grad(t.to(DEV), wd2, padding, None if u is None else u.to(DEV), o.to(DEV),
                                             d.to(DEV), rad.to(DEV))
    ((means * gm.to(DEV)).sum() + (covs * gc.to(DEV)).sum() + (nt * g_t.to(DEV)).sum()).backward()
    err2 = float((wd2.grad.cpu() - ref2).norm() / ref2.norm())
    assert err2 <= 1e-3, err2
