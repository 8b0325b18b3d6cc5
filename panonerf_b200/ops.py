"""Host-side operators: thin, checked wrappers over the C ABI plus the autograd Functions (explicit forward and
hand-written backward) the models are assembled from.  PyTorch is used here for device memory, streams and the
autograd tape only; every arithmetic step is a kernel of libpanonerf_b200.so.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import os

import torch

from . import _lib
from ._lib import PNB_BF16, PNB_F32, check

_VP = ctypes.c_void_p

# The reference trains under Lightning's `precision='16-mixed'` (train.py:86): torch.autocast(fp16) + GradScaler.  Every
# Function below runs its kernels in the precision it was built for, so under autocast floating-point inputs are cast
# to fp32 and autocast is switched off inside forward / backward (SURVEY.md section 8b "precision context").
_amp_fwd = torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
_amp_bwd = torch.amp.custom_bwd(device_type="cuda")


def _p(t: Optional[torch.Tensor]):
    return None if t is None else _VP(t.data_ptr())


def _stream():
    return _VP(torch.cuda.current_stream().cuda_stream)


def _req(t: torch.Tensor, name: str, dtype=torch.float32):
    """Inputs must already live on the GPU in the layout the kernels read (no silent CPU fallback, no copies)."""
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"panonerf_b200: `{name}` must be a CUDA tensor (there is no CPU path)")
    if t.dtype != dtype:
        raise RuntimeError(f"panonerf_b200: `{name}` must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"panonerf_b200: `{name}` must be contiguous")
    return t


def _f32c(t: torch.Tensor) -> torch.Tensor:
    """Normalise user-facing ray tensors (fp16 env rays, fp64 radii, expanded views) to contiguous fp32."""
    if not t.is_cuda:
        raise RuntimeError("panonerf_b200: tensors must be CUDA tensors (there is no CPU path)")
    return t.detach().to(torch.float32).contiguous() if (t.dtype != torch.float32 or not t.is_contiguous()) else t


def dt_code(dtype) -> int:
    return PNB_BF16 if dtype == torch.bfloat16 else PNB_F32


def launch_count() -> int:
    return int(_lib.lib().pnb_launch_count())


# --------------------------------------------------------------------------------------------------------------
# rays / sampling / encodings (no autograd needed: inputs are data)
# --------------------------------------------------------------------------------------------------------------
_lin_cache = {}


def linspace01(n: int, device) -> torch.Tensor:
    """torch.linspace(0, 1, n) on `device`, computed by torch once so that its rounding matches the reference."""
    key = ("lin", n, str(device))
    if key not in _lin_cache:
        _lin_cache[key] = torch.linspace(0.0, 1.0, n, device=device)
    return _lin_cache[key]


def linspace_u(n: int, device) -> torch.Tensor:
    """torch.linspace(0, 1 - eps, n) of models/mip.py:279."""
    key = ("u", n, str(device))
    if key not in _lin_cache:
        _lin_cache[key] = torch.linspace(0.0, 1.0 - torch.finfo(torch.float32).eps, n, device=device)
    return _lin_cache[key]


def raygen_equirect(h: int, w: int, c2w, near: float, far: float, device, row0: int = 0, nrows: Optional[int] = None):
    nrows = h - row0 if nrows is None else nrows
    n = nrows * w
    import numpy as np
    c = np.ascontiguousarray(np.asarray(c2w, dtype=np.float32)[:3, :4])
    c44 = np.zeros((3, 4), dtype=np.float32)
    c44[:] = c
    mk = lambda k: torch.empty(n, k, device=device, dtype=torch.float32)
    o, d, v = mk(3), mk(3), mk(3)
    rad, lm, nr, fr, nv = mk(1), mk(1), mk(1), mk(1), mk(1)
    with torch.cuda.device(o.device):
        check(_lib.lib().pnb_raygen_equirect(h, w, row0, nrows, c44.ctypes.data_as(_VP), float(near), float(far),
                                             _p(o), _p(d), _p(v), _p(rad), _p(lm), _p(nr), _p(fr), _p(nv),
                                             _stream()), "raygen_equirect")
    return o, d, v, rad, lm, nr, fr, nv


def sample_cast(origins, directions, radii, near, far, n: int, t_rand=None, disparity=False, o_div=1, d_mod=0,
                n_rays=None, rand_shared=False):
    r = origins.shape[0] * o_div if n_rays is None else n_rays
    dev = origins.device
    t = torch.empty(r, n + 1, device=dev, dtype=torch.float32)
    means = torch.empty(r, n, 3, device=dev, dtype=torch.float32)
    covs = torch.empty(r, n, 3, device=dev, dtype=torch.float32)
    if t_rand is not None:
        _req(t_rand, "t_rand")
    with torch.cuda.device(dev):
        check(_lib.lib().pnb_sample_cast(r, n, _p(_req(origins, "origins")), o_div, _p(_req(directions, "directions")),
                                         _p(_req(radii, "radii")), _p(_req(near, "near")), _p(_req(far, "far")), d_mod,
                                         _p(linspace01(n + 1, dev)), _p(t_rand),
                                         0 if (t_rand is None or rand_shared) else n + 1, int(bool(disparity)),
                                         _p(t), _p(means), _p(covs), _stream()), "sample_cast")
    return t, means, covs


def cast_rays_t(t, origins, directions, radii, o_div=1, d_mod=0):
    r, n = t.shape[0], t.shape[1] - 1
    means = torch.empty(r, n, 3, device=t.device, dtype=torch.float32)
    covs = torch.empty(r, n, 3, device=t.device, dtype=torch.float32)
    with torch.cuda.device(t.device):
        check(_lib.lib().pnb_cast_rays(r, n, _p(_req(t, "t")), _p(_req(origins, "origins")), o_div,
                                       _p(_req(directions, "directions")), _p(_req(radii, "radii")), d_mod, _p(means),
                                       _p(covs), _stream()), "cast_rays")
    return means, covs


def ipe_into(means, covs, min_deg, max_deg, out: torch.Tensor):
    """IPE written into `out` ([M, 6L] view, any row stride, fp32 or bf16)."""
    m = means.numel() // 3
    with torch.cuda.device(means.device):
        check(_lib.lib().pnb_ipe_fwd(m, _p(_req(means, "means")), _p(_req(covs, "covs")), min_deg, max_deg,
                                     _p(out), out.stride(0), dt_code(out.dtype), _stream()), "ipe_fwd")
    return out


def ipe_vjp(means, covs, min_deg, max_deg, d_enc: torch.Tensor):
    m = means.numel() // 3
    out = torch.empty(m, 3, device=means.device, dtype=torch.float32)
    with torch.cuda.device(means.device):
        check(_lib.lib().pnb_ipe_vjp(m, _p(means), _p(covs), min_deg, max_deg, _p(d_enc), d_enc.stride(0),
                                     dt_code(d_enc.dtype), _p(out), _stream()), "ipe_vjp")
    return out


def ipe_jvp_into(means, covs, min_deg, max_deg, v, out: torch.Tensor):
    m = means.numel() // 3
    with torch.cuda.device(means.device):
        check(_lib.lib().pnb_ipe_jvp(m, _p(means), _p(covs), min_deg, max_deg, _p(_req(v, "v")), _p(out),
                                     out.stride(0), dt_code(out.dtype), _stream()), "ipe_jvp")
    return out


def pos_enc(x, deg: int):
    x = _req(x, "x")
    out = torch.empty(x.shape[0], 3 + 6 * deg, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        check(_lib.lib().pnb_pos_enc(x.shape[0], _p(x), deg, _p(out), _stream()), "pos_enc")
    return out


def resample(t, weights, padding: float, u=None, return_inds=False, blur_pool=True, cast=None):
    """models/mip.py:304-352 (stop_grad branch): new fence-posts [R,N+1] (+ searchsorted indices for tests).
    `cast=(origins, directions, radii)` also runs cast_rays on the new fence-posts in the same kernel and appends
    (means, covs) to the result."""
    r, n = weights.shape
    dev = t.device
    if u is None:
        u, u_ld = linspace_u(n + 1, dev), 0
    else:
        u, u_ld = _req(u, "u"), n + 1
    new_t = torch.empty(r, n + 1, device=dev, dtype=torch.float32)
    inds = torch.empty(r, n + 1, device=dev, dtype=torch.int64) if return_inds else None
    o = d = rad = means = covs = None
    if cast is not None:
        o, d, rad = (_req(x, nm) for x, nm in zip(cast, ("origins", "directions", "radii")))
        means = torch.empty(r, n, 3, device=dev, dtype=torch.float32)
        covs = torch.empty(r, n, 3, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        check(_lib.lib().pnb_resample_cast(r, n, _p(_req(t, "t")), _p(_req(weights, "weights")), float(padding),
                                           int(blur_pool), _p(u), u_ld, _p(new_t), _p(inds), _p(o), _p(d), _p(rad),
                                           _p(means), _p(covs), _stream()), "resample")
    out = (new_t,) + ((inds,) if return_inds else ()) + ((means, covs) if cast is not None else ())
    return out if len(out) > 1 else new_t


class _ResampleCast(torch.autograd.Function):
    """resample_along_rays with stop_grad=False (the else branch of models/mip.py:336-350): blur-pool + PDF sampling +
    cast_rays in one launch; the backward pulls the gradients of the new fence-posts (direct, and through the
    Gaussians: pnb_cast_rays_bwd) back to the coarse weights (pnb_resample_bwd).  Bins, origins, directions, radii and
    the uniform draws receive no gradient (upstream's bins are the coarse fence-posts, which are off the tape)."""

    @staticmethod
    @_amp_fwd
    def forward(ctx, weights, t, padding, u, origins, directions, radii):
        new_t, means, covs = resample(t, weights, padding, u=u, cast=(origins, directions, radii))
        ctx.save_for_backward(weights, t, u if u is not None else torch.empty(0, device=t.device), new_t, directions,
                              radii)
        ctx.cfg = (float(padding), u is not None)
        return new_t, means, covs

    @staticmethod
    @_amp_bwd
    def backward(ctx, g_t, g_means, g_covs):
        weights, t, u, new_t, directions, radii = ctx.saved_tensors
        padding, has_u = ctx.cfg
        r, n = weights.shape
        dev = weights.device
        cg = lambda g: None if g is None else g.contiguous()
        g_t, g_means, g_covs = cg(g_t), cg(g_means), cg(g_covs)
        if g_t is None and g_means is None and g_covs is None:
            return (None,) * 7
        g_total = g_t.clone() if g_t is not None else torch.zeros(r, n + 1, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            if g_means is not None or g_covs is not None:
                check(_lib.lib().pnb_cast_rays_bwd(r, n, _p(new_t), _p(directions), _p(radii), _p(g_means), _p(g_covs),
                                                   _p(g_total), 1, _stream()), "cast_rays_bwd")
            d_w = torch.empty_like(weights)
            uu, u_ld = (u, n + 1) if has_u else (linspace_u(n + 1, dev), 0)
            check(_lib.lib().pnb_resample_bwd(r, n, _p(t), _p(weights), padding, 1, _p(uu), u_ld, _p(g_total), _p(d_w),
                                              _stream()), "resample_bwd")
        return d_w, None, None, None, None, None, None


def resample_cast_grad(t, weights, padding, u, origins, directions, radii):
    """Differentiable (w.r.t. `weights`) resample_along_rays: -> new_t, means, covs."""
    return _ResampleCast.apply(_req(weights, "weights"), _req(t, "t"), padding, u, _req(origins, "origins"),
                               _req(directions, "directions"), _req(radii, "radii"))


def ipe_cov_hess(means, covs, min_deg, max_deg, d_enc=None, h_enc=None, d_v=None, d_means=None, d_covs=None,
                 accumulate=False):
    """pnb_ipe_cov_hess: the variance gradient of the encoding and the explicit second-order terms of the
    density-gradient normals (include/panonerf_b200.h)."""
    m = means.numel() // 3
    any_rows = d_enc if d_enc is not None else h_enc
    with torch.cuda.device(means.device):
        check(_lib.lib().pnb_ipe_cov_hess(m, _p(means), _p(covs), min_deg, max_deg, _p(d_enc),
                                          d_enc.stride(0) if d_enc is not None else 0,
                                          dt_code(d_enc.dtype) if d_enc is not None else dt_code(any_rows.dtype),
                                          _p(h_enc), h_enc.stride(0) if h_enc is not None else 0, _p(d_v), _p(d_means),
                                          _p(d_covs), int(accumulate), _stream()), "ipe_cov_hess")


def hdr_to_ldr(x, quantize=False):
    x = _req(x, "x")
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(_lib.lib().pnb_hdr_to_ldr(x.numel(), _p(x), int(quantize), _p(out), _stream()), "hdr_to_ldr")
    return out


def dsum(x, scale=1.0):
    """Deterministic sum of an fp32 tensor -> 0-dim tensor."""
    x = _req(x, "x")
    out = torch.empty(1, device=x.device, dtype=torch.float32)
    ws = torch.empty(1024, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        check(_lib.lib().pnb_sum(x.numel(), _p(x), float(scale), _p(out), _p(ws), _stream()), "sum")
    return out[0]


def adam_step(p, g, m, v, lr, step, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0):
    with torch.cuda.device(p.device):
        check(_lib.lib().pnb_adam_step(p.numel(), _p(_req(p, "p")), _p(_req(g, "g")), _p(_req(m, "m")), _p(_req(v, "v")),
                                       float(lr), beta1, beta2, eps, int(step), float(grad_scale), _stream()), "adam")


def adam_step_dev(p, g, m, v, hyper, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0):
    """Adam update with {lr, 1-beta1^t, sqrt(1-beta2^t)} read from the device tensor `hyper` (CUDA-graph friendly)."""
    with torch.cuda.device(p.device):
        check(_lib.lib().pnb_adam_step_dev(p.numel(), _p(_req(p, "p")), _p(_req(g, "g")), _p(_req(m, "m")),
                                           _p(_req(v, "v")), _p(_req(hyper, "hyper")), beta1, beta2, eps,
                                           float(grad_scale), _stream()), "adam_dev")


# --------------------------------------------------------------------------------------------------------------
# autograd Functions (explicit forward / backward kernels)
# --------------------------------------------------------------------------------------------------------------
class _Act(torch.autograd.Function):
    """rgb/density/albedo activations of compute_graph (models/pano_mip_nerf.py:264-278)."""

    @staticmethod
    @_amp_fwd
    def forward(ctx, raw_rgb, raw_den, density_bias, rgb_padding, want_albedo):
        m, c = raw_den.shape
        dev = raw_den.device
        rgb = torch.empty(m, 3, device=dev, dtype=torch.float32)
        den = torch.empty(m, device=dev, dtype=torch.float32)
        alb = torch.empty(m, 3, device=dev, dtype=torch.float32) if want_albedo else None
        with torch.cuda.device(dev):
            check(_lib.lib().pnb_act_fwd(m, c, _p(_req(raw_rgb, "raw_rgb")), _p(_req(raw_den, "raw_den")),
                                         float(density_bias), float(rgb_padding), _p(rgb), _p(den), _p(alb),
                                         _stream()), "act_fwd")
        ctx.save_for_backward(raw_rgb, raw_den)
        ctx.cfg = (float(density_bias), float(rgb_padding))
        if want_albedo:
            return rgb, den, alb
        return rgb, den, None

    @staticmethod
    @_amp_bwd
    def backward(ctx, d_rgb, d_den, d_alb):
        raw_rgb, raw_den = ctx.saved_tensors
        m, c = raw_den.shape
        d_raw_rgb = torch.empty_like(raw_rgb)
        d_raw_den = torch.empty_like(raw_den)
        cg = lambda g: None if g is None else g.contiguous()
        d_rgb, d_den, d_alb = cg(d_rgb), cg(d_den), cg(d_alb)
        with torch.cuda.device(raw_den.device):
            check(_lib.lib().pnb_act_bwd(m, c, _p(raw_rgb), _p(raw_den), ctx.cfg[0], ctx.cfg[1], _p(d_rgb), _p(d_den),
                                         _p(d_alb), _p(d_raw_rgb), _p(d_raw_den), _stream()), "act_bwd")
        return d_raw_rgb, d_raw_den, None, None, None


def activations(raw_rgb, raw_den, density_bias, rgb_padding, want_albedo):
    return _Act.apply(raw_rgb, raw_den, density_bias, rgb_padding, want_albedo)


class _Composite(torch.autograd.Function):
    """volumetric_rendering (models/mip.py:444-483)."""

    @staticmethod
    @_amp_fwd
    def forward(ctx, rgb, density, t, dirs, d_mod, white_bkgd):
        r, n = density.shape
        dev = rgb.device
        comp = torch.empty(r, 3, device=dev, dtype=torch.float32)
        dist = torch.empty(r, device=dev, dtype=torch.float32)
        acc = torch.empty(r, device=dev, dtype=torch.float32)
        w = torch.empty(r, n, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            check(_lib.lib().pnb_composite_fwd(r, n, _p(_req(rgb, "rgb")), _p(_req(density, "density")), _p(_req(t, "t")),
                                               _p(_req(dirs, "dirs")), d_mod, int(white_bkgd), _p(comp), _p(dist),
                                               _p(acc), _p(w), _stream()), "composite_fwd")
        ctx.save_for_backward(rgb, density, t, dirs)
        ctx.cfg = (d_mod, int(white_bkgd))
        return comp, dist, acc, w

    @staticmethod
    @_amp_bwd
    def backward(ctx, g_comp, g_dist, g_acc, g_w):
        rgb, density, t, dirs = ctx.saved_tensors
        r, n = density.shape
        d_rgb = torch.empty_like(rgb)
        d_den = torch.empty_like(density)
        cg = lambda g: None if g is None else g.contiguous()
        g_comp, g_dist, g_acc, g_w = cg(g_comp), cg(g_dist), cg(g_acc), cg(g_w)
        with torch.cuda.device(rgb.device):
            check(_lib.lib().pnb_composite_bwd(r, n, _p(rgb), _p(density), _p(t), _p(dirs), ctx.cfg[0], ctx.cfg[1],
                                               _p(g_comp), _p(g_dist), _p(g_acc), _p(g_w), _p(d_rgb), _p(d_den),
                                               _stream()), "composite_bwd")
        return d_rgb, d_den, None, None, None, None


def composite(rgb, density, t, dirs, white_bkgd, d_mod=0, attenuate=False):
    """rgb [R,N,3], density [R,N], t [R,N+1], dirs [R,3] (or [D,3] with d_mod=D) -> comp, distance, acc, weights.
    `attenuate`: the 1 / (1 + t_mid^2) colour attenuation of volumetric_lighting_composing (models/mip.py:486-527)."""
    return _Composite.apply(rgb, density, t, dirs, d_mod, int(bool(white_bkgd)) | (2 if attenuate else 0))


class _ActComposite(torch.autograd.Function):
    """compute_graph's activations (models/pano_mip_nerf.py:264-278) + volumetric_rendering (models/mip.py:444-483) in
    one kernel each way: the activated colours / densities never reach HBM.  Bit-identical to _Act followed by
    _Composite."""

    @staticmethod
    @_amp_fwd
    def forward(ctx, raw_rgb, raw_den, t, dirs, r, n, d_mod, white_bkgd, density_bias, rgb_padding, want_albedo):
        m, c = raw_den.shape
        dev = raw_den.device
        comp = torch.empty(r, 3, device=dev, dtype=torch.float32)
        dist = torch.empty(r, device=dev, dtype=torch.float32)
        acc = torch.empty(r, device=dev, dtype=torch.float32)
        w = torch.empty(r, n, device=dev, dtype=torch.float32)
        alb = torch.empty(m, 3, device=dev, dtype=torch.float32) if want_albedo else None
        with torch.cuda.device(dev):
            check(_lib.lib().pnb_act_composite_fwd(r, n, c, _p(_req(raw_rgb, "raw_rgb")), _p(_req(raw_den, "raw_den")),
                                                   float(density_bias), float(rgb_padding), _p(_req(t, "t")),
                                                   _p(_req(dirs, "dirs")), d_mod, int(white_bkgd), _p(comp), _p(dist),
                                                   _p(acc), _p(w), _p(alb), _stream()), "act_composite_fwd")
        ctx.save_for_backward(raw_rgb, raw_den, t, dirs)
        ctx.cfg = (r, n, d_mod, int(white_bkgd), float(density_bias), float(rgb_padding))
        return comp, dist, acc, w, alb

    @staticmethod
    @_amp_bwd
    def backward(ctx, g_comp, g_dist, g_acc, g_w, g_alb):
        raw_rgb, raw_den, t, dirs = ctx.saved_tensors
        r, n, d_mod, white, bias, pad = ctx.cfg
        d_raw_rgb = torch.empty_like(raw_rgb)
        d_raw_den = torch.empty_like(raw_den)
        # fence-post gradient: only with stop_resample_grad=False (the resampled t is then on the tape)
        d_t = torch.empty_like(t) if ctx.needs_input_grad[2] else None
        cg = lambda g: None if g is None else g.contiguous()
        g_comp, g_dist, g_acc, g_w, g_alb = cg(g_comp), cg(g_dist), cg(g_acc), cg(g_w), cg(g_alb)
        with torch.cuda.device(raw_den.device):
            check(_lib.lib().pnb_act_composite_bwd(r, n, raw_den.shape[1], _p(raw_rgb), _p(raw_den), bias, pad, _p(t),
                                                   _p(dirs), d_mod, white, _p(g_comp), _p(g_dist), _p(g_acc), _p(g_w),
                                                   _p(g_alb), _p(d_raw_rgb), _p(d_raw_den), _p(d_t), _stream()),
                  "act_composite_bwd")
        return (d_raw_rgb, d_raw_den, d_t) + (None,) * 8


def act_composite(raw_rgb, raw_den, t, dirs, white_bkgd, density_bias, rgb_padding, want_albedo, d_mod=0):
    """raw_rgb [R*N,3], raw_den [R*N,C], t [R,N+1], dirs [R,3] (or [D,3] with d_mod=D) -> comp, distance, acc, weights,
    albedos ([R*N,3] or None).  One launch when N <= 256 (the fused kernel), otherwise activations() + composite()."""
    r, n = t.shape[0], t.shape[1] - 1
    if t.requires_grad and n > 256:
        raise NotImplementedError("stop_resample_grad=False needs the fused activation + compositing kernels (N <= 256)")
    if (n > 256 or os.environ.get("PNB_UNFUSED_ACT") == "1") and not t.requires_grad:
        rgb, den, alb = activations(raw_rgb, raw_den, density_bias, rgb_padding, want_albedo)
        comp, dist, acc, w = composite(rgb.view(r, n, 3), den.view(r, n), t, dirs, white_bkgd, d_mod=d_mod)
        return comp, dist, acc, w, alb
    return _ActComposite.apply(raw_rgb, raw_den, t, dirs, r, n, d_mod, int(bool(white_bkgd)), density_bias,
                               rgb_padding, bool(want_albedo))


class _Normals(torch.autograd.Function):
    """Weighted, normalised density-gradient normals + orientation loss + albedo compositing
    (models/pano_mip_nerf.py:296-317, models/mip_nerf.py:258-277)."""

    @staticmethod
    @_amp_fwd
    def forward(ctx, n_raw, weights, dirs, albedos):
        r, n = weights.shape
        dev = weights.device
        normal = torch.empty(r, 3, device=dev, dtype=torch.float32)
        ort = torch.empty(r, device=dev, dtype=torch.float32)
        alb = torch.empty(r, 3, device=dev, dtype=torch.float32) if albedos is not None else None
        with torch.cuda.device(dev):
            check(_lib.lib().pnb_normals_fwd(r, n, _p(_req(n_raw, "n_raw")), _p(_req(weights, "weights")),
                                             _p(_req(dirs, "dirs")), _p(albedos), _p(normal), _p(ort), _p(alb),
                                             _stream()), "normals_fwd")
        ctx.save_for_backward(n_raw, weights, dirs, albedos)
        return normal, ort, alb

    @staticmethod
    @_amp_bwd
    def backward(ctx, g_normal, g_ort, g_alb):
        n_raw, weights, dirs, albedos = ctx.saved_tensors
        r, n = weights.shape
        d_n = torch.empty_like(n_raw)
        d_w = torch.empty_like(weights)
        d_a = torch.empty_like(albedos) if albedos is not None else None
        cg = lambda g: None if g is None else g.contiguous()
        g_normal, g_ort, g_alb = cg(g_normal), cg(g_ort), cg(g_alb)
        with torch.cuda.device(weights.device):
            check(_lib.lib().pnb_normals_bwd(r, n, _p(n_raw), _p(weights), _p(dirs), _p(albedos), _p(g_normal),
                                             _p(g_ort), _p(g_alb), _p(d_n), _p(d_w), _p(d_a), _stream()), "normals_bwd")
        return d_n, d_w, None, d_a


def normals_aggregate(n_raw, weights, dirs, albedos=None):
    return _Normals.apply(n_raw, weights, dirs, albedos)


class _EnvCast(torch.autograd.Function):
    """Surface point o + d*distance (models/pano_mip_nerf.py:321-324) and the env-ray Gaussians of
    sample_each_points (models/mip.py:154-194).  Differentiable w.r.t. `distance` (detach_dist=False upstream)."""

    @staticmethod
    @_amp_fwd
    def forward(ctx, origins, dirs, distance, env_dirs, env_radii, env_near, env_far, n_env, t_rand):
        r, d = origins.shape[0], env_dirs.shape[0]
        dev = origins.device
        pts = torch.empty(r, 3, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            check(_lib.lib().pnb_surface_point_fwd(r, _p(_req(origins, "origins")), _p(_req(dirs, "dirs")),
                                                   _p(_req(distance, "distance")), _p(pts), _stream()), "surface_point")
        t, means, covs = sample_cast(pts, env_dirs, env_radii, env_near, env_far, n_env, t_rand=t_rand, o_div=d,
                                     d_mod=d, n_rays=r * d, rand_shared=True)
        ctx.save_for_backward(dirs)
        ctx.k = d * n_env
        ctx.mark_non_differentiable(t, covs)
        return t, means, covs

    @staticmethod
    @_amp_bwd
    def backward(ctx, g_t, g_means, g_covs):
        (dirs,) = ctx.saved_tensors
        r = dirs.shape[0]
        d_dist = torch.empty(r, device=dirs.device, dtype=torch.float32)
        g_means = g_means.contiguous()
        with torch.cuda.device(dirs.device):
            check(_lib.lib().pnb_surface_point_bwd(r, ctx.k, _p(dirs), _p(g_means), _p(d_dist), _stream()),
                  "surface_point_bwd")
        return None, None, d_dist, None, None, None, None, None, None


def env_cast(origins, dirs, distance, env_dirs, env_radii, env_near, env_far, n_env, t_rand=None):
    return _EnvCast.apply(origins, dirs, distance, env_dirs, env_radii, env_near, env_far, n_env, t_rand)


class _Shade(torch.autograd.Function):
    """Lambertian surface rendering (utils/surface_rendering.py:104-165, roughness=None branch)."""

    @staticmethod
    @_amp_fwd
    def forward(ctx, env_rgb, albedo, normal, light_dirs, solid_angle):
        r, d = env_rgb.shape[0], env_rgb.shape[1]
        dev = env_rgb.device
        rgb = torch.empty(r, 3, device=dev, dtype=torch.float32)
        shading = torch.empty(r, 3, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            check(_lib.lib().pnb_shade_fwd(r, d, _p(_req(env_rgb, "env_rgb")), _p(_req(albedo, "albedo")),
                                           _p(_req(normal, "normal")), _p(_req(light_dirs, "light_dirs")),
                                           _p(_req(solid_angle, "solid_angle")), _p(rgb), _p(shading), _stream()),
                  "shade_fwd")
        ctx.save_for_backward(env_rgb, albedo, normal, light_dirs, solid_angle)
        return rgb, shading

    @staticmethod
    @_amp_bwd
    def backward(ctx, g_rgb, g_shading):
        env_rgb, albedo, normal, light_dirs, solid_angle = ctx.saved_tensors
        r, d = env_rgb.shape[0], env_rgb.shape[1]
        d_env, d_alb, d_nrm = torch.empty_like(env_rgb), torch.empty_like(albedo), torch.empty_like(normal)
        cg = lambda g: None if g is None else g.contiguous()
        g_rgb, g_shading = cg(g_rgb), cg(g_shading)
        with torch.cuda.device(env_rgb.device):
            check(_lib.lib().pnb_shade_bwd(r, d, _p(env_rgb), _p(albedo), _p(normal), _p(light_dirs), _p(solid_angle),
                                           _p(g_rgb), _p(g_shading), _p(d_env), _p(d_alb), _p(d_nrm), _stream()),
                  "shade_bwd")
        return d_env, d_alb, d_nrm, None, None


def shade(env_rgb, albedo, normal, light_dirs, solid_angle):
    return _Shade.apply(env_rgb, albedo, normal, light_dirs, solid_angle)


# ---- variants the upstream hot path does not call today (csrc/variants.cu) ----------
BRDF_MICROFACET, BRDF_BLINN_PHONG = 0, 1


class _BrdfTerms(torch.autograd.Function):
    """Specular ratio and NoL per (ray, light direction): `microfeast_brdf` / `blinn_phong_brdf`
    (utils/surface_rendering.py:6-61, 64-101).  Differentiable w.r.t. normal and roughness."""

    @staticmethod
    @_amp_fwd
    def forward(ctx, kind, normal, roughness, l, v):
        r = normal.shape[0]
        per_ray = int(l.dim() == 3)
        d = l.shape[-2]
        spec = torch.empty(r, d, device=normal.device, dtype=torch.float32)
        nol = torch.empty(r, d, device=normal.device, dtype=torch.float32)
        with torch.cuda.device(normal.device):
            check(_lib.lib().pnb_brdf_terms_fwd(kind, r, d, _p(_req(normal, "normal")), _p(_req(roughness, "roughness")),
                                                _p(_req(l, "l")), per_ray, _p(_req(v, "v")), _p(spec), _p(nol),
                                                _stream()), "brdf_terms_fwd")
        ctx.save_for_backward(normal, roughness, l, v)
        ctx.cfg = (kind, per_ray, d)
        return spec, nol

    @staticmethod
    @_amp_bwd
    def backward(ctx, g_spec, g_nol):
        normal, roughness, l, v = ctx.saved_tensors
        kind, per_ray, d = ctx.cfg
        d_n, d_r = torch.empty_like(normal), torch.empty_like(roughness)
        cg = lambda g: None if g is None else g.contiguous()
        g_spec, g_nol = cg(g_spec), cg(g_nol)
        with torch.cuda.device(normal.device):
            check(_lib.lib().pnb_brdf_terms_bwd(kind, normal.shape[0], d, _p(normal), _p(roughness), _p(l), per_ray,
                                                _p(v), _p(g_spec), _p(g_nol), _p(d_n), _p(d_r), _stream()),
                  "brdf_terms_bwd")
        return None, d_n, d_r, None, None


def brdf_terms(kind, normal, roughness, l, v):
    """normal [R,3], roughness [R], l [D,3] or [R,D,3], v [R,3] -> spec [R,D], NoL [R,D]."""
    return _BrdfTerms.apply(kind, normal, roughness, l, v)


class _ShadeSum(torch.autograd.Function):
    """diffuse / specular sums of the `roughness is not None` branch of surface_rendering
    (utils/surface_rendering.py:147-151,159)."""

    @staticmethod
    @_amp_fwd
    def forward(ctx, env_rgb, albedo, spec, nol, solid_angle):
        r, d = env_rgb.shape[0], env_rgb.shape[1]
        dev = env_rgb.device
        rgb, dif, spc = (torch.empty(r, 3, device=dev, dtype=torch.float32) for _ in range(3))
        with torch.cuda.device(dev):
            check(_lib.lib().pnb_shade_sum_fwd(r, d, _p(_req(env_rgb, "env_rgb")), _p(_req(albedo, "albedo")),
                                               _p(_req(spec, "spec")), _p(_req(nol, "nol")),
                                               _p(_req(solid_angle, "solid_angle")), _p(rgb), _p(dif), _p(spc),
                                               _stream()), "shade_sum_fwd")
        ctx.save_for_backward(env_rgb, albedo, spec, nol, solid_angle)
        return rgb, dif, spc

    @staticmethod
    @_amp_bwd
    def backward(ctx, g_rgb, g_dif, g_spc):
        env_rgb, albedo, spec, nol, solid_angle = ctx.saved_tensors
        r, d = env_rgb.shape[0], env_rgb.shape[1]
        d_env, d_alb = torch.empty_like(env_rgb), torch.empty_like(albedo)
        d_spec, d_nol = torch.empty_like(spec), torch.empty_like(nol)
        cg = lambda g: None if g is None else g.contiguous()
        g_rgb, g_dif, g_spc = cg(g_rgb), cg(g_dif), cg(g_spc)
        with torch.cuda.device(env_rgb.device):
            check(_lib.lib().pnb_shade_sum_bwd(r, d, _p(env_rgb), _p(albedo), _p(spec), _p(nol), _p(solid_angle),
                                               _p(g_rgb), _p(g_dif), _p(g_spc), _p(d_env), _p(d_alb), _p(d_spec),
                                               _p(d_nol), _stream()), "shade_sum_bwd")
        return d_env, d_alb, d_spec, d_nol, None


def shade_sum(env_rgb, albedo, spec, nol, solid_angle):
    return _ShadeSum.apply(env_rgb, albedo, spec, nol, solid_angle)


class _RotToTarget(torch.autograd.Function):
    """`RotToTarget.rot2t` (utils/vector_rotation.py:57-89)."""

    @staticmethod
    @_amp_fwd
    def forward(ctx, tvec):
        r = tvec.shape[0]
        rot = torch.empty(r, 3, 3, device=tvec.device, dtype=torch.float32)
        with torch.cuda.device(tvec.device):
            check(_lib.lib().pnb_rot_to_target_fwd(r, _p(_req(tvec, "tvec")), _p(rot), _stream()), "rot_to_target_fwd")
        ctx.save_for_backward(tvec)
        return rot

    @staticmethod
    @_amp_bwd
    def backward(ctx, g_rot):
        (tvec,) = ctx.saved_tensors
        d_t = torch.empty_like(tvec)
        with torch.cuda.device(tvec.device):
            check(_lib.lib().pnb_rot_to_target_bwd(tvec.shape[0], _p(tvec), _p(g_rot.contiguous()), _p(d_t), _stream()),
                  "rot_to_target_bwd")
        return d_t


def rot_to_target(tvec):
    return _RotToTarget.apply(tvec)


class _EnvCastHemisp(torch.autograd.Function):
    """Env-ray Gaussians of `sample_each_points_hemisp` (models/mip.py:197-237): every surface point brings its own D
    directions.  Differentiable w.r.t. the surface points (the means are `point + dir * t_mean`); the directions are
    treated as data."""

    @staticmethod
    @_amp_fwd
    def forward(ctx, points, dirs, env_radii, env_near, env_far, n_env, t_rand):
        b, d = dirs.shape[0], dirs.shape[1]
        dev = points.device
        r = b * d
        t = torch.empty(r, n_env + 1, device=dev, dtype=torch.float32)
        means = torch.empty(r, n_env, 3, device=dev, dtype=torch.float32)
        covs = torch.empty(r, n_env, 3, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            check(_lib.lib().pnb_sample_cast_hemisp(r, n_env, _p(_req(points, "points")), d, _p(_req(dirs, "dirs")),
                                                    _p(_req(env_radii, "radii")), _p(_req(env_near, "near")),
                                                    _p(_req(env_far, "far")), d, _p(linspace01(n_env + 1, dev)),
                                                    _p(None if t_rand is None else _req(t_rand, "t_rand")), 0, _p(t),
                                                    _p(means), _p(covs), _stream()), "sample_cast_hemisp")
        ctx.k = d * n_env
        ctx.b = b
        ctx.mark_non_differentiable(t, covs)
        return t, means, covs

    @staticmethod
    @_amp_bwd
    def backward(ctx, g_t, g_means, g_covs):
        # d point = sum of the mean gradients of its D * Ne samples
        g = g_means.contiguous().view(ctx.b, ctx.k * 3)
        out = torch.empty(ctx.b * 3, device=g.device, dtype=torch.float32)
        with torch.cuda.device(g.device):
            check(_lib.lib().pnb_group_sum(ctx.b * ctx.k, 3, ctx.k, _p(g), 3, PNB_F32, _p(out), _stream()), "group_sum")
        return out.view(ctx.b, 3), None, None, None, None, None, None


def env_cast_hemisp(points, dirs, env_radii, env_near, env_far, n_env, t_rand=None):
    return _EnvCastHemisp.apply(points, dirs, env_radii, env_near, env_far, n_env, t_rand)


class _TonemapMSE(torch.autograd.Function):
    """sum(mask * (hdr_to_ldr(pred) - gt)^2) / sum(mask)   (systems/panonerf_system.py:44-50).
    `inv_mask_sum` is a python float or a 0-dim device tensor (1 / mask.sum() computed on the device)."""

    @staticmethod
    @_amp_fwd
    def forward(ctx, pred, gt_ldr, mask, inv_mask_sum):
        r = pred.shape[0]
        partial = torch.empty(r, device=pred.device, dtype=torch.float32)
        with torch.cuda.device(pred.device):
            check(_lib.lib().pnb_tonemap_se_fwd(r, _p(_req(pred, "pred")), _p(_req(gt_ldr, "gt")), _p(_req(mask, "mask")),
                                                _p(partial), _stream()), "tonemap_se_fwd")
        ctx.save_for_backward(pred, gt_ldr, mask)
        ctx.inv = inv_mask_sum
        if isinstance(inv_mask_sum, torch.Tensor):
            return dsum(partial) * inv_mask_sum
        return dsum(partial, inv_mask_sum)

    @staticmethod
    @_amp_bwd
    def backward(ctx, g):
        pred, gt_ldr, mask = ctx.saved_tensors
        d_pred = torch.empty_like(pred)
        gs = (g * ctx.inv).reshape(1).contiguous()
        with torch.cuda.device(pred.device):
            check(_lib.lib().pnb_tonemap_se_bwd(pred.shape[0], _p(pred), _p(gt_ldr), _p(mask), _p(gs), _p(d_pred),
                                                _stream()), "tonemap_se_bwd")
        return d_pred, None, None, None


def tonemap_mse(pred, gt_ldr, mask, inv_mask_sum):
    return _TonemapMSE.apply(pred, gt_ldr, mask, inv_mask_sum)


class _Chroma(torch.autograd.Function):
    """mean((normalize(gt) - normalize(albedo))^2)   (systems/panonerf_system.py:58-63)."""

    @staticmethod
    @_amp_fwd
    def forward(ctx, gt_ldr, albedo):
        r = albedo.shape[0]
        partial = torch.empty(r, device=albedo.device, dtype=torch.float32)
        with torch.cuda.device(albedo.device):
            check(_lib.lib().pnb_chroma_fwd(r, _p(_req(gt_ldr, "gt")), _p(_req(albedo, "albedo")), _p(partial),
                                            _stream()), "chroma_fwd")
        ctx.save_for_backward(gt_ldr, albedo)
        return dsum(partial, 1.0 / (3 * r))

    @staticmethod
    @_amp_bwd
    def backward(ctx, g):
        gt_ldr, albedo = ctx.saved_tensors
        r = albedo.shape[0]
        d_alb = torch.empty_like(albedo)
        gs = (g / (3 * r)).reshape(1).contiguous()
        with torch.cuda.device(albedo.device):
            check(_lib.lib().pnb_chroma_bwd(r, _p(gt_ldr), _p(albedo), _p(gs), _p(d_alb), _stream()), "chroma_bwd")
        return None, d_alb


def chroma_loss(gt_ldr, albedo):
    return _Chroma.apply(gt_ldr, albedo)


class _Mean(torch.autograd.Function):
    """Deterministic mean of a per-ray vector (ort_loss `.mean()`, models/pano_mip_nerf.py:311)."""

    @staticmethod
    @_amp_fwd
    def forward(ctx, x):
        ctx.n = x.numel()
        ctx.dev = x.device
        return dsum(x.contiguous(), 1.0 / x.numel())

    @staticmethod
    @_amp_bwd
    def backward(ctx, g):
        return (g / ctx.n).expand(ctx.n)


def dmean(x):
    return _Mean.apply(x)


# --------------------------------------------------------------------------------------------------------------
# render driver / validation outputs (csrc/image.cu)
# --------------------------------------------------------------------------------------------------------------
def pack_chw(per_ray, out, pix0: int):
    """Scatter per-ray results (list of [R, c_i] fp32 tensors) of rays [pix0, pix0+R) into `out` [sum(c_i), H*W]."""
    r = per_ray[0].shape[0]
    n = len(per_ray)
    chans = [int(x.shape[1]) if x.dim() == 2 else 1 for x in per_ray]
    if sum(chans) != out.shape[0]:
        raise RuntimeError("pack_chw: channel counts do not add up to the output planes")
    keep = [_req(x.reshape(r, c), "per-ray result") for x, c in zip(per_ray, chans)]
    src = (ctypes.c_void_p * n)(*[x.data_ptr() for x in keep])
    ch = (ctypes.c_int * n)(*chans)
    with torch.cuda.device(out.device):
        check(_lib.lib().pnb_pack_chw(r, out.shape[1], int(pix0), n, src, ch, _p(_req(out, "out")), _stream()),
              "pack_chw")
    return out


def image_sqerr_sum(pred, gt, row_weights=None):
    """sum((pred - gt)^2 [* row weight]) over a [C,H,W] image pair, fixed summation order -> 0-dim tensor."""
    pred, gt = _req(pred, "pred"), _req(gt, "gt")
    if pred.shape != gt.shape or pred.dim() != 3:
        raise RuntimeError("image metrics expect two [C,H,W] tensors of the same shape")
    c, h, w = pred.shape
    tmp = torch.empty(pred.numel(), device=pred.device, dtype=torch.float32)
    with torch.cuda.device(pred.device):
        check(_lib.lib().pnb_image_sqerr(pred.numel(), w, h * w, _p(pred), _p(gt),
                                         _p(None if row_weights is None else _req(row_weights, "row_weights")),
                                         _p(tmp), _stream()), "image_sqerr")
    return dsum(tmp)


def exr_payload(chw):
    """Scan-line blocks (uint8 device tensor) of an uncompressed float32 OpenEXR file for a [C,H,W] image."""
    chw = _req(chw, "image")
    c, h, w = chw.shape
    out = torch.empty(int(_lib.lib().pnb_exr_payload_bytes(h, w)), device=chw.device, dtype=torch.uint8)
    with torch.cuda.device(chw.device):
        check(_lib.lib().pnb_exr_pack(h, w, c, _p(chw), _p(out), _stream()), "exr_pack")
    return out


def png_payload(chw):
    """Filtered 8-bit RGB scan lines (uint8 device tensor) of a PNG for a [C,H,W] image in [0,1]."""
    chw = _req(chw, "image")
    c, h, w = chw.shape
    out = torch.empty(int(_lib.lib().pnb_png_payload_bytes(h, w)), device=chw.device, dtype=torch.uint8)
    with torch.cuda.device(chw.device):
        check(_lib.lib().pnb_png_pack(h, w, c, _p(chw), _p(out), _stream()), "png_pack")
    return out
