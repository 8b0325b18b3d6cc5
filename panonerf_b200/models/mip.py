"""CUDA-backed mirror of the free functions of the reference's models/mip.py (same names, argument meaning and
error behaviour).  Everything here launches kernels of libpanonerf_b200.so; tensors must be CUDA fp32."""
import torch

from .. import ops


def _check_shape(ray_shape):
    if ray_shape == "cone":
        return
    if ray_shape == "cylinder":
        raise NotImplementedError          # models/mip.py:83-84
    assert False                           # models/mip.py:86


def cast_rays(t_samples, origins, directions, radii, ray_shape, diagonal=True):
    """models/mip.py:67-89 (diagonal covariances only, as every live call site uses)."""
    _check_shape(ray_shape)
    if not diagonal:
        raise NotImplementedError("full covariances are only used by dead code upstream (sample_along_rays_360)")
    f = ops._f32c
    return ops.cast_rays_t(f(t_samples), f(origins), f(directions), f(radii))


def sample_along_rays(origins, directions, radii, num_samples, near, far, randomized, disparity, ray_shape,
                      t_rand=None):
    """models/mip.py:113-151.  `t_rand` optionally supplies the uniform draws (otherwise torch.rand on the ray
    device is consumed exactly like upstream, so a seeded run sees the same stream)."""
    _check_shape(ray_shape)
    f = ops._f32c
    if randomized and t_rand is None:
        t_rand = torch.rand(origins.shape[0], num_samples + 1, device=origins.device)
    t, means, covs = ops.sample_cast(f(origins), f(directions), f(radii), f(near), f(far), num_samples,
                                     t_rand=t_rand if randomized else None, disparity=disparity)
    return t, (means, covs)


def stratified_u(num_rays, num_samples, device):
    """The randomized `u` of models/mip.py:270-276 (same RNG consumption)."""
    s = 1 / num_samples
    eps = torch.finfo(torch.float32).eps
    u = (torch.arange(num_samples, device=device) * s)[None, ...]
    u = u + torch.empty([num_rays, num_samples], device=device).uniform_(to=(s - eps))
    return torch.minimum(u, torch.full_like(u, 1.0 - eps)).contiguous()


def sorted_piecewise_constant_pdf(bins, weights, num_samples, randomized, u=None, return_inds=False):
    """models/mip.py:240-301 on the given (already blurred/padded) weights; `num_samples` must be len(bins) as at the
    only upstream call site (mip.py:331-336)."""
    f = ops._f32c
    bins, weights = f(bins), f(weights)
    if num_samples != bins.shape[-1]:
        raise NotImplementedError("num_samples must equal the number of fence-posts")
    if randomized and u is None:
        u = stratified_u(weights.shape[0], num_samples, weights.device)
    return ops.resample(bins, weights, 0.0, u=u if randomized else None, return_inds=return_inds, blur_pool=False)


def resample_along_rays(origins, directions, radii, t_samples, weights, randomized, ray_shape, stop_grad,
                        resample_padding, u=None, return_inds=False):
    """models/mip.py:304-352.  `stop_grad=True` (configs/*.yaml `stop_resample_grad: True`) runs off the tape; with
    False the new fence-posts and Gaussians carry the gradient back to `weights` (ops.resample_cast_grad)."""
    _check_shape(ray_shape)
    f = ops._f32c
    if not stop_grad and weights.requires_grad and torch.is_grad_enabled():
        if return_inds:
            raise NotImplementedError("return_inds is a test hook of the stop_grad=True path")
        t_samples, weights = f(t_samples.detach()), f(weights)
        if randomized and u is None:
            u = stratified_u(weights.shape[0], t_samples.shape[-1], weights.device)
        new_t, means, covs = ops.resample_cast_grad(t_samples, weights, resample_padding, u if randomized else None,
                                                    f(origins), f(directions), f(radii))
        return new_t, (means, covs)
    t_samples, weights = f(t_samples.detach()), f(weights.detach())
    if randomized and u is None:
        u = stratified_u(weights.shape[0], t_samples.shape[-1], weights.device)
    res = ops.resample(t_samples, weights, resample_padding, u=u if randomized else None, return_inds=return_inds,
                       cast=(f(origins), f(directions), f(radii)))
    if return_inds:
        new_t, inds, means, covs = res
        return new_t, (means, covs), inds
    new_t, means, covs = res
    return new_t, (means, covs)


def integrated_pos_enc(means_covs, min_deg, max_deg, diagonal=True):
    """models/mip.py:394-428 (forward only; inside the model the encoding is written straight into the MLP's
    input buffer and differentiated by hand)."""
    if not diagonal:
        raise NotImplementedError
    means, covs = means_covs
    f = ops._f32c
    means, covs = f(means), f(covs)
    out = torch.empty(*means.shape[:-1], 6 * (max_deg - min_deg), device=means.device, dtype=torch.float32)
    ops.ipe_into(means.reshape(-1, 3), covs.reshape(-1, 3), min_deg, max_deg, out.view(-1, out.shape[-1]))
    return out


def pos_enc(x, min_deg, max_deg, append_identity=True):
    """models/mip.py:431-441 (min_deg must be 0 as at every call site)."""
    if min_deg != 0:
        raise NotImplementedError
    out = ops.pos_enc(ops._f32c(x), max_deg)
    return out if append_identity else out[:, 3:]


def volumetric_rendering(rgb, density, t_samples, dirs, white_bkgd, output_t=False):
    """models/mip.py:444-483.  density is [B,N,>=1]; channel 0 is used."""
    f = ops._f32c
    den = density[..., 0].contiguous() if density.dim() == 3 else density
    comp, dist, acc, w = ops.composite(rgb.contiguous(), den, f(t_samples), f(dirs), white_bkgd)
    if output_t:
        return comp, dist, acc, w, 0.5 * (t_samples[..., :-1] + t_samples[..., 1:])
    return comp, dist, acc, w


def volumetric_lighting_composing(rgb, density, t_samples, dirs, white_bkgd, output_t=False):
    """models/mip.py:486-527: volumetric_rendering with every sample's colour attenuated by 1 / (1 + t_mid^2)
    (the commented-out `enable_dist_att` branch of pano_mip_nerf.py:340-343)."""
    f = ops._f32c
    den = density[..., 0].contiguous() if density.dim() == 3 else density
    comp, dist, acc, w = ops.composite(rgb.contiguous(), den, f(t_samples), f(dirs), white_bkgd, attenuate=True)
    if output_t:
        return comp, dist, acc, w, 0.5 * (t_samples[..., :-1] + t_samples[..., 1:])
    return comp, dist, acc, w


def sample_each_points_hemisp(point_origins, directions, num_samples, near, far, radii, randomized, t_rand=None):
    """models/mip.py:197-237 with num_points == 1: `directions` is [batch, num_lit_rays, 3], one hemisphere of light
    directions per surface point (e.g. `RotToTarget.rot2t(normal) @ dirs`)."""
    b, npts, _ = point_origins.shape
    if npts != 1:
        raise NotImplementedError("num_points must be 1")
    f = ops._f32c
    if randomized and t_rand is None:
        t_rand = torch.rand(1, num_samples + 1, device=point_origins.device)
    dirs = f(directions)
    pts = point_origins.reshape(b, 3)
    pts = pts if pts.requires_grad else f(pts)
    t, means, covs = ops.env_cast_hemisp(pts.contiguous(), dirs, f(radii), f(near), f(far), num_samples,
                                         t_rand if randomized else None)
    return t, (means, covs), dirs.reshape(-1, 3)


def sample_each_points(point_origins, directions, num_samples, near, far, radii, randomized, t_rand=None):
    """models/mip.py:154-194 with num_points == 1 (the only shape the model passes, pano_mip_nerf.py:327)."""
    b, npts, _ = point_origins.shape
    if npts != 1:
        raise NotImplementedError("num_points must be 1")
    f = ops._f32c
    d = directions.shape[0]
    if randomized and t_rand is None:
        t_rand = torch.rand(1, num_samples + 1, device=point_origins.device)
    t, means, covs = ops.sample_cast(f(point_origins.reshape(b, 3)), f(directions), f(radii), f(near), f(far),
                                     num_samples, t_rand=t_rand if randomized else None, o_div=d, d_mod=d,
                                     n_rays=b * d, rand_shared=True)
    dirs = f(directions)[None].expand(b, d, 3).reshape(-1, 3)
    return t, (means, covs), dirs
