"""CPU oracle for the Pano-NeRF mip-NeRF volumetric-rendering hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``panonerf_b200/`` imports this file; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may.  It is a from-scratch restatement, in plain
PyTorch-on-CPU fp32, of the algorithm the reference implements in pure Python/PyTorch (there is no native
code upstream).  Every function cites the reference ``file:line`` it restates (paths relative to the upstream
repository root).  Parity status: PINNED — ``tests/test_oracle_vs_reference.py`` executes the unmodified
reference (imported from /root/reference in the build container) against these functions, and
``tests/golden/*.npz`` (made by ``tests/golden/make_golden.py`` from the reference) pin it wherever the reference
tree is absent (the GPU box).

Layout conventions are the reference's: rays are rows, samples are the second axis, fence-posts ``t`` have N+1
entries per ray, encodings are ``[sin(48) | cos(48)]`` with index ``3*l + c``.
"""
from __future__ import annotations

import collections
import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

# datasets/base_datasets.py:13-16
Rays = collections.namedtuple(
    "Rays", ("origins", "directions", "viewdirs", "radii", "lossmult", "near", "far", "noise_var"))

F32_EPS = float(torch.finfo(torch.float32).eps)


# ----------------------------------------------------------------------------------------------------------
# rays
# ----------------------------------------------------------------------------------------------------------
def equirect_rays(h: int, w: int, c2w, near: float, far: float) -> Rays:
    """datasets/pano_datasets.py:152-216 (one camera).  NumPy fp32 like the reference (numpy 1.24 semantics:
    the radius stays fp32; NumPy>=2 would promote it to fp64 - SURVEY.md App. B.19)."""
    c2w = np.asarray(c2w, dtype=np.float32)
    col, row = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32), indexing="xy")
    theta = -(col + 0.5) / w * 2 * np.pi                                   # :163
    phi = (row + 0.5) / h * np.pi                                          # :164
    cam = np.stack([np.sin(phi) * np.sin(theta), np.cos(phi), np.sin(phi) * np.cos(theta)], -1)  # :166-173
    noise = (np.sin(phi) * np.pi / w).reshape(h, w, 1)                     # :170-171
    dirs = (cam @ c2w[:3, :3].T).copy()                                    # :175
    orig = np.broadcast_to(c2w[:3, -1], dirs.shape).copy()                 # :176-179
    view = dirs / np.linalg.norm(dirs, axis=-1, keepdims=True)             # :182
    mid = dirs[h // 2]
    dx = np.sqrt(np.sum((mid[:-1] - mid[1:]) ** 2, -1))                    # :201
    dx = np.tile(dx[None, :], (h, 1))
    dx = np.concatenate([dx, dx[:, -2:-1]], 1)                             # :202
    radii = (dx[..., None] * 2 / np.float32(np.sqrt(12))).astype(np.float32)  # :203
    ones = np.ones_like(orig[..., :1])
    fields = (orig, dirs, view, radii, 1 * ones, near * ones, far * ones, noise)
    return Rays(*[torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).reshape(-1, x.shape[-1])
                  for x in fields])


def fibonacci_env_rays(num: int, radius: float, near: float = 0.0, far: float = 10.0,
                       dtype=torch.float16) -> Rays:
    """datasets/pano_datasets.py:218-263: golden-angle sphere directions, solid angle 4pi/num, cast to fp16."""
    ga = np.pi * (3.0 - np.sqrt(5.0))
    i = np.arange(num, dtype=np.float64)
    y = 1 - (i / float(num - 1)) * 2
    r = np.sqrt(1 - y * y)
    d = np.stack([np.cos(ga * i) * r, y, np.sin(ga * i) * r], -1)
    one = np.ones((num, 1))
    view = d / np.linalg.norm(d, axis=-1, keepdims=True)
    fields = (np.zeros_like(d), d, view, radius * one, (4 * np.pi / num) * one, near * one, far * one, 0 * one)
    return Rays(*[torch.tensor(x).to(dtype) for x in fields])


# ----------------------------------------------------------------------------------------------------------
# sampling / casting / encoding
# ----------------------------------------------------------------------------------------------------------
def cast_cone(t, origins, directions, radii):
    """models/mip.py:67-89 + :36-58 (stable branch) + :8-22 (diagonal lift)."""
    t0, t1 = t[..., :-1], t[..., 1:]
    mu, hw = (t0 + t1) / 2, (t1 - t0) / 2
    den = 3 * mu ** 2 + hw ** 2
    t_mean = mu + (2 * mu * hw ** 2) / den
    t_var = (hw ** 2) / 3 - (4 / 15) * ((hw ** 4 * (12 * mu ** 2 - hw ** 2)) / den ** 2)
    r_var = radii ** 2 * ((mu ** 2) / 4 + (5 / 12) * hw ** 2 - 4 / 15 * (hw ** 4) / den)
    mean = directions[..., None, :] * t_mean[..., None]
    dd = directions ** 2
    null = 1 - dd / (torch.sum(dd, -1, keepdim=True) + 1e-10)
    cov = t_var[..., None] * dd[..., None, :] + r_var[..., None] * null[..., None, :]
    return mean + origins[..., None, :], cov


def stratified_t(near, far, n, randomized: bool, disparity: bool = False, t_rand=None):
    """models/mip.py:131-149.  `t_rand` (if given) replaces the internal torch.rand draw."""
    b = near.shape[0]
    s = torch.linspace(0.0, 1.0, n + 1)
    t = 1.0 / (1.0 / near * (1.0 - s) + 1.0 / far * s) if disparity else near + (far - near) * s
    if randomized:
        mids = 0.5 * (t[..., 1:] + t[..., :-1])
        upper = torch.cat([mids, t[..., -1:]], -1)
        lower = torch.cat([t[..., :1], mids], -1)
        if t_rand is None:
            t_rand = torch.rand(b, n + 1)
        t = lower + (upper - lower) * t_rand
    else:
        t = torch.broadcast_to(t, (b, n + 1))
    return t


def sample_along_rays(origins, directions, radii, n, near, far, randomized, disparity=False, t_rand=None):
    """models/mip.py:113-151."""
    t = stratified_t(near, far, n, randomized, disparity, t_rand)
    return t, cast_cone(t, origins, directions, radii)


def ipe(mean, cov, min_deg: int, max_deg: int):
    """models/mip.py:394-428 (diagonal) with expected_sin :355-361 (first return value only)."""
    scales = torch.tensor([2.0 ** i for i in range(min_deg, max_deg)])
    y = (mean[..., None, :] * scales[:, None]).flatten(-2)
    yv = (cov[..., None, :] * scales[:, None] ** 2).flatten(-2)
    half_pi = 0.5 * torch.tensor(np.pi)          # fp64 0-dim tensor; the sum below stays fp32 like upstream
    arg = torch.cat([y, y + half_pi], -1)
    return torch.exp(-0.5 * torch.cat([yv, yv], -1)) * torch.sin(arg)


def pos_enc(x, min_deg: int, max_deg: int, append_identity=True):
    """models/mip.py:431-441."""
    scales = torch.tensor([2.0 ** i for i in range(min_deg, max_deg)])
    xb = (x[..., None, :] * scales[:, None]).flatten(-2)
    feat = torch.sin(torch.cat([xb, xb + 0.5 * torch.tensor(np.pi)], -1))
    return torch.cat([x, feat], -1) if append_identity else feat


# ----------------------------------------------------------------------------------------------------------
# MLP (state-dict driven so the same weights feed the oracle, the reference and the CUDA path)
# ----------------------------------------------------------------------------------------------------------
def mlp_forward(sd: Dict[str, torch.Tensor], enc, venc, skip_index: int = 4):
    """models/pano_mip_nerf.py:78-114 == models/mip_nerf.py:70-102.  `sd` uses the reference's parameter
    names (layers.{i}.0.weight ..., density_layer, extra_layer, view_layers.0.0, color_layer)."""
    depth = len([k for k in sd if k.startswith("layers.") and k.endswith(".weight")])
    x = enc
    for i in range(depth):
        x = torch.relu(F.linear(x, sd[f"layers.{i}.0.weight"], sd[f"layers.{i}.0.bias"]))
        if i % skip_index == 0 and i > 0:
            x = torch.cat([x, enc], -1)
    raw_density = F.linear(x, sd["density_layer.weight"], sd["density_layer.bias"])
    if venc is not None:
        bott = F.linear(x, sd["extra_layer.weight"], sd["extra_layer.bias"])
        v = venc[:, None, :].expand(-1, enc.shape[1], -1)
        x = torch.cat([bott, v], -1)
        j = 0
        while f"view_layers.{j}.0.weight" in sd:
            x = torch.relu(F.linear(x, sd[f"view_layers.{j}.0.weight"], sd[f"view_layers.{j}.0.bias"]))
            j += 1
    raw_rgb = F.linear(x, sd["color_layer.weight"], sd["color_layer.bias"])
    return raw_rgb, raw_density


def softplus(x):
    """torch.nn.Softplus() defaults: beta=1, threshold=20."""
    return F.softplus(x)


def radiance_field(sd, mean, cov, viewdirs, cfg):
    """compute_graph closure: models/pano_mip_nerf.py:235-280 / models/mip_nerf.py:206-243."""
    if cfg.get("disable_integration", False):
        cov = torch.zeros_like(cov)
    enc = ipe(mean, cov, cfg["min_deg_point"], cfg["max_deg_point"])
    venc = pos_enc(viewdirs, 0, cfg["deg_view"], True) if cfg.get("use_viewdirs", True) else None
    raw_rgb, raw_den = mlp_forward(sd, enc, venc, cfg.get("skip_index", 4))
    pad = cfg["rgb_padding"]
    rgb = softplus(raw_rgb) * (1 + 2 * pad) - pad
    density = softplus(raw_den[..., :1] + cfg["density_bias"])
    out = {"rgb": rgb, "density": density}
    if raw_den.shape[-1] >= 5:                       # panonerf split, pano_mip_nerf.py:264-278
        out["albedo"] = torch.sigmoid(raw_den[..., 1:-1]) * 0.77 + 0.03
        out["roughness"] = softplus(raw_den[..., -1:] - 1)
    return out


# ----------------------------------------------------------------------------------------------------------
# compositing / resampling
# ----------------------------------------------------------------------------------------------------------
def composite(rgb, density, t, dirs, white_bkgd: bool):
    """models/mip.py:444-483."""
    t_mid = 0.5 * (t[..., :-1] + t[..., 1:])
    delta = (t[..., 1:] - t[..., :-1]) * torch.linalg.norm(dirs[..., None, :], dim=-1)
    sd_ = density[..., 0] * delta
    alpha = 1 - torch.exp(-sd_)
    trans = torch.exp(-torch.cat([torch.zeros_like(sd_[..., :1]), torch.cumsum(sd_[..., :-1], -1)], -1))
    w = alpha * trans
    comp = (w[..., None] * rgb).sum(-2)
    acc = w.sum(-1)
    dist = (w * t_mid).sum(-1) / acc
    dist = torch.clamp(torch.nan_to_num(dist), t[:, 0], t[:, -1])
    if white_bkgd:
        comp = comp + (1.0 - acc[..., None])
    return comp, dist, acc, w


def blur_weights(w, padding: float):
    """models/mip.py:324-329."""
    wp = torch.cat([w[..., :1], w, w[..., -1:]], -1)
    wm = torch.maximum(wp[..., :-1], wp[..., 1:])
    return 0.5 * (wm[..., :-1] + wm[..., 1:]) + padding


def pdf_sample(bins, weights, num_samples: int, randomized: bool, u=None, return_aux=False):
    """models/mip.py:240-301.  `u` (if given) replaces the internal draw (deterministic: linspace(0,1-eps))."""
    eps = 1e-5
    wsum = torch.sum(weights, -1, keepdim=True)
    padding = torch.maximum(torch.zeros_like(wsum), eps - wsum)
    weights = weights + padding / weights.shape[-1]
    wsum = wsum + padding
    pdf = weights / wsum
    cdf = torch.cumsum(pdf[..., :-1], -1)
    cdf = torch.minimum(torch.ones_like(cdf), cdf)      # (minimum, not clamp: its tie rule is what autograd sees)
    z = torch.zeros(list(cdf.shape[:-1]) + [1])
    cdf = torch.cat([z, cdf, z + 1.0], -1)
    if u is None:
        if randomized:
            s = 1 / num_samples
            u = (torch.arange(num_samples) * s)[None, :]
            u = u + torch.empty(list(cdf.shape[:-1]) + [num_samples]).uniform_(to=(s - F32_EPS))
            u = torch.clamp_max(u, 1.0 - F32_EPS)
        else:
            u = torch.linspace(0.0, 1.0 - F32_EPS, num_samples)
            u = torch.broadcast_to(u, list(cdf.shape[:-1]) + [num_samples])
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    lo = torch.clamp_min(inds - 1, 0)
    hi = torch.clamp_max(inds, cdf.shape[-1] - 1)
    c0, c1 = torch.gather(cdf, -1, lo), torch.gather(cdf, -1, hi)
    b0, b1 = torch.gather(bins, -1, lo), torch.gather(bins, -1, hi)
    den = c1 - c0
    den = torch.where(den < 1e-5, torch.ones_like(den), den)
    out = b0 + ((u - c0) / den) * (b1 - b0)
    return (out, inds, cdf) if return_aux else out


def resample_along_rays(origins, directions, radii, t, weights, randomized, padding, u=None, stop_grad=True):
    """models/mip.py:304-352.  stop_grad=True is what the configs use (*.yaml `stop_resample_grad`); with False the
    new fence-posts stay on the tape (the else branch, mip.py:336-350): the fine level's loss reaches the coarse weights
    through the CDF inversion."""
    if stop_grad:
        with torch.no_grad():
            new_t = pdf_sample(t, blur_weights(weights, padding), t.shape[-1], randomized, u)
    else:
        new_t = pdf_sample(t, blur_weights(weights, padding), t.shape[-1], randomized, u)
    return new_t, cast_cone(new_t, origins, directions, radii)


# ----------------------------------------------------------------------------------------------------------
# validation outputs: metrics and image quantisation
# ----------------------------------------------------------------------------------------------------------
def solid_angle_refinement(h=8, w=16):
    """utils/surface_rendering.py:294-316 (spherical, torch branch): [1, h*w, 1] fp32."""
    d_phi, d_theta = np.pi / h, 2 * np.pi / w
    y = (np.arange(h) + 0.5) / h
    x = (np.arange(w) + 0.5) / w
    _, yy = np.meshgrid(x, y)
    return torch.Tensor((np.sin(yy * np.pi) * d_theta * d_phi).reshape(1, -1, 1))


def calc_psnr(x, y):
    """utils/metrics.py:210-214, 231-237."""
    return -10.0 * torch.log10(torch.mean((x - y) ** 2))


def calc_ws_psnr(pred, gt):
    """utils/metrics.py:318-326."""
    c, h, w = pred.shape
    weights = solid_angle_refinement(h=h, w=w).reshape(1, h, w)
    weights = weights / weights.sum()
    return -10.0 * torch.log10(torch.sum((pred - gt) ** 2 * weights))


def png_pixels(image_1chw):
    """utils/vis.py:29-35: the uint8 [H,W,3] array save_results hands to PIL (single channel replicated)."""
    img = image_1chw[0].permute(1, 2, 0).cpu().numpy()
    if img.shape[-1] == 1:
        img = np.concatenate([img] * 3, axis=-1)
    return (img * 255).astype(np.uint8)


def ipe_exact(mean, cov, min_deg: int, max_deg: int):
    """float64 value of models/mip.py:394-428 on the fp32 ARGUMENTS upstream builds (scaled means, `y + pi/2` rounded
    to fp32, fp32 exp argument): what a correctly rounded fp32 sin / exp would return, independent of the host libm."""
    scales = torch.tensor([2.0 ** i for i in range(min_deg, max_deg)])
    y = (mean[..., None, :] * scales[:, None]).flatten(-2)
    yv = (cov[..., None, :] * scales[:, None] ** 2).flatten(-2)
    arg = torch.cat([y, y + 0.5 * torch.tensor(np.pi)], -1)
    ex = -0.5 * torch.cat([yv, yv], -1)
    return torch.exp(ex.double()) * torch.sin(arg.double())


def resample_case(n: int, rays: int = 4096):
    """Seeded config-size inputs of the resampling tests / golden vectors (tests/golden/make_golden.py): peaky
    weights incl. an all-zero ray and a single spike, sorted fence-posts, random ray geometry."""
    gen = torch.Generator().manual_seed(1000 + n)
    w = torch.rand(rays, n, generator=gen) ** 4
    w[0] = 0.0
    w[1] = 0.0
    w[1, n // 2] = 1.0
    w[2] = 1e-9                                   # weight_sum below eps: the padding branch of mip.py:253-257
    t = torch.sort(torch.rand(rays, n + 1, generator=gen) * 10, dim=-1).values
    o = torch.rand(rays, 3, generator=gen) - 0.5
    d = torch.nn.functional.normalize(torch.randn(rays, 3, generator=gen), dim=-1)
    rad = torch.full((rays, 1), 0.0035)
    return t, w, o, d, rad


# ----------------------------------------------------------------------------------------------------------
# surface branch / tone mapping
# ----------------------------------------------------------------------------------------------------------
def env_samples(points, env: Rays, n_env: int, randomized: bool, t_rand=None):
    """models/mip.py:154-194 with num_points == 1 (pano_mip_nerf.py:327).  Returns t[B*D,Ne+1], (mean,cov), dirs."""
    b, d = points.shape[0], env.directions.shape[0]
    o = points[:, None, :].expand(b, d, 3).reshape(-1, 3)
    rep = lambda x: x[None].expand(b, *x.shape).reshape(-1, x.shape[-1])
    dirs, radii, near, far = rep(env.directions), rep(env.radii), rep(env.near), rep(env.far)
    t = near + (far - near) * torch.linspace(0.0, 1.0, n_env + 1)
    if randomized:
        mids = 0.5 * (t[..., 1:] + t[..., :-1])
        upper = torch.cat([mids, t[..., -1:]], -1)
        lower = torch.cat([t[..., :1], mids], -1)
        if t_rand is None:
            t_rand = torch.rand(1, n_env + 1)        # one draw shared by every env ray (mip.py:186)
        t = lower + (upper - lower) * t_rand
    return t, cast_cone(t, o, dirs, radii), dirs


def lambert_shade(env_rgb, albedo, normal, light_dirs, solid_angle):
    """utils/surface_rendering.py:129-165 (roughness=None) + :104-126."""
    nol = torch.relu((normal[:, None, :] * light_dirs).sum(-1, keepdim=True))
    shading = torch.sum(env_rgb * nol * solid_angle, dim=1)
    diffuse = albedo / np.pi * shading
    return diffuse + torch.zeros_like(diffuse), diffuse, shading


# ---- variants the upstream hot path does not call today ---------------------------
def _unit_dots(normal, l, v):
    """Half vector and the clamped cosines shared by the two specular BRDFs (utils/surface_rendering.py:28-44)."""
    d = l.shape[1]
    vv = v[:, None, :].expand(-1, d, -1)
    nn = normal[:, None, :].expand(-1, d, -1)
    h = F.normalize(l + vv, dim=-1)
    dot = lambda a, b: (a * b).sum(-1, keepdim=True)
    return dot(nn, h), dot(vv, h), dot(nn, l), dot(nn, vv)


def microfacet_terms(albedo, normal, roughness, l, v, masked: bool = False):
    """utils/surface_rendering.py:6-61 (`microfeast_brdf`, UE4 GGX / Schlick / Smith for image-based lighting).
    Returns (diffuse_brdf [B,D,3], specular_brdf [B,D,1], NoL [B,D,1]).  `masked=True` evaluates the same specular
    term with the 0/0 entries masked BEFORE the division: identical forward values, but a finite gradient (upstream's
    `nan_to_num` after the division back-propagates NaN into every ray that has a light below its horizon)."""
    d = l.shape[1]
    noh, voh, nol, nov = [torch.relu(x) for x in _unit_dots(normal, l, v)]
    r = roughness[:, None, :].expand(-1, d, -1)
    alpha = r ** 2
    k = r ** 2 / 2
    dist = alpha ** 2 / (np.pi * ((noh ** 2) * (alpha ** 2 - 1) + 1) ** 2)
    fres = 0.04 + (1 - 0.04) * 2 ** (-(5.55473 * voh + 6.98316) * voh)
    geom = nol / ((1 - k) * nol + k) * (nov / ((1 - k) * nov + k))
    den = 4 * nol * nov
    if masked:
        ok = den > 0
        spec = torch.where(ok, dist * fres * geom / torch.where(ok, den, torch.ones_like(den)), torch.zeros_like(den))
    else:
        spec = (dist * fres * geom / den).nan_to_num(nan=0, posinf=0)
    return (albedo / np.pi)[:, None, :].expand(-1, d, -1), spec, nol


def blinn_phong_terms(albedo, normal, roughness, l, v):
    """utils/surface_rendering.py:64-101 (`blinn_phong_brdf`): specular = relu(n.h) ** roughness, NoL NOT clamped."""
    d = l.shape[1]
    noh, _, nol, _ = _unit_dots(normal, l, v)
    spec = torch.pow(torch.relu(noh), roughness[:, None, :].expand(-1, d, -1)).nan_to_num(nan=0, posinf=0)
    return (albedo / np.pi)[:, None, :].expand(-1, d, -1), spec, nol


def rough_shade(env_rgb, albedo, normal, roughness, l, v, solid_angle, masked: bool = False):
    """utils/surface_rendering.py:147-151,159 (the `roughness is not None` branch of `surface_rendering`)."""
    dif_b, spec_b, nol = microfacet_terms(albedo, normal, roughness, l, v, masked)
    diffuse = torch.sum(dif_b * env_rgb * nol * solid_angle, dim=1)
    specular = torch.sum(spec_b * env_rgb * solid_angle, dim=1)
    return diffuse + specular, diffuse, specular


def rot_to_target(tvec):
    """utils/vector_rotation.py:57-89 (`RotToTarget.rot2t`): Rodrigues rotation taking (0,1,0) onto each row of
    tvec [B,3] -> [B,3,3]; the antipodal case (theta == pi) is the fixed reflection diag(1,-1,1)."""
    up = torch.tensor([0.0, 1.0, 0.0], dtype=tvec.dtype)
    theta = torch.acos((tvec * up).sum(-1)).view(-1, 1, 1)
    axis = F.normalize(torch.cross(up[None].expand_as(tvec), tvec, dim=-1), dim=-1)
    z = torch.zeros_like(axis[:, 0])
    skew = torch.stack([z, -axis[:, 2], axis[:, 1], axis[:, 2], z, -axis[:, 0], -axis[:, 1], axis[:, 0], z], -1)
    skew = skew.view(-1, 3, 3)
    rm = torch.eye(3, dtype=tvec.dtype)[None] + torch.sin(theta) * skew + torch.bmm(skew, skew) * (1 - torch.cos(theta))
    flip = (theta.view(-1) == np.pi)
    return torch.where(flip[:, None, None], torch.diag(torch.tensor([1.0, -1.0, 1.0], dtype=tvec.dtype))[None], rm)


def env_samples_hemisp(points, dirs, env: Rays, n_env: int, randomized: bool, t_rand=None):
    """models/mip.py:197-237 (`sample_each_points_hemisp`, num_points == 1): like env_samples, but every surface
    point brings its own D directions, dirs [B,D,3]."""
    b, d = dirs.shape[0], dirs.shape[1]
    o = points[:, None, :].expand(b, d, 3).reshape(-1, 3)
    rep = lambda x: x[None].expand(b, *x.shape).reshape(-1, x.shape[-1])
    radii, near, far = rep(env.radii), rep(env.near), rep(env.far)
    t = near + (far - near) * torch.linspace(0.0, 1.0, n_env + 1)
    if randomized:
        mids = 0.5 * (t[..., 1:] + t[..., :-1])
        upper = torch.cat([mids, t[..., -1:]], -1)
        lower = torch.cat([t[..., :1], mids], -1)
        if t_rand is None:
            t_rand = torch.rand(1, n_env + 1)
        t = lower + (upper - lower) * t_rand
    flat = dirs.reshape(-1, 3)
    return t, cast_cone(t, o, flat, radii), flat


def composite_lighting(rgb, density, t, dirs, white_bkgd: bool):
    """models/mip.py:486-527 (`volumetric_lighting_composing`): alpha compositing with the colour of every sample
    attenuated by 1 / (1 + t_mid^2); weights, acc and distance are those of `composite`."""
    _, dist, acc, w = composite(rgb, density, t, dirs, False)
    t_mid = 0.5 * (t[..., :-1] + t[..., 1:])
    comp = (w[..., None] * (1 / (1 + t_mid ** 2))[..., None] * rgb).sum(-2)
    if white_bkgd:
        comp = comp + (1.0 - acc[..., None])
    return comp, dist, acc, w


def hdr_to_ldr(c, gamma=2.2, quantize=False, clamp=True):
    """utils/surface_rendering.py:319-344 (ACES + gamma; `quantize` == dtype='uint8')."""
    c = (c * (2.51 * c + 0.03)) / (c * (2.43 * c + 0.59) + 0.14)
    if clamp:
        c = torch.clamp(c, 0, 1)
    if quantize:
        c = ((c * 255.0).to(torch.uint8) / 255.0).to(torch.float32)
    return c ** (1 / gamma)


# ----------------------------------------------------------------------------------------------------------
# models
# ----------------------------------------------------------------------------------------------------------
DEFAULT_CFG = dict(num_samples=64, num_levels=2, resample_padding=0.01, min_deg_point=0, max_deg_point=16,
                   deg_view=4, density_bias=-1.0, rgb_padding=0.0, skip_index=4, use_viewdirs=True,
                   disparity=False, num_env_samples=10, disable_integration=False)


def _density_normals(sd, mean, cov, viewdirs, cfg, create_graph):
    """n = -d(density)/d(mean), per sample.

    cfg["normals_impl"] == "autograd" (default): one reverse pass, autograd.grad(density.sum(), mean) - equal to the
    reference's result (max abs diff 6e-7) at a fraction of its cost.
    cfg["normals_impl"] == "jacrev": literally what the reference executes, vmap(jacrev(compute_graph, argnums=0))
    over every sample with all of compute_graph's outputs differentiated and only the density row kept
    (models/pano_mip_nerf.py:299-302, models/mip_nerf.py:261-264).  bench.py times THIS variant as the CPU arm."""
    if cfg.get("normals_impl", "autograd") == "jacrev":
        from torch.func import jacrev, vmap

        def compute_graph(m1, v1, d1):
            f = radiance_field(sd, m1.view(1, 1, -1), v1.view(1, 1, -1), d1.view(1, -1), cfg)
            keys = ("rgb", "density", "albedo", "roughness") if "albedo" in f else ("rgb", "density")
            return tuple(f[k] for k in keys)

        n = mean.shape[1]
        vd = viewdirs.view(-1, 1, 3).repeat(1, n, 1).view(-1, 3)
        jac = vmap(jacrev(compute_graph, argnums=0))(mean.reshape(-1, 3), cov.reshape(-1, 3), vd)[1]
        return -jac.view(mean.shape[0], n, 3)
    m = mean.detach().requires_grad_(True) if not mean.requires_grad else mean
    with torch.enable_grad():
        vd = viewdirs
        den = radiance_field(sd, m, cov, vd, cfg)["density"]
        (g,) = torch.autograd.grad(den.sum(), m, create_graph=create_graph)
    return -g


def _fine_level_normals(sd, mean, cov, rays, weights, norm_den, cfg, use_ort, train):
    nrm = _density_normals(sd, mean, cov, rays.viewdirs, cfg, create_graph=train)
    nrm = F.normalize(nrm, dim=-1)
    nw = weights[..., None] / norm_den.view(-1, 1, 1)
    normal = F.normalize(torch.sum(nw * nrm, dim=1), dim=-1)
    ort = None
    if use_ort:
        dot = torch.bmm(nrm, rays.directions.view(-1, 3, 1))
        ort = torch.sum(nw * torch.relu(dot) ** 2, dim=1).mean()
    return nrm, nw, normal, ort


def mipnerf_forward(sd, rays: Rays, cfg, randomized=False, white_bkgd=False, use_ort_loss=False, train=False,
                    rand=None):
    """models/mip_nerf.py:170-283.  Returns [(comp_rgb, distance, ort_loss, den_normal)] per level and an aux
    dict with t / weights per level (test visibility only)."""
    cfg = {**DEFAULT_CFG, **cfg}
    rand = rand or {}
    ret, aux, t, w = [], [], None, None
    for lvl in range(cfg["num_levels"]):
        if lvl == 0:
            t, (mean, cov) = sample_along_rays(rays.origins, rays.directions, rays.radii, cfg["num_samples"],
                                               rays.near, rays.far, randomized, cfg["disparity"],
                                               rand.get("t_rand"))
        else:
            t, (mean, cov) = resample_along_rays(rays.origins, rays.directions, rays.radii, t, w, randomized,
                                                 cfg["resample_padding"], rand.get("u"),
                                                 cfg.get("stop_resample_grad", True))
        f = radiance_field(sd, mean, cov, rays.viewdirs, cfg)
        comp, dist, acc, w = composite(f["rgb"], f["density"], t, rays.directions, white_bkgd)
        if lvl == 1 and use_ort_loss:
            _, _, normal, ort = _fine_level_normals(sd, mean, cov, rays, w, acc, cfg, True, train)
            ret.append((comp, dist, ort, normal))
        else:
            ret.append((comp, dist, None, torch.ones_like(comp)))
        aux.append(dict(t=t, weights=w, acc=acc))
    return ret, aux


def panonerf_forward(sd, rays: Rays, env: Rays, cfg, randomized=False, white_bkgd=False, enable_surf=True,
                     use_ort_loss=True, train=False, rand=None):
    """models/pano_mip_nerf.py:197-363.  9-tuples per level (coarse level: None past `distance`)."""
    cfg = {**DEFAULT_CFG, **cfg}
    rand = rand or {}
    ret, aux, t, w = [], [], None, None
    for lvl in range(cfg["num_levels"]):
        if lvl == 0:
            t, (mean, cov) = sample_along_rays(rays.origins, rays.directions, rays.radii, cfg["num_samples"],
                                               rays.near, rays.far, randomized, cfg["disparity"],
                                               rand.get("t_rand"))
        else:
            t, (mean, cov) = resample_along_rays(rays.origins, rays.directions, rays.radii, t, w, randomized,
                                                 cfg["resample_padding"], rand.get("u"),
                                                 cfg.get("stop_resample_grad", True))
        f = radiance_field(sd, mean, cov, rays.viewdirs, cfg)
        comp, dist, acc, w = composite(f["rgb"], f["density"], t, rays.directions, white_bkgd)
        normal = surf = albedo = diffuse = ort = shading = None
        if lvl == 1:
            nrm, nw, normal, ort = _fine_level_normals(sd, mean, cov, rays, w, torch.sum(w, -1), cfg,
                                                      use_ort_loss, train)
            if enable_surf:
                albedo = torch.sum(nw * f["albedo"], dim=1)
                pts = rays.origins + rays.directions * dist.view(-1, 1)
                lt, (lm, lc), ldir = env_samples(pts, env, cfg["num_env_samples"], randomized,
                                                 rand.get("env_t_rand"))
                g = radiance_field(sd, lm, lc, ldir, cfg)
                env_rgb = composite(g["rgb"], g["density"], lt, ldir, False)[0].view(normal.shape[0], -1, 3)
                surf, diffuse, shading = lambert_shade(env_rgb, albedo, normal, ldir.view(env_rgb.shape),
                                                       env.lossmult)
        ret.append((comp, dist, ort, normal, albedo, None, surf, diffuse, shading))
        aux.append(dict(t=t, weights=w, acc=acc))
    return ret, aux


# ----------------------------------------------------------------------------------------------------------
# losses (systems/*_system.py training_step bodies)
# ----------------------------------------------------------------------------------------------------------
def mipnerf_loss(outputs, rays, gt_hdr, coarse_mult=0.1, ort_mult=0.0):
    """systems/mipnerf_system.py:22-53."""
    gt = hdr_to_ldr(gt_hdr[..., :3], quantize=True)
    (c, *_), (f, _, ort, _) = outputs
    m = rays.lossmult
    lc = (m * (hdr_to_ldr(c) - gt) ** 2).sum() / m.sum()
    lf = (m * (hdr_to_ldr(f) - gt) ** 2).sum() / m.sum()
    loss = coarse_mult * lc + lf
    if ort_mult > 0:
        loss = loss + ort_mult * ort
    return loss


def panonerf_loss(outputs, rays, gt_hdr, surface_on=True, coarse_mult=0.1, surface_mult=1.0, ort_mult=0.1,
                  chrom_mult=0.1):
    """systems/panonerf_system.py:15-75."""
    gt = hdr_to_ldr(gt_hdr[..., :3], quantize=True)
    (c, *_), (f, _, ort, _, alb, _, sf, _, _) = outputs
    m = rays.lossmult
    lc = (m * (hdr_to_ldr(c) - gt) ** 2).sum() / m.sum()
    lf = (m * (hdr_to_ldr(f) - gt) ** 2).sum() / m.sum()
    loss = coarse_mult * lc + lf
    if surface_on:
        ls = (m * (hdr_to_ldr(sf) - gt) ** 2).sum() / m.sum()
        loss = loss + surface_mult * ls
        if chrom_mult > 0:
            loss = loss + chrom_mult * ((F.normalize(gt, dim=-1) - F.normalize(alb, dim=-1)) ** 2).mean()
    if ort is not None:
        loss = loss + ort_mult * ort
    return loss


# ----------------------------------------------------------------------------------------------------------
# optimiser (systems/base_system.py:81-87, utils/lr_schedule.py:51-60)
# ----------------------------------------------------------------------------------------------------------
def mip_lr(step: int, lr_init=2e-4, lr_final=2e-5, max_steps=44000, delay_steps=120, delay_mult=0.01) -> float:
    rate = 1.0
    if delay_steps > 0:
        rate = delay_mult + (1 - delay_mult) * math.sin(0.5 * math.pi * min(max(step / delay_steps, 0), 1))
    tt = min(max(step / max_steps, 0), 1)
    return rate * math.exp(math.log(lr_init) * (1 - tt) + math.log(lr_final) * tt)


def adam_step(p, g, m, v, step: int, lr: float, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam defaults (base_system.py:82), one tensor, in place; `step` is 1-based."""
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    p.addcdiv_(m, (v.sqrt() / math.sqrt(bc2)).add_(eps), value=-lr / bc1)


# ----------------------------------------------------------------------------------------------------------
# deterministic parameter factory shared by tests / bench (no reference needed)
# ----------------------------------------------------------------------------------------------------------
def mlp_shapes(width=256, depth=8, width_cond=128, skip=4, c_density=1, xyz=96, view=27, rgb=3):
    """Parameter names/shapes of models/pano_mip_nerf.py:38-76 in state-dict order."""
    shp = collections.OrderedDict()
    for i in range(depth):
        din = xyz if i == 0 else (width + xyz if (i - 1) % skip == 0 and i > 1 else width)
        shp[f"layers.{i}.0.weight"] = (width, din)
        shp[f"layers.{i}.0.bias"] = (width,)
    shp["density_layer.weight"], shp["density_layer.bias"] = (c_density, width), (c_density,)
    shp["extra_layer.weight"], shp["extra_layer.bias"] = (width, width), (width,)
    shp["view_layers.0.0.weight"], shp["view_layers.0.0.bias"] = (width_cond, width + view), (width_cond,)
    shp["color_layer.weight"], shp["color_layer.bias"] = (rgb, width_cond), (rgb,)
    return shp


def synth_state_dict(seed=4, gain=1.0, **kw) -> Dict[str, torch.Tensor]:
    """Xavier-uniform-like weights from a CPU generator: identical on every machine for a given torch build."""
    g = torch.Generator().manual_seed(seed)
    sd = collections.OrderedDict()
    for k, s in mlp_shapes(**kw).items():
        if k.endswith("weight"):
            bound = gain * math.sqrt(6.0 / (s[0] + s[1]))
        else:
            bound = 0.05
        sd[k] = (torch.rand(s, generator=g) * 2 - 1) * bound
    return sd
