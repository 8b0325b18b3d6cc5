"""CUDA replacements for the live functions of utils/surface_rendering.py."""
import math

import torch

from .. import ops


def hdr_to_ldr(color, gamma=2.2, dtype="float32", clamp=True):
    """ACES tone-mapping + gamma (utils/surface_rendering.py:319-344).  Forward only (ground truth / image dumps);
    the differentiable use inside the losses is fused in ops.tonemap_mse."""
    if gamma != 2.2 or not clamp:
        raise NotImplementedError("only the reference defaults gamma=2.2, clamp=True are implemented")
    if not isinstance(color, torch.Tensor):
        raise TypeError("color must be a CUDA torch.Tensor")
    return ops.hdr_to_ldr(ops._f32c(color), quantize=(dtype == "uint8"))


def _specular_brdf(kind, albedo, normal, roughness, l, v):
    f = ops._f32c
    d = l.shape[1]
    rough = roughness.reshape(-1)
    spec, nol = ops.brdf_terms(kind, normal.contiguous(), rough.contiguous(), f(l), f(v))
    diffuse_brdf = (albedo / math.pi)[:, None, :].expand(-1, d, -1)
    return diffuse_brdf, spec[..., None], nol[..., None]


def microfeast_brdf(albedo, normal, roughness, l, v):
    """utils/surface_rendering.py:6-61 (UE4 microfacet BRDF for image-based lighting): albedo [B,3], normal [B,3],
    roughness [B,1], l [B,D,3], v [B,3] -> diffuse_brdf [B,D,3], specular_brdf [B,D,1], NoL [B,D,1] (clamped).
    Differentiable w.r.t. albedo, normal, roughness; entries upstream zeroes with `nan_to_num` (lights at or below the
    horizon) are 0 with zero gradient (upstream back-propagates NaN through them)."""
    return _specular_brdf(ops.BRDF_MICROFACET, albedo, normal, roughness, l, v)


def blinn_phong_brdf(albedo, normal, roughness, l, v):
    """utils/surface_rendering.py:64-101: specular = relu(n.h) ** roughness, NoL returned un-clamped as upstream."""
    return _specular_brdf(ops.BRDF_BLINN_PHONG, albedo, normal, roughness, l, v)


def surface_rendering(env, albedo, normal, roughness, l, v, solid_angle, output_sd=False):
    """utils/surface_rendering.py:129-165.  `roughness=None` (the live call, pano_mip_nerf.py:349-358): Lambertian
    shading, `l` is [B,D,3] (every row identical: the D env directions) or [D,3].  With a roughness [B,1] tensor: the
    microfacet branch (:147-151), `l` [B,D,3] per ray and `v` [B,3]; `shading` is None there, as upstream."""
    if roughness is not None:
        f = ops._f32c
        spec, nol = ops.brdf_terms(ops.BRDF_MICROFACET, normal.contiguous(), roughness.reshape(-1).contiguous(), f(l),
                                   f(v))
        rgb, diffuse, specular = ops.shade_sum(env.contiguous(), albedo.contiguous(), spec, nol,
                                               f(solid_angle).reshape(-1))
        return (rgb, diffuse, specular, None) if output_sd else (rgb, diffuse, specular)
    ld = l[0] if l.dim() == 3 else l
    rgb, shading = ops.shade(env.contiguous(), albedo.contiguous(), normal.contiguous(), ops._f32c(ld),
                             ops._f32c(solid_angle).reshape(-1))
    diffuse = rgb
    specular = torch.zeros_like(rgb)
    return (rgb, diffuse, specular, shading) if output_sd else (rgb, diffuse, specular)
