"""CUDA replacements for the live functions of utils/surface_rendering.py."""
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


def surface_rendering(env, albedo, normal, roughness, l, v, solid_angle, output_sd=False):
    """Lambertian branch of utils/surface_rendering.py:129-165 (`roughness=None`; the microfacet branch is dead
    code upstream).  `l` is [B,D,3] (every row identical: the D env directions) or [D,3]."""
    if roughness is not None:
        raise NotImplementedError("microfacet BRDF branch is unused by the reference hot path")
    ld = l[0] if l.dim() == 3 else l
    rgb, shading = ops.shade(env.contiguous(), albedo.contiguous(), normal.contiguous(), ops._f32c(ld),
                             ops._f32c(solid_angle).reshape(-1))
    diffuse = rgb
    specular = torch.zeros_like(rgb)
    return (rgb, diffuse, specular, shading) if output_sd else (rgb, diffuse, specular)
