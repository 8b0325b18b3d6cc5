"""B200-native `PanoMipNeRF` (models/pano_mip_nerf.py:117-363): mip-NeRF + density-gradient normals + albedo +
environment-ray irradiance + Lambertian surface rendering, same constructor / forward / return tuples."""
from collections import namedtuple

import torch

from .. import ops
from .mip_nerf import PureMLP, _NerfBase, _DEFAULTS


class MLP(PureMLP):
    """models/pano_mip_nerf.py:17-114 (identical architecture to PureMLP)."""


class PanoMipNeRF(_NerfBase):
    def __init__(self, **kwargs):
        super().__init__()
        kw = {**_DEFAULTS, "solid_angle_height": 8, "solid_angle_width": 16, "num_env_samples": 10, **kwargs}
        if kw["alb_activation"] != "sigmoid":
            raise NotImplementedError
        self._setup(kw, MLP)
        self.num_env_samples = kw["num_env_samples"]
        self.detach_dist = False          # pano_mip_nerf.py:189

    def forward(self, rays: namedtuple, env_rays: namedtuple, randomized: bool, white_bkgd: bool, enable_surf: bool,
                use_ort_loss: bool):
        rays = self._prep_rays(rays)
        venc = ops.pos_enc(rays.viewdirs, self.deg_view)
        ret = []
        t, weights = None, None
        for lvl in range(self.num_levels):
            t, (means, covs) = self._sample_level(lvl, rays, t, weights, randomized)
            R, S = means.shape[0], means.shape[1]
            fine = lvl == 1
            raw_rgb, raw_den, n_raw = self._field(means, covs, venc, S, fine)
            C = raw_den.shape[-1]
            comp_rgb, distance, acc, weights, albedos = ops.act_composite(
                raw_rgb.view(R * S, raw_rgb.shape[-1]), raw_den.view(R * S, C), t, rays.directions, white_bkgd, self.density_bias,
                self.rgb_padding, C >= 5 and fine and enable_surf)
            normal = surface_rgb = albedo = diffuse = ort_loss = shading = None
            if fine:
                normal, ort, albedo = ops.normals_aggregate(
                    n_raw, weights, rays.directions, albedos.view(R, S, 3) if albedos is not None else None)
                if use_ort_loss:
                    ort_loss = ops.dmean(ort)
                if enable_surf:
                    env = self._prep_rays(env_rays)
                    D = env.directions.shape[0]
                    Ne = self.num_env_samples
                    t_rand = torch.rand(1, Ne + 1, device=distance.device) if randomized else None
                    dist_in = distance.detach() if self.detach_dist else distance
                    lit_t, lit_means, lit_covs = ops.env_cast(rays.origins, rays.directions, dist_in, env.directions,
                                                             env.radii, env.near, env.far, Ne, t_rand)
                    lit_venc = ops.pos_enc(env.directions, self.deg_view)            # [D,27], one per env direction
                    e_rgb, e_den, _ = self._field(lit_means, lit_covs, lit_venc, Ne, False, venc_mod=D)
                    env_rgb = ops.act_composite(e_rgb.view(R * D * Ne, e_rgb.shape[-1]), e_den.view(R * D * Ne, C), lit_t,
                                                env.directions, False, self.density_bias, self.rgb_padding, False,
                                                d_mod=D)[0].view(R, D, 3)
                    surface_rgb, shading = ops.shade(env_rgb, albedo, normal, env.directions,
                                                     env.lossmult.reshape(-1).contiguous())
                    diffuse = surface_rgb
            ret.append((comp_rgb, distance, ort_loss, normal, albedo, None, surface_rgb, diffuse, shading))
        return ret
