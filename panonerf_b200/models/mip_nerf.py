"""B200-native `MipNeRF` with the reference's constructor, `.mlp` parameter names/shapes, forward signature and
return tuples (models/mip_nerf.py:15-283).  The arithmetic runs in libpanonerf_b200.so (no PyTorch math ops)."""
import os
from collections import OrderedDict, namedtuple

import torch

from .. import field, ops


def _xavier_init(linear):
    torch.nn.init.xavier_uniform_(linear.weight.data)


class PureMLP(torch.nn.Module):
    """Parameter container with the exact module tree of the reference MLP (models/mip_nerf.py:15-102 ==
    models/pano_mip_nerf.py:17-114) so that checkpoints interchange; evaluation goes through field.radiance_field."""

    def __init__(self, net_depth, net_width, net_depth_condition, net_width_condition, skip_index, num_rgb_channels,
                 num_density_channels, activation, xyz_dim, view_dim):
        super().__init__()
        if activation != "relu":
            raise NotImplementedError
        if net_depth_condition != 1:
            raise NotImplementedError("net_depth_condition must be 1 (configs/*.yaml)")
        self.skip_index = skip_index
        layers = []
        for i in range(net_depth):
            if i == 0:
                dim_in = xyz_dim
            elif (i - 1) % skip_index == 0 and i > 1:
                dim_in = net_width + xyz_dim
            else:
                dim_in = net_width
            linear = torch.nn.Linear(dim_in, net_width)
            _xavier_init(linear)
            layers.append(torch.nn.Sequential(linear, torch.nn.ReLU(True)))
        if sum(1 for i in range(net_depth) if (i - 1) % skip_index == 0 and i > 1) > 1:
            raise NotImplementedError("at most one skip connection (net_depth <= 2*skip_index+1) is supported")
        self.layers = torch.nn.ModuleList(layers)
        self.density_layer = torch.nn.Linear(net_width, num_density_channels)
        _xavier_init(self.density_layer)
        self.extra_layer = torch.nn.Linear(net_width, net_width)
        _xavier_init(self.extra_layer)
        layers = []
        for i in range(net_depth_condition):
            dim_in = net_width + view_dim if i == 0 else net_width_condition
            linear = torch.nn.Linear(dim_in, net_width_condition)
            _xavier_init(linear)
            layers.append(torch.nn.Sequential(linear, torch.nn.ReLU(True)))
        self.view_layers = torch.nn.Sequential(*layers)
        self.color_layer = torch.nn.Linear(net_width_condition, num_rgb_channels)

    def named_field_params(self):
        return OrderedDict((k, v) for k, v in self.named_parameters())

    def forward(self, x, view_direction=None):
        raise RuntimeError("the MLP is evaluated by the fused CUDA field (panonerf_b200.field), not layer by layer")


def default_precision():
    return os.environ.get("PANONERF_PRECISION", "bf16")


class _NerfBase(torch.nn.Module):
    """Shared plumbing of MipNeRF / PanoMipNeRF."""

    def _setup(self, kw, mlp_cls):
        self.num_levels = kw["num_levels"]
        self.num_samples = kw["num_samples"]
        self.disparity = kw["disparity"]
        self.ray_shape = kw["ray_shape"]
        self.disable_integration = kw["disable_integration"]
        self.min_deg_point = kw["min_deg_point"]
        self.max_deg_point = kw["max_deg_point"]
        self.use_viewdirs = kw["use_viewdirs"]
        self.deg_view = kw["deg_view"]
        self.density_noise = kw["density_noise"]
        self.density_bias = kw["density_bias"]
        self.resample_padding = kw["resample_padding"]
        self.stop_resample_grad = kw["stop_resample_grad"]
        self.rgb_padding = kw["rgb_padding"]
        self.precision = kw.get("precision") or default_precision()
        self.jac_precision = kw.get("jac_precision") or os.environ.get("PANONERF_JAC_PRECISION")
        if kw["rgb_activation"] != "softplus":
            raise NotImplementedError
        if kw["density_activation"] != "softplus":
            raise NotImplementedError
        if not self.use_viewdirs:
            raise NotImplementedError("use_viewdirs=False is not supported (the reference's colour head needs it)")
        if self.density_noise:
            raise NotImplementedError("density_noise > 0 crashes upstream on GPU (SURVEY App. B.17); not implemented")
        if self.ray_shape == "cylinder":
            pass  # raised at call time like upstream (models/mip.py:83-84)
        xyz = (self.max_deg_point - self.min_deg_point) * 3 * 2
        view = self.deg_view * 3 * 2
        view = view + 3 if kw["append_identity"] else view
        if not kw["append_identity"]:
            raise NotImplementedError("compute_graph always appends the identity (pano_mip_nerf.py:257)")
        self.mlp = mlp_cls(kw["mlp_net_depth"], kw["mlp_net_width"], kw["mlp_net_depth_condition"],
                           kw["mlp_net_width_condition"], kw["mlp_skip_index"], kw["mlp_num_rgb_channels"],
                           kw["mlp_num_density_channels"], kw["mlp_net_activation"], xyz, view)

    # -- one level of the hot path -------------------------------------------------------------------------------
    def _field(self, means, covs, venc, samples_per_ray, with_normals, venc_mod=0):
        if self.disable_integration:
            covs = torch.zeros_like(covs)
        return field.radiance_field(means, covs, venc, self.mlp.named_field_params(), precision=self.precision,
                                    samples_per_ray=samples_per_ray, min_deg=self.min_deg_point,
                                    max_deg=self.max_deg_point, density_bias=self.density_bias,
                                    skip=self.mlp.skip_index, with_normals=with_normals,
                                    jac_precision=self.jac_precision, venc_mod=venc_mod)

    def _prep_rays(self, rays):
        f = ops._f32c
        return type(rays)(*[f(x) for x in rays])

    def _sample_level(self, lvl, rays, t, weights, randomized):
        from . import mip
        if lvl == 0:
            return mip.sample_along_rays(rays.origins, rays.directions, rays.radii, self.num_samples, rays.near,
                                         rays.far, randomized, self.disparity, self.ray_shape)
        return mip.resample_along_rays(rays.origins, rays.directions, rays.radii, t,
                                       weights.detach() if self.stop_resample_grad else weights, randomized,
                                       self.ray_shape, self.stop_resample_grad,
                                       resample_padding=self.resample_padding)


_DEFAULTS = dict(num_samples=128, num_levels=2, resample_padding=0.01, stop_resample_grad=True, use_viewdirs=True,
                 disparity=False, ray_shape="cone", min_deg_point=0, max_deg_point=16, deg_view=4,
                 density_activation="softplus", density_noise=0.0, density_bias=-1.0, rgb_activation="sigmoid",
                 alb_activation="sigmoid", rgb_padding=0.001, disable_integration=False, append_identity=True,
                 mlp_net_depth=8, mlp_net_width=256, mlp_net_depth_condition=1, mlp_net_width_condition=128,
                 mlp_skip_index=4, mlp_num_rgb_channels=3, mlp_num_density_channels=1, mlp_net_activation="relu")


class MipNeRF(_NerfBase):
    """models/mip_nerf.py:105-283.  Extra keyword `precision` ('bf16' tensor-core path | 'fp32' parity path);
    unknown keywords are swallowed like upstream (`**kwargs`, mip_nerf.py:134)."""

    def __init__(self, **kwargs):
        super().__init__()
        kw = {**_DEFAULTS, **kwargs}
        self._setup(kw, PureMLP)

    def forward(self, rays: namedtuple, randomized: bool, white_bkgd: bool, use_ort_loss: bool):
        rays = self._prep_rays(rays)
        venc = ops.pos_enc(rays.viewdirs, self.deg_view)
        ret = []
        t, weights = None, None
        for lvl in range(self.num_levels):
            t, (means, covs) = self._sample_level(lvl, rays, t, weights, randomized)
            want_normals = lvl == 1 and use_ort_loss
            raw_rgb, raw_den, n_raw = self._field(means, covs, venc, means.shape[1], want_normals)
            R, S = means.shape[0], means.shape[1]
            comp_rgb, distance, acc, weights, _ = ops.act_composite(
                raw_rgb.view(R * S, raw_rgb.shape[-1]), raw_den.view(R * S, raw_den.shape[-1]), t, rays.directions, white_bkgd, self.density_bias,
                self.rgb_padding, False)
            if want_normals:
                normal, ort, _ = ops.normals_aggregate(n_raw, weights, rays.directions, None)
                ret.append((comp_rgb, distance, ops.dmean(ort), normal))
            else:
                ret.append((comp_rgb, distance, None, torch.ones_like(comp_rgb)))
        return ret
