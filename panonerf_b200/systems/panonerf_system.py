"""systems/panonerf_system.py without Lightning: training_step (15-75) and the chunked render (133-192)."""
import torch

from .. import ops
from ..models.mip import rearrange_render_image
from .base_system import BaseSystem


class PanoNeRFSystem(BaseSystem):
    def training_step(self, batch, batch_index=0):
        rays, rgbs = batch[0], batch[1]
        ldr_rgb_gt = self._gt_ldr(rgbs)
        hp = self.hparams
        surf_on = bool(self.global_step >= hp["train.surface_start_step"] and hp["train.surface"])
        use_ort_loss = True if hp["loss.ort_loss"] > 0 else False
        outputs = self.mip_nerf(rays=rays, env_rays=self.env_rays, randomized=self.train_randomized,
                                white_bkgd=self.white_bkgd, enable_surf=surf_on, use_ort_loss=use_ort_loss)
        mask = ops._f32c(rays.lossmult).reshape(-1)
        inv = self._inv_mask_sum(mask)
        (rgb_c, *_), (rgb_f, _, ort_loss, _, alb, _, sf_rgb, _, _) = outputs
        vol_coarse = self._masked_mse(rgb_c, ldr_rgb_gt, mask, inv)
        vol_fine = self._masked_mse(rgb_f, ldr_rgb_gt, mask, inv)
        loss = hp["loss.coarse_loss_mult"] * vol_coarse + vol_fine
        if surf_on:
            loss = loss + hp["loss.surface_loss"] * self._masked_mse(sf_rgb, ldr_rgb_gt, mask, inv)
            if hp["loss.chrom_loss"] > 0:
                loss = loss + hp["loss.chrom_loss"] * ops.chroma_loss(ldr_rgb_gt, alb)
        if ort_loss is not None:
            loss = loss + hp["loss.ort_loss"] * ort_loss
        return loss

    def render_image(self, batch, chunk_size=None):
        rays, rgbs = batch[:2]
        _, height, width, _ = rgbs.shape
        chunks, _ = rearrange_render_image(rays, chunk_size or self.render_chunk())
        keys = ("coarse_rgb", "fine_rgb", "coarse_dep", "fine_dep", "fine_nor", "albedo", "surface_rgb", "shading")
        acc = {k: [] for k in keys}
        with torch.no_grad():
            for batch_rays in chunks:
                (c_rgb, c_dep, *_), (f_rgb, f_dep, _, f_nor, alb, rhn, sf_rgb, _, sd) = self.mip_nerf(
                    rays=batch_rays, env_rays=self.env_rays, randomized=self.val_randomized,
                    white_bkgd=self.white_bkgd, enable_surf=True, use_ort_loss=True)
                for k, v in zip(keys, (c_rgb, f_rgb, c_dep, f_dep, f_nor, alb, sf_rgb, sd)):
                    if v is not None:
                        acc[k].append(v)

        def compose(x, dim=3):
            return torch.cat(x, dim=0).view(1, height, width, dim).permute(0, 3, 1, 2) if len(x) else None

        return (compose(acc["coarse_rgb"]), compose(acc["fine_rgb"]), compose(acc["coarse_dep"], 1),
                compose(acc["fine_dep"], 1), compose(acc["fine_nor"]), compose(acc["albedo"]), [],
                compose(acc["surface_rgb"]), compose(acc["shading"]))
