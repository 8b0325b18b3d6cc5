"""systems/panonerf_system.py without Lightning: training_step (15-75) and the chunked render (133-192)."""
import torch

from .. import ops
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
        """systems/panonerf_system.py:133-192: same 9-tuple of [1,C,H,W] images, produced by the render driver."""
        rays, rgbs = batch[:2]
        _, height, width, _ = rgbs.shape

        def forward(part):
            (c_rgb, c_dep, *_), (f_rgb, f_dep, _, f_nor, alb, rhn, sf_rgb, _, sd) = self.mip_nerf(
                rays=part, env_rays=self.env_rays, randomized=self.val_randomized, white_bkgd=self.white_bkgd,
                enable_surf=True, use_ort_loss=True)
            return [c_rgb, f_rgb, c_dep, f_dep, f_nor, alb, sf_rgb, sd]

        c_rgb, f_rgb, c_dep, f_dep, nor, alb, sf, sd = self._render_into(rays, height, width, (3, 3, 1, 1, 3, 3, 3, 3),
                                                                        forward, chunk_size)
        return c_rgb, f_rgb, c_dep, f_dep, nor, alb, [], sf, sd
