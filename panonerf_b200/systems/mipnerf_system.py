"""systems/mipnerf_system.py without Lightning: training_step (22-53) and the chunked render (95-127)."""
import torch

from .. import ops
from .base_system import BaseSystem


class MipNeRFSystem(BaseSystem):
    def forward(self, batch_rays, randomized, white_bkgd, use_ort_loss=False):
        return self.mip_nerf(rays=batch_rays, randomized=randomized, white_bkgd=white_bkgd, use_ort_loss=use_ort_loss)

    def training_step(self, batch, batch_nb=0):
        rays, rgbs = batch[0], batch[1]
        ldr_rgb_gt = self._gt_ldr(rgbs)
        use_ort_loss = True if self.hparams["loss.ort_loss"] > 0 else False
        outputs = self.mip_nerf(rays=rays, randomized=self.train_randomized, white_bkgd=self.white_bkgd,
                                use_ort_loss=use_ort_loss)
        mask = ops._f32c(rays.lossmult).reshape(-1)
        inv = self._inv_mask_sum(mask)
        (vol_c, *_), (vol_f, _, ort_loss, _) = outputs
        vol_coarse = self._masked_mse(vol_c, ldr_rgb_gt, mask, inv)
        vol_fine = self._masked_mse(vol_f, ldr_rgb_gt, mask, inv)
        loss = self.hparams["loss.coarse_loss_mult"] * vol_coarse + vol_fine
        if use_ort_loss:
            loss = loss + self.hparams["loss.ort_loss"] * ort_loss
        return loss

    def render_image(self, batch, chunk_size=None):
        """systems/mipnerf_system.py:95-127: same 6-tuple of [1,C,H,W] images, produced by the render driver."""
        rays, rgbs = batch[:2]
        _, height, width, _ = rgbs.shape

        def forward(part):
            (vol_c, dep_c, _, nor_c), (vol_f, dep_f, _, nor_f) = self.mip_nerf(
                rays=part, randomized=self.val_randomized, white_bkgd=self.white_bkgd, use_ort_loss=True)
            return [vol_c, vol_f, dep_c, dep_f, nor_c, nor_f]

        return tuple(self._render_into(rays, height, width, (3, 3, 1, 1, 3, 3), forward, chunk_size))
