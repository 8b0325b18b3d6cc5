"""systems/mipnerf_system.py without Lightning: training_step (22-53) and the chunked render (95-127)."""
import torch

from .. import ops
from ..models.mip import rearrange_render_image
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
        rays, rgbs = batch[:2]
        _, height, width, _ = rgbs.shape
        chunks, _ = rearrange_render_image(rays, chunk_size or self.render_chunk())
        outs = [[] for _ in range(6)]
        with torch.no_grad():
            for batch_rays in chunks:
                (vol_c, dep_c, _, nor_c), (vol_f, dep_f, _, nor_f) = self.mip_nerf(
                    rays=batch_rays, randomized=self.val_randomized, white_bkgd=self.white_bkgd, use_ort_loss=True)
                for lst, v in zip(outs, (vol_c, vol_f, dep_c, dep_f, nor_c, nor_f)):
                    lst.append(v)

        def compose(x, dim=3):
            return torch.cat(x, dim=0).view(1, height, width, dim).permute(0, 3, 1, 2)

        return (compose(outs[0]), compose(outs[1]), compose(outs[2], 1), compose(outs[3], 1), compose(outs[4]),
                compose(outs[5]))
