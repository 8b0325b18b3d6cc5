"""Device-side image metrics (utils/metrics.py:210-237 calc_mse / calc_psnr, :318-326 calc_ws_psnr): the squared
error is formed and reduced on the GPU in a fixed order (csrc/image.cu + pnb_sum); results are 0-dim device tensors
like upstream's, so validation does not round-trip the images through the host."""
import numpy as np
import torch

from .. import ops


def _chw(x):
    x = ops._f32c(x)
    if x.dim() == 4 and x.shape[0] == 1:
        x = x[0]
    if x.dim() != 3:
        raise RuntimeError("expected a [C,H,W] (or [1,C,H,W]) image")
    return x.contiguous()


def calc_mse(x, y):
    """utils/metrics.py:210-214."""
    x, y = _chw(x), _chw(y)
    return ops.image_sqerr_sum(x, y) * (1.0 / x.numel())


def calc_psnr(x, y):
    """utils/metrics.py:231-237."""
    return -10.0 * torch.log10(calc_mse(x, y))


_sa_cache = {}


def solid_angle_rows(h, w, device):
    """Per-row weights of calc_ws_psnr: utils/surface_rendering.py:294-316 (sin(phi) d_theta d_phi, constant along a
    row), cast to fp32 like `torch.Tensor(solid_angle)` and normalised by the fp32 sum over all H*W pixels."""
    key = (h, w, str(device))
    if key not in _sa_cache:
        d_phi, d_theta = np.pi / h, 2 * np.pi / w
        y = (np.arange(h) + 0.5) / h
        row = torch.tensor(np.sin(y * np.pi) * d_theta * d_phi, dtype=torch.float64).to(torch.float32)
        total = row[:, None].expand(h, w).reshape(1, -1, 1).sum()       # same tensor, same fp32 sum as upstream
        _sa_cache[key] = (row / total).to(device).contiguous()
    return _sa_cache[key]


def calc_ws_psnr(pred, gt):
    """utils/metrics.py:318-326: PSNR weighted by the normalised solid angle of every equirect pixel."""
    pred, gt = _chw(pred), _chw(gt)
    c, h, w = pred.shape
    mse = ops.image_sqerr_sum(pred, gt, solid_angle_rows(h, w, pred.device))
    return -10.0 * torch.log10(mse)
