"""SURVEY.md section 8f ranks 2 and 3 on the GPU: the render driver (whole image in one forward, per-ray results
scattered straight into [1,C,H,W]) against the CPU oracle and against the chunked walk, device-side PSNR / WS-PSNR
against the oracle's restatement of utils/metrics.py, and the OpenEXR / PNG writers read back with OpenCV / PIL."""
import os

import numpy as np
import pytest
import torch

from conftest import load_golden
from util import O, T, assert_close, golden_state_dict

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _mip_system(precision="fp32"):
    from panonerf_b200.systems.base_system import default_hparams
    from panonerf_b200.systems.mipnerf_system import MipNeRFSystem
    g = load_golden("mipnerf_w64.npz")
    hp = default_hparams("mipnerf", precision=precision)
    hp.update({"nerf.num_samples": 16, "nerf.mlp.net_width": 64, "train.randomized": False})
    system = MipNeRFSystem(hp).to(DEV)
    system.mip_nerf.mlp.load_state_dict(golden_state_dict(g))
    return system, g


def test_render_driver_64x128_panorama_matches_oracle():
    """A 64 x 128 panorama (8192 rays) rendered by ONE forward into [1,C,H,W] equals the oracle's per-ray outputs
    reshaped the way systems/mipnerf_system.py:118-127 does, and the chunked walk bit for bit."""
    from panonerf_b200.datasets.pano_datasets import generate_rays
    system, g = _mip_system()
    h, w = 64, 128
    rays = generate_rays(h, w, g["c2w"], 0.0, 10.0, DEV)
    batch = (type(rays)(*[x.view(1, h, w, -1) for x in rays]), torch.zeros(1, h, w, 3, device=DEV))
    full = system.render_image(batch)
    chunked = system.render_image(batch, chunk_size=1000)          # ragged last chunk
    for a, b in zip(full, chunked):
        assert a.is_contiguous() and torch.equal(a, b)
    assert full[1].shape == (1, 3, h, w) and full[3].shape == (1, 1, h, w)
    rays_c = O.equirect_rays(h, w, g["c2w"], 0.0, 10.0)
    ref, _ = O.mipnerf_forward(golden_state_dict(g), rays_c, dict(num_samples=16), use_ort_loss=True)
    compose = lambda x, d: x.reshape(1, h, w, d).permute(0, 3, 1, 2)
    for img, r, d, tol in ((full[0], ref[0][0], 3, 1e-5), (full[1], ref[1][0], 3, 1e-5), (full[2], ref[0][1], 1, 1e-5),
                           (full[3], ref[1][1], 1, 1e-5)):
        assert_close(img.cpu(), compose(r.detach(), d), tol, floor=1e-2)
    cos = (full[5].cpu() * compose(ref[1][3].detach(), 3)).sum(1)
    assert float((cos > 0.999).float().mean()) > 0.99


def test_device_metrics_match_oracle():
    from panonerf_b200.utils import metrics
    gen = torch.Generator().manual_seed(0)
    for (h, w) in ((16, 32), (128, 256)):
        a, b = torch.rand(3, h, w, generator=gen) * 2, torch.rand(3, h, w, generator=gen) * 2
        ad, bd = a.to(DEV), b.to(DEV)
        assert abs(float(metrics.calc_psnr(ad, bd)) - float(O.calc_psnr(a, b))) < 1e-4
        assert abs(float(metrics.calc_ws_psnr(ad, bd)) - float(O.calc_ws_psnr(a, b))) < 1e-4
        assert abs(float(metrics.calc_psnr(ad[None], bd[None])) - float(O.calc_psnr(a, b))) < 1e-4
    w_rows = metrics.solid_angle_rows(16, 32, torch.device(DEV)).cpu()
    ref = O.solid_angle_refinement(16, 32).reshape(16, 32)
    ref = ref / ref.sum()
    assert torch.equal(w_rows[:, None].expand(16, 32), ref)


def test_exr_and_png_writers_roundtrip(tmp_path):
    from panonerf_b200.utils.vis import save_results
    gen = torch.Generator().manual_seed(1)
    img = torch.rand(1, 3, 24, 40, generator=gen)
    img[0, :, 0, 0] = torch.tensor([0.0, 1.0, 0.5])
    hdr = img * 7.5
    save_results(img.to(DEV), tmp_path / "a.png")
    save_results(hdr.to(DEV), tmp_path / "a.exr")
    save_results(img[:, :1].to(DEV), tmp_path / "mono.png")
    save_results(hdr[:, :1].to(DEV), tmp_path / "mono.exr")
    from PIL import Image
    png = np.asarray(Image.open(tmp_path / "a.png"))
    assert png.dtype == np.uint8 and np.array_equal(png, O.png_pixels(img))            # utils/vis.py:35
    assert np.array_equal(np.asarray(Image.open(tmp_path / "mono.png")), O.png_pixels(img[:, :1]))
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    import cv2
    exr = cv2.imread(str(tmp_path / "a.exr"), cv2.IMREAD_UNCHANGED)
    if exr is None:
        pytest.skip("this OpenCV build cannot read OpenEXR")
    assert exr.dtype == np.float32 and np.array_equal(exr[..., ::-1], hdr[0].permute(1, 2, 0).numpy())
    mono = cv2.imread(str(tmp_path / "mono.exr"), cv2.IMREAD_UNCHANGED)
    assert np.array_equal(mono[..., 0], hdr[0, 0].numpy()) and np.array_equal(mono[..., 2], hdr[0, 0].numpy())
    with pytest.raises(NotImplementedError):
        save_results(img.to(DEV), tmp_path / "a.jpg")
