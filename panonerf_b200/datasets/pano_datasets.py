"""Device-side replacements for the ray producers of datasets/pano_datasets.py:
`PanoDataset._generate_rays` (152-216) -> `generate_rays` (kernel K1) and `generate_lit_rays` (218-263).
Image / EXR loading is out of scope."""
import numpy as np
import torch

from .. import ops
from .base_datasets import Rays


def generate_rays(h, w, c2w, near=0.0, far=10.0, device="cuda", row0=0, nrows=None) -> Rays:
    """Equirectangular rays of one camera as flat [rows*W, .] fp32 CUDA tensors (rows [row0,row0+nrows) only, so
    that each GPU of a ray-sharded render generates exactly its own block)."""
    return Rays(*ops.raygen_equirect(h, w, c2w, near, far, torch.device(device), row0, nrows))


def pixel_radius(h, w, c2w, device="cuda") -> float:
    """`self.radii = radii[0][0,0,0]` of pano_datasets.py:215 (radius of pixel (0,0))."""
    r = ops.raygen_equirect(h, w, c2w, 0.0, 1.0, torch.device(device), 0, 1)[3]
    return float(r[0, 0])


def generate_lit_rays(radius, num=80, near=0, far=10.0, type=torch.float16, device="cuda") -> Rays:
    """Fibonacci-sphere environment directions (pano_datasets.py:218-263).  D<=80 values, computed once on the host
    in fp64 exactly like the reference, then quantised to `type` (fp16 upstream) and moved to the device."""
    phi = np.pi * (3.0 - np.sqrt(5.0))
    i = np.arange(num, dtype=np.float64)
    y = 1 - (i / float(num - 1)) * 2
    rad = np.sqrt(1 - y * y)
    d = np.stack([np.cos(phi * i) * rad, y, np.sin(phi * i) * rad], -1)
    one = np.ones((num, 1))
    view = d / np.linalg.norm(d, axis=-1, keepdims=True)
    fields = (np.zeros_like(d), d, view, float(radius) * one, (4 * np.pi / num) * one, near * one, far * one, 0 * one)
    return Rays(*[torch.tensor(x).to(type).to(device) for x in fields])


class DeviceRayFeed:
    """Device-side training-ray feed: replaces `DataLoader` + the per-ray
    `PanoDataset.__getitem__` (datasets/pano_datasets.py:271-281, systems/base_system.py:89-96), which hands the
    trainer one 8-field namedtuple per ray from 28 worker processes.

    The HDR image pool and the rays of every camera (kernel K1, generated once) stay on the GPU as one packed
    [n_pixels, 14] fp32 table; a batch is `batch_size` pixel ids drawn on the device ('all_images' batching: uniform
    over all pixels of all images) and two gathers.  Nothing touches the host per step."""

    WIDTHS = (3, 3, 3, 1, 1, 1, 1, 1)          # Rays fields: 14 floats = 56 B per ray

    def __init__(self, images, c2ws, near=0.0, far=10.0, device="cuda", seed=0):
        device = torch.device(device)
        assert len(images) == len(c2ws) and len(images) > 0
        packed, gts = [], []
        for img, c2w in zip(images, c2ws):
            img = torch.as_tensor(img, dtype=torch.float32)
            h, w = img.shape[:2]
            rays = generate_rays(h, w, np.asarray(c2w, dtype=np.float32), near, far, device)
            packed.append(torch.cat(list(rays), dim=1))
            gts.append(img.reshape(h * w, -1)[:, :3].to(device))
        self.packed = torch.cat(packed, dim=0).contiguous()
        self.gt = torch.cat(gts, dim=0).contiguous()
        self.n = self.packed.shape[0]
        self.gen = torch.Generator(device=device)
        self.gen.manual_seed(seed)

    def __len__(self):
        return self.n

    def unpack(self, packed) -> Rays:
        out, o = [], 0
        for wd in self.WIDTHS:
            out.append(packed[:, o:o + wd].contiguous())
            o += wd
        return Rays(*out)

    def sample(self, batch_size, return_ids=False):
        """-> (Rays of [batch_size, .], HDR ground truth [batch_size, 3]) on the device."""
        ids = torch.randint(self.n, (batch_size,), device=self.packed.device, generator=self.gen)
        rays = self.unpack(self.packed.index_select(0, ids))
        gt = self.gt.index_select(0, ids)
        return (rays, gt, ids) if return_ids else (rays, gt)
