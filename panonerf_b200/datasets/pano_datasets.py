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
