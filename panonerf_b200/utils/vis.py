"""Image dumps of validation (utils/vis.py:25-41 save_results): PNG (8-bit, `(image * 255).astype(uint8)`) and
OpenEXR (float32).  Quantisation / channel interleave / scan-line layout happen on the device (csrc/image.cu); the
host only frames the bytes (zlib + CRC for PNG)."""
import os
import struct
import zlib

from .. import ops
from .io_exr import write_exr


def _chunk(tag, data):
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def save_results(image, save_path, compress_level=6):
    """`image`: CUDA tensor [1,C,H,W] with C in {1,3} (single-channel images are replicated like upstream)."""
    save_path = str(save_path)
    d = os.path.dirname(save_path)
    if d:
        os.makedirs(d, exist_ok=True)
    img = ops._f32c(image[0] if image.dim() == 4 else image).contiguous()
    c, h, w = img.shape
    if save_path.endswith(".png"):
        raw = ops.png_payload(img).cpu().numpy().tobytes()
        png = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0))
        png += _chunk(b"IDAT", zlib.compress(raw, compress_level)) + _chunk(b"IEND", b"")
        with open(save_path, "wb") as f:
            f.write(png)
    elif save_path.endswith(".exr"):
        write_exr(save_path, img)
    else:
        raise NotImplementedError
