"""OpenEXR writer (utils/io_exr.py:30-47) without the OpenEXR dependency: the scan-line payload (FLOAT channels
B, G, R, no compression, increasing-Y) is laid out on the device (csrc/image.cu), the host adds the ~300-byte header
and the scan-line offset table and writes the file - one device-to-host copy per image."""
import struct

import numpy as np
import torch

from .. import ops


def _attr(name, typ, payload):
    return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(payload)) + payload


def exr_header(h, w):
    ch = b"".join(n + b"\0" + struct.pack("<iBxxxii", 2, 0, 1, 1) for n in (b"B", b"G", b"R")) + b"\0"   # FLOAT = 2
    box = struct.pack("<iiii", 0, 0, w - 1, h - 1)
    hdr = struct.pack("<ii", 20000630, 2)                                  # magic, version 2 (scan lines)
    hdr += _attr("channels", "chlist", ch)
    hdr += _attr("compression", "compression", b"\0")                      # NO_COMPRESSION
    hdr += _attr("dataWindow", "box2i", box)
    hdr += _attr("displayWindow", "box2i", box)
    hdr += _attr("lineOrder", "lineOrder", b"\0")                          # INCREASING_Y
    hdr += _attr("pixelAspectRatio", "float", struct.pack("<f", 1.0))
    hdr += _attr("screenWindowCenter", "v2f", struct.pack("<ff", 0.0, 0.0))
    hdr += _attr("screenWindowWidth", "float", struct.pack("<f", 1.0))
    return hdr + b"\0"


def write_exr(filename, data):
    """`data`: CUDA tensor [C,H,W] / [1,C,H,W] (C in {1,3}) - or an [H,W,C] float32 numpy array as upstream takes."""
    assert ".exr" in filename, "extension must be .exr"
    if isinstance(data, np.ndarray):
        assert data.dtype == np.float32, f"Data type is {data.dtype}, should be np.float32"
        data = torch.from_numpy(np.ascontiguousarray(data.transpose(2, 0, 1))).cuda()
    if data.dim() == 4:
        data = data[0]
    data = ops._f32c(data).contiguous()
    c, h, w = data.shape
    payload = ops.exr_payload(data).cpu().numpy().tobytes()
    hdr = exr_header(h, w)
    line = 8 + 12 * w
    base = len(hdr) + 8 * h
    table = struct.pack(f"<{h}Q", *[base + y * line for y in range(h)])
    with open(filename, "wb") as f:
        f.write(hdr + table + payload)
