"""Micro-benchmark of pnb_wgrad_batch alone on the job lists of one configs/panonerf.yaml backward pass
(field.py:_backward_fused): coarse / env level (15 jobs) and fine level (15 + 10 adjoint jobs).
python tools/bench_wgrad.py [M]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from panonerf_b200 import field  # noqa: E402


def jobs(wb, acts, dz, enc, pads, G, q=None, u=None):
    f32 = dict(device=acts.device, dtype=torch.float32)
    if q is not None:
        wb.add(acts, 10, 256, u, 0, 96, G["w0"])
        for i in range(1, 8):
            if i == 5:
                wb.add(acts, 15, 256, q, 4, 256, G["w5"][:, :256])
                wb.add(acts, 15, 256, u, 0, 96, G["w5"][:, 256:])
            else:
                wb.add(acts, 10 + i, 256, q, i - 1, 256, G[f"w{i}"])
        wb.add(q, 7, 256, u, 0, 16, torch.zeros(256, 16, **f32), G["ws"])
    wb.add(acts, 9, 128, pads[0], 0, 64, torch.zeros(128, 64, **f32))
    wb.add(dz, 0, 128, acts, 8, 256, G["wv"], G["bv"])
    wb.add(dz, 1, 256, acts, 7, 256, G["we"], G["be"])
    wb.add(acts, 7, 256, pads[1], 0, 64, torch.zeros(256, 64, **f32))
    for i in range(7, 0, -1):
        if i == 5:
            wb.add(dz, 4, 256, acts, 4, 256, G["w5"][:, :256], G["b5"])
            wb.add(dz, 4, 256, enc, 0, 96, G["w5"][:, 256:])
        else:
            wb.add(dz, 9 - i, 256, acts, i - 1, 256, G[f"w{i}"], G[f"b{i}"])
    wb.add(dz, 9, 256, enc, 0, 96, G["w0"], G["b0"])


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    M = int(args[0]) if args else 8192 * 64
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    bf = dict(device=dev, dtype=torch.bfloat16)
    acts = torch.randn(18, M, 256, **bf)
    dz = torch.randn(10, M, 256, **bf)
    q = torch.randn(8, M, 256, **bf)
    enc, u = torch.randn(M, 96, **bf), torch.randn(M, 96, **bf)
    pads = [torch.randn(M, 64, **bf), torch.randn(M, 64, **bf)]
    f32 = dict(device=dev, dtype=torch.float32)
    G = {f"w{i}": torch.zeros(256, 352 if i == 5 else (96 if i == 0 else 256), **f32) for i in range(8)}
    G.update({f"b{i}": torch.zeros(256, **f32) for i in range(8)})
    G.update(wv=torch.zeros(128, 256, **f32), bv=torch.zeros(128, **f32), we=torch.zeros(256, 256, **f32),
             be=torch.zeros(256, **f32), ws=torch.zeros(256, **f32))
    for name, with_adj in (("base level (15 jobs)", False), ("fine level (25 jobs)", True)):
        wb = field.WgradBatch(M, dev)
        jobs(wb, acts, dz, enc, pads, G, q if with_adj else None, u)
        nbytes = sum(M * 2 * (j[4] + j[5]) for j in wb.jobs)
        flops = sum(2 * M * j[4] * j[5] for j in wb.jobs)
        for _ in range(3):
            wb.launch()
        torch.cuda.synchronize()
        times = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                wb.launch()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) / 5)
        ms = min(times)
        print(json.dumps(dict(kernel="wgrad_batch", jobs=name, M=M, ms=round(ms, 4), GBps=round(nbytes / ms / 1e6, 1),
                              TFLOPs=round(flops / ms / 1e9, 1), lib=os.environ.get("PNB_LIB_PATH", "in-tree"))))


if __name__ == "__main__":
    main()
