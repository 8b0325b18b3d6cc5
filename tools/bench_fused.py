"""Micro-benchmark of the fused MLP kernel alone: TFLOP/s vs the measured bf16 peak (MEASURED_PEAKS.json).
python tools/bench_fused.py [samples] [--normals] [--save]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from panonerf_b200 import field, ops  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    M = int(args[0]) if args else 128 * 148 * 64
    normals, save = "--normals" in sys.argv, "--save" in sys.argv
    S, C = int(os.environ.get("BF_S", "64")), 5
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    from oracle import panonerf_oracle as O
    sd = O.synth_state_dict(seed=4, width=256, c_density=C)
    names = field.param_names(8, 1)
    params = [sd[n].to(dev).contiguous() for n in names]
    pack = field.fused_pack(names, params)
    means = (torch.rand(M, 3, device=dev) * 2 - 1) * 3
    covs = torch.rand(M, 3, device=dev) * 1e-3
    enc = torch.empty(M, 96, device=dev, dtype=torch.bfloat16)
    ops.ipe_into(means, covs, 0, 16, enc)
    vb = torch.randn((M + S - 1) // S, 128, device=dev)
    acts = torch.empty(18, M, 256, device=dev, dtype=torch.bfloat16) if save else None
    g_enc = torch.empty(M, 96, device=dev) if normals else None
    flops = M * (2 * (96 * 256 + 6 * 256 * 256 + 352 * 256 + 256 * C + 256 * 256 + 256 * 128 + 128 * 3)
                 + (2 * (6 * 256 * 256 + 2 * 96 * 256) if normals else 0))
    masks = field.fused_masks(M, dev, save)
    kernel = "mlp_fused"
    run = lambda: field.fused_forward(enc, vb, S, C, pack, acts, g_enc, masks, save)
    if "--ipe" in sys.argv:      # IPE computed inside the kernel (encoder warps)
        run = lambda: field.fused_forward_ipe(means, covs, 0, vb, 0, S, C, pack, g_enc)
        kernel = "mlp_fused(in-kernel ipe)"
    if "--bwd" in sys.argv or "--jadj" in sys.argv:
        masks = field.fused_masks(M, dev, True)
        field.fused_forward(enc, vb, S, C, pack, None, None, masks, True)        # real sign bits
        if "--bwd" in sys.argv:
            d_rgb, d_den = torch.randn(M, 3, device=dev), torch.randn(M, C, device=dev)
            d_enc = torch.empty(M, 96, device=dev)
            flops = M * 2 * (128 * 256 + 256 * 256 + 7 * 256 * 256 + 2 * 96 * 256)
            kernel = "mlp_fused_bwd"
            run = lambda: field.fused_backward(M, C, pack, d_rgb, d_den, masks, d_enc)
        else:
            flops = M * 2 * (96 * 256 + 6 * 256 * 256 + 352 * 256)
            kernel = "mlp_fused_jadj"
            run = lambda: field.fused_jadj(enc, pack, masks)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    times = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = sorted(times)[len(times) // 2]
    peaks = {}
    pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pp):
        peaks = json.load(open(pp))
    tf = flops / ms / 1e9
    print(json.dumps({"kernel": kernel, "samples": M, "normals": normals, "save": save, "ms": ms,
                      "tflops": tf, "frac_of_burst_peak": tf / peaks.get("bf16_tflops", 1685.0),
                      "samples_per_s": M / ms * 1e3, "all_ms": times}))


if __name__ == "__main__":
    main()
