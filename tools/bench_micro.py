"""Per-kernel micro-benchmarks of the memory-bound stages against the HBM roofline (BASELINE.json configs[4],
SURVEY.md section 8d): IPE, sample+cast, compositing (fwd), resampling.  One JSON line per kernel:
algorithmic bytes / CUDA-event time vs the measured copy bandwidth in MEASURED_PEAKS.json.
python tools/bench_micro.py [log2_samples]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from panonerf_b200 import ops  # noqa: E402


def timeit(fn, groups=5, per_group=10):
    """Median over `groups` of the mean device time of `per_group` back-to-back calls (CUDA events on the launching
    stream).  One event pair per CALL would add the host-side cost of the Python wrapper (output allocation + ctypes,
    20-40 us) to every sample - the device idles between the start event and the launch - which is a fifth of a 0.1 ms
    kernel; back to back the host runs ahead of the device and only the first call of a group pays it.  Every buffer is
    far larger than the 126 MB L2, so consecutive calls do not feed each other from cache."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(groups):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(per_group):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / per_group)
    return sorted(ts)[len(ts) // 2]


def run(lg=24, dev=None, peak=None):
    """-> one dict per kernel: algorithmic bytes / device time against the measured copy bandwidth."""
    N = 64
    R = (1 << lg) // N
    M = R * N
    dev = dev or torch.device("cuda", 0)
    if peak is None:
        peak = 6650.0
        pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pp):
            peak = json.load(open(pp))["hbm_gbs"]
    g = torch.Generator(device=dev).manual_seed(0)
    means = (torch.rand(M, 3, device=dev, generator=g) * 10 - 5)
    covs = torch.rand(M, 3, device=dev, generator=g)
    enc16 = torch.empty(M, 96, device=dev, dtype=torch.bfloat16)
    enc32 = torch.empty(M, 96, device=dev, dtype=torch.float32)
    vdir = torch.randn(M, 3, device=dev, generator=g)
    rgb = torch.rand(R, N, 3, device=dev, generator=g)
    den = -torch.log(torch.rand(R, N, device=dev, generator=g).clamp_min(1e-6))
    raw_den5 = torch.randn(M, 5, device=dev, generator=g)
    t = torch.sort(torch.rand(R, N + 1, device=dev, generator=g) * 10, dim=1).values.contiguous()
    dirs = torch.nn.functional.normalize(torch.randn(R, 3, device=dev, generator=g), dim=-1)
    w = torch.rand(R, N, device=dev, generator=g)
    origins = torch.zeros(R, 3, device=dev)
    radii = torch.full((R, 1), 1e-3, device=dev)
    near, far = torch.zeros(R, 1, device=dev), torch.full((R, 1), 10.0, device=dev)
    cases = [
        ("ipe_fwd(bf16 out)", lambda: ops.ipe_into(means, covs, 0, 16, enc16), M * (24 + 192), M),
        ("ipe_fwd(fp32 out)", lambda: ops.ipe_into(means, covs, 0, 16, enc32), M * (24 + 384), M),
        ("ipe_vjp(bf16 rows)", lambda: ops.ipe_vjp(means, covs, 0, 16, enc16), M * (36 + 192), M),
        ("ipe_vjp(fp32 rows)", lambda: ops.ipe_vjp(means, covs, 0, 16, enc32), M * (36 + 384), M),
        ("ipe_jvp(bf16 rows)", lambda: ops.ipe_jvp_into(means, covs, 0, 16, vdir, enc16), M * (36 + 192), M),
        ("sample_cast", lambda: ops.sample_cast(origins, dirs, radii, near, far, N), R * (36 + 4 * (N + 1) + 24 * N), M),
        ("composite_fwd", lambda: ops.composite(rgb, den, t, dirs, False), M * 24 + R * 36, M),
        # activations inside the compositing kernel: raw heads in (12 + 4 C B / sample, C = 5), weights + albedos out
        ("act_composite_fwd (C = 5, with albedos)",
         lambda: ops.act_composite(rgb.view(M, 3), raw_den5, t, dirs, False, -1.0, 0.001, True), M * (12 + 20 + 4 + 4 + 12) + R * 36, M),
        ("resample", lambda: ops.resample(t, w, 0.01), M * 12 + R * 8, M),
        ("resample_cast (resample_along_rays in one launch)", lambda: ops.resample(t, w, 0.01, cast=(origins, dirs, radii)),
         M * 36 + R * (8 + 28), M),
    ]
    out = []
    with torch.no_grad():
        for name, fn, nbytes, units in cases:
            ms = timeit(fn)
            gbs = nbytes / ms / 1e6
            out.append({"kernel": name, "samples": units, "ms": ms, "algorithmic_GBps": gbs, "hbm_peak_GBps": peak,
                        "frac_of_hbm_roofline": gbs / peak, "Gsamples_per_s": units / ms / 1e6})
    return out


def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24          # 16.8 M samples: every buffer >> 126 MB L2
    for rec in run(lg):
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
