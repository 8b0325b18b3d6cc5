"""Shared helpers for the parity tests."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import panonerf_oracle as O  # noqa: E402  (the checker; never imported by the product)


def T(x, device="cpu"):
    return torch.from_numpy(np.ascontiguousarray(x)).to(device)


def rel_err(a, b, floor=None):
    """max |a-b| / max(|b|, floor) with floor defaulting to the mean magnitude of b (scale-aware relative error)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    if floor is None:
        floor = float(b.abs().mean()) + 1e-30
    return float(((a - b).abs() / torch.clamp_min(b.abs(), floor)).max())


def assert_close(a, b, rtol, name="", floor=None):
    e = rel_err(a, b, floor)
    assert e <= rtol, f"{name}: relative error {e:.3e} > {rtol:.1e}"


def golden_state_dict(g):
    if "sd_seed" in g:
        width = int(g["width"])
        c = 5 if "out/1/albedo" in g else 1
        sd = O.synth_state_dict(seed=int(g["sd_seed"]), width=width, c_density=c)
        chk = np.array([float(sum(v.double().sum() for v in sd.values())),
                        float(sum((v.double() ** 2).sum() for v in sd.values()))])
        assert np.allclose(chk, g["sd_checksum"], rtol=1e-10), "synthetic weights are not reproducible here"
        return sd
    return {k[3:]: T(v) for k, v in g.items() if k.startswith("sd/")}


def golden_rays(g, device="cpu"):
    h, w = [int(v) for v in g["hw"]]
    rays = O.equirect_rays(h, w, g["c2w"], 0.0, 10.0)
    perm = T(g["perm"]).long()
    rays = O.Rays(*[x[perm].contiguous().to(device) for x in rays])
    env = O.fibonacci_env_rays(10, float(g["env_radius"]))
    env = O.Rays(*[x.to(device) for x in env])
    return rays, env
