"""Measure (not assert) the parity numbers the tests bound: config-size goldens (C1 MipNeRF 4096 rays x 128+128,
C2 subset PanoMipNeRF 2048 rays x 64+64) in fp32 parity mode and on the bf16 tensor-core product path, and the
resampling index mismatch counts at 4096 rays.  Writes gpurun_out/parity_report.json.

    python tools/parity_report.py
"""
import hashlib
import json
import math
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import O, T, golden_rays, golden_state_dict  # noqa: E402
from conftest import load_golden  # noqa: E402

DEV = "cuda"
MIP_NAMES = ["comp_rgb", "distance", "ort_loss", "normal"]
PANO_NAMES = ["comp_rgb", "distance", "ort_loss", "normal", "albedo", "roughness", "surface_rgb", "diffuse", "shading"]


def build(g, pano, precision, **over):
    from panonerf_b200.systems.base_system import default_hparams
    from panonerf_b200.systems.mipnerf_system import MipNeRFSystem
    from panonerf_b200.systems.panonerf_system import PanoNeRFSystem
    hp = default_hparams("panonerf" if pano else "mipnerf", precision=precision)
    hp.update({"nerf.num_samples": int(g["n"]), "nerf.mlp.net_width": int(g["width"]), "train.randomized": False,
               "loss.ort_loss": 0.1})
    hp.update(over)
    system = (PanoNeRFSystem if pano else MipNeRFSystem)(hp).to(DEV)
    system.mip_nerf.mlp.load_state_dict(golden_state_dict(g))
    rays, env = golden_rays(g, DEV)
    system.env_rays = env
    return system, rays, T(g["gt"]).to(DEV)


def rel(a, b, floor=1e-3):
    a, b = a.double().cpu(), b.double().cpu()
    return float(((a - b).abs() / torch.clamp_min(b.abs(), max(float(b.abs().mean()), floor))).max())


def grads_report(system, g, prefix=""):
    rep = {}
    for k, p in system.mip_nerf.mlp.named_parameters():
        ref = T(g[prefix + "gslice/" + k]).double()
        n = ref.numel()
        got = p.grad.detach().cpu().double().reshape(-1)[:: max(1, p.numel() // n)][:n]
        gn_ref = float(g[prefix + "gnorm/" + k])
        rep[k] = {"slice_rel_err": float((got - ref).norm() / max(float(ref.norm()), 1e-30)),
                  "slice_cos": float((got @ ref) / (got.norm() * ref.norm() + 1e-30)),
                  "gnorm_ratio": float(p.grad.norm()) / max(gn_ref, 1e-30)}
        tkey = prefix + "g64/gslice/" + k
        if tkey in g:                       # float64 oracle value: how far are we / is the fp32 reference from it
            tr = T(g[tkey]).double()
            rep[k].update(truth_rel_err=float((got - tr).norm() / tr.norm()), truth_cos=float((got @ tr) / (got.norm() * tr.norm())),
                          ref_truth_rel_err=float((ref - tr).norm() / tr.norm()),
                          ref_truth_cos=float((ref @ tr) / (ref.norm() * tr.norm())))
    return rep


def model_report(name, pano):
    g = load_golden(name)
    names = PANO_NAMES if pano else MIP_NAMES
    out = {}
    for prec in ("fp32", "bf16"):
        t0 = time.time()
        system, rays, gt = build(g, pano, prec)
        if pano:
            res = system.mip_nerf(rays=rays, env_rays=system.env_rays, randomized=False, white_bkgd=False,
                                  enable_surf=True, use_ort_loss=True)
        else:
            res = system.mip_nerf(rays=rays, randomized=False, white_bkgd=False, use_ort_loss=True)
        rep = {}
        for lvl in range(2):
            for nm, v in zip(names, res[lvl]):
                key = f"out/{lvl}/{nm}"
                if key not in g or v is None:
                    continue
                ref = T(g[key])
                v = v.detach().cpu().reshape(ref.shape)
                r = {"rel": rel(v, ref), "max_abs": float((v - ref).abs().max())}
                if nm == "normal":
                    cos = (v * ref).sum(-1)
                    r.update(cos_mean=float(cos.mean()), cos_p01=float(cos.quantile(0.01)),
                             cos_p001=float(cos.quantile(0.001)), frac_cos_gt_0999=float((cos > 0.999).float().mean()),
                             frac_cos_gt_09999=float((cos > 0.9999).float().mean()))
                if nm == "comp_rgb":
                    mse = float(((v - ref) ** 2).mean())
                    r["psnr"] = 10 * math.log10(1.0 / max(mse, 1e-20))
                rep[key] = r
        loss = system.training_step((rays, gt))
        rep["loss"] = {"ours": float(loss), "ref": float(g["loss"]), "rel": abs(float(loss) - float(g["loss"])) / abs(float(g["loss"]))}
        loss.backward()
        rep["grads"] = grads_report(system, g)
        if "noort/loss" in g:
            system2, rays2, gt2 = build(g, pano, prec, **{"loss.ort_loss": 0})
            l0 = system2.training_step((rays2, gt2))
            rep["noort_loss"] = {"ours": float(l0), "ref": float(g["noort/loss"]),
                                 "rel": abs(float(l0) - float(g["noort/loss"])) / abs(float(g["noort/loss"]))}
            l0.backward()
            rep["noort_grads"] = grads_report(system2, g, "noort/")
        torch.cuda.synchronize()
        rep["seconds"] = time.time() - t0
        out[prec] = rep
    return out


def resample_report():
    from panonerf_b200 import ops
    g = load_golden("resample_large.npz")
    out = {}
    for n in (64, 128, 256):
        t, w, o, d, rad = O.resample_case(n)
        ref, inds_ref, _ = O.pdf_sample(t, O.blur_weights(w, 0.01), n + 1, False, return_aux=True)
        new_t, inds, means, covs = ops.resample(t.to(DEV), w.to(DEV), 0.01, return_inds=True,
                                                cast=(o.to(DEV), d.to(DEV), rad.to(DEV)))
        sha = np.frombuffer(hashlib.sha256(new_t.cpu().numpy().tobytes()).digest(), dtype=np.uint8)
        out[str(n)] = {"index_mismatch_vs_host_oracle": int((inds.cpu() != inds_ref).sum()),
                       "index_mismatch_vs_reference_golden": int((inds.cpu() != T(g[f"inds/{n}"].astype(np.int64))).sum()),
                       "new_t_bit_identical_to_golden": bool((sha == g[f"new_t_sha256/{n}"]).all()),
                       "new_t_values_differing_from_host_oracle": int((new_t.cpu() != ref).sum()),
                       "new_t_max_abs_vs_host_oracle": float((new_t.cpu() - ref).abs().max()),
                       "host_oracle_equals_golden": bool(torch.equal(inds_ref, T(g[f"inds/{n}"].astype(np.int64)))),
                       "mean_rows_max_abs": float((means.cpu()[::64] - T(g[f"mean_rows64/{n}"])).abs().max()),
                       "cov_rows_rel": rel(covs.cpu()[::64], T(g[f"cov_rows64/{n}"]), 1e-9)}
    return out


def main():
    rep = {"resample": resample_report()}
    print(json.dumps(rep["resample"], indent=1), flush=True)
    for name, pano in (("mipnerf_c1.npz", False), ("panonerf_c2s.npz", True)):
        rep[name] = model_report(name, pano)
        print(name, json.dumps(rep[name], indent=1)[:6000], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rep, open(os.path.join(ROOT, "gpurun_out", "parity_report.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
