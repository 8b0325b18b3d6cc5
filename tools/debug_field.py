import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,'tests'))
import torch
from conftest import load_golden
from util import golden_rays, golden_state_dict
from panonerf_b200 import field, ops
DEV='cuda'
g = load_golden("panonerf_w256.npz"); sd = golden_state_dict(g)
rays,_ = golden_rays(g, DEV); S=64
t, means, covs = ops.sample_cast(rays.origins, rays.directions, rays.radii, rays.near, rays.far, S)
venc = ops.pos_enc(rays.viewdirs, 4)
outs={}
for prec in ("bf16","bf16_simt","fp32"):
    params = {k: v.clone().to(DEV) for k, v in sd.items()}
    with torch.no_grad():
        outs[prec] = field.radiance_field(means, covs, venc, params, precision=prec, samples_per_ray=S, min_deg=0, max_deg=16, density_bias=-1.0, skip=4, with_normals=True)
for i,nm in enumerate(("raw_rgb","raw_den","n_raw")):
    a,b,c = [outs[p][i].reshape(-1, outs[p][i].shape[-1]).float() for p in ("bf16","bf16_simt","fp32")]
    d=(a-b).abs(); rows=d.max(dim=1).values
    print(nm, "tc-vs-twin max", float(d.max()), "mean", float(d.mean()), "| mean|b|", float(b.abs().mean()), " tc-vs-fp32 max", float((a-c).abs().max()), "twin-vs-fp32 max", float((b-c).abs().max()))
    bad = torch.nonzero(rows > 0.05*float(b.abs().mean())).flatten()
    print("   bad rows", bad.numel(), bad[:40].tolist())
