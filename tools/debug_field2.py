import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,'tests'))
import torch
from conftest import load_golden
from util import golden_rays, golden_state_dict
from test_models_gpu import _field_grads, _cos
from panonerf_b200 import field, ops
DEV='cuda'
g = load_golden("panonerf_w256.npz"); sd = golden_state_dict(g)
rays,_ = golden_rays(g, DEV); S=64
t, means, covs = ops.sample_cast(rays.origins, rays.directions, rays.radii, rays.near, rays.far, S)
venc = ops.pos_enc(rays.viewdirs, 4)
gen = torch.Generator().manual_seed(3); R=means.shape[0]
gens = (torch.randn(R, S, 3, generator=gen).to(DEV), torch.randn(R, S, 5, generator=gen).to(DEV), (torch.randn(R, S, 3, generator=gen) * 1e-2).to(DEV))
for need_means in (False, True):
    res={}
    for prec in ("bf16","bf16_simt","fp32"):
        res[prec]=_field_grads(prec, sd, means, covs, venc, S, gens, need_means)
        torch.cuda.synchronize()
    for i,nm in ((1,"n_raw"),(2,"raw_rgb"),(3,"raw_den")):
        a,b,c=[res[p][i].float() for p in ("bf16","bf16_simt","fp32")]
        print(need_means, nm, "tc-twin", float((a-b).abs().max()), "tc-f32", float((a-c).abs().max()), "twin-f32", float((b-c).abs().max()), "mean|.|", float(c.abs().mean()))
    for k in res["bf16"][0]:
        a,b,c=[res[p][0][k] for p in ("bf16","bf16_simt","fp32")]
        print(f"   {k:26s} cos tc-twin {_cos(a,b):+.4f} tc-f32 {_cos(a,c):+.4f} twin-f32 {_cos(b,c):+.4f}  norms {float(a.norm()):.3e} {float(b.norm()):.3e} {float(c.norm()):.3e}")
