import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
from conftest import load_golden
from test_models_gpu import build
from util import T

def grads(name, pano, prec, surf=True, ort=0.1, chrom=0.1, jac=None):
    g = load_golden(name)
    system, rays, gt = build(g, pano, prec)
    system.mip_nerf.jac_precision = jac
    system.hparams['train.surface'] = surf
    system.hparams['loss.ort_loss'] = ort
    system.hparams['loss.chrom_loss'] = chrom
    loss = system.training_step((rays, gt))
    loss.backward()
    return float(loss), {k: p.grad.detach().double().flatten() for k, p in system.mip_nerf.mlp.named_parameters()}

for name, pano, kw in (("mipnerf_w256.npz", False, {}), ("panonerf_w256.npz", True, {})):
    la, a = grads(name, pano, "bf16_simt", **kw)
    lb, b = grads(name, pano, "bf16", **kw)
    print(name, kw, "loss", la, lb)
    for k in a:
        cos = float((a[k] @ b[k]) / (a[k].norm() * b[k].norm() + 1e-30))
        print(f"   {k:28s} cos={cos:+.4f} |fp32|={float(a[k].norm()):.4e} |bf16|={float(b[k].norm()):.4e}")
