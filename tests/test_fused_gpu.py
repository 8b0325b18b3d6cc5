"""GPU parity of the fused MLP kernel (csrc/mlp_fused.cu) against the layer-by-layer tensor-core path it replaces
(csrc/gemm_tc.cu), which is itself pinned to the fp32 parity path and the reference's golden vectors by
tests/test_models_gpu.py.  Both paths push the same bf16 operands through tcgen05 with fp32 accumulation and the same
epilogue order, so they must agree to bf16 rounding noise; the weight-blob packer is checked bit-exactly."""
import os

import pytest
import torch

from conftest import load_golden
from util import assert_close, golden_rays, golden_state_dict

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _tc_or_skip():
    from panonerf_b200 import _lib
    if not _lib.lib().pnb_tc_available():
        pytest.skip("not an sm_100 device")


def _inputs(S, R=None, seed=0):
    from panonerf_b200 import ops
    g = load_golden("panonerf_w256.npz")
    sd = golden_state_dict(g)
    rays, _ = golden_rays(g, DEV)
    if R is not None:
        rays = type(rays)(*[x[:R].contiguous() for x in rays])
    t, means, covs = ops.sample_cast(rays.origins, rays.directions, rays.radii, rays.near, rays.far, S)
    venc = ops.pos_enc(rays.viewdirs, 4)
    return sd, means, covs, venc


def _field(sd, means, covs, venc, S, fused, with_normals, grads=None, need_means=False):
    from panonerf_b200 import field
    old = os.environ.pop("PNB_NO_FUSED", None)
    if not fused:
        os.environ["PNB_NO_FUSED"] = "1"
    try:
        params = {k: v.clone().to(DEV).requires_grad_(grads is not None) for k, v in sd.items()}
        m = means.clone().requires_grad_(need_means)
        ctx = torch.enable_grad() if grads is not None else torch.no_grad()
        with ctx:
            raw_rgb, raw_den, n_raw = field.radiance_field(m, covs, venc, params, precision="bf16", samples_per_ray=S,
                                                           min_deg=0, max_deg=16, density_bias=-1.0, skip=4,
                                                           with_normals=with_normals)
            out = dict(raw_rgb=raw_rgb.detach(), raw_den=raw_den.detach(),
                       n_raw=None if n_raw is None else n_raw.detach())
            if grads is not None:
                g1, g2, g3 = grads
                loss = (raw_rgb * g1).sum() + (raw_den * g2).sum()
                if n_raw is not None:
                    loss = loss + (n_raw * g3).sum()
                loss.backward()
                out["grads"] = {k: p.grad.detach().double().flatten() for k, p in params.items()}
                if need_means:
                    out["d_means"] = m.grad.detach().double().flatten()
        return out
    finally:
        os.environ.pop("PNB_NO_FUSED", None)
        if old is not None:
            os.environ["PNB_NO_FUSED"] = old


def test_pack_blob_is_the_swizzled_bf16_image_of_the_weights():
    """Un-swizzle two tiles of the blob on the host and compare with the parameters, bit-exactly."""
    _tc_or_skip()
    from panonerf_b200 import field
    sd = golden_state_dict(load_golden("panonerf_w256.npz"))
    names = field.param_names(8, 1)
    params = [sd[n].to(DEV).contiguous() for n in names]
    pack = field.fused_pack(names, params)
    torch.cuda.synchronize()
    blob = pack["wblob"].cpu()

    def tile(off, rows):
        raw = blob[off:off + rows * 128].view(rows, 8, 16)          # [row][stored chunk][bytes]
        r = torch.arange(rows).view(rows, 1)
        j = torch.arange(8).view(1, 8)
        src = (j ^ (r & 7))                                           # logical chunk j lives at stored chunk j ^ (r&7)
        logical = torch.gather(raw, 1, src.view(rows, 8, 1).expand(rows, 8, 16))
        return logical.reshape(rows, 128).view(torch.bfloat16)       # [rows, 64]

    w0 = sd["layers.0.0.weight"]
    assert torch.equal(tile(0, 256).float(), w0[:, :64].bfloat16().float())
    t1 = tile(256 * 128, 256).float()
    assert torch.equal(t1[:, :32], w0[:, 64:96].bfloat16().float()) and float(t1[:, 32:].abs().max()) == 0.0
    # entry 2 of the blob: the bias of layer 0 as an un-swizzled [256 x 16] K-major tile (8-row core matrices of
    # 16-byte rows) whose row n holds three bf16 terms that sum to the fp32 bias exactly, followed by zeros and a
    # [128 x 16] tile of ones - the operands of the "bias" MMA step
    off = 2 * 256 * 128
    rows = blob[off:off + 4096].view(256, 16).view(torch.bfloat16).float()           # [256, 8]
    assert torch.equal(rows[:, :3].double().sum(-1).float(), sd["layers.0.0.bias"]) and float(rows[:, 3:].abs().max()) == 0
    assert float(blob[off + 4096:off + 8192].float().abs().max()) == 0.0
    assert torch.equal(blob[off + 8192:off + 12288].view(torch.bfloat16).float(), torch.ones(2048))
    w1 = sd["layers.1.0.weight"]
    assert torch.equal(tile(off + 12288 + 2 * 256 * 128, 256).float(), w1[:, 128:192].bfloat16().float())
    bb = pack["bblob"].cpu()
    assert torch.equal(bb[:256], sd["layers.0.0.bias"]) and torch.equal(bb[2048:2304], sd["extra_layer.bias"])
    assert torch.equal(bb[2304:2309], sd["density_layer.bias"]) and torch.equal(bb[2320:2323], sd["color_layer.bias"])
    assert torch.equal(bb[2336:2592], sd["density_layer.weight"][0])


@pytest.mark.parametrize("S,R,normals", [(64, None, False), (64, None, True), (10, 77, True), (7, 3, False)])
def test_fused_forward_matches_layered_path(S, R, normals):
    """Same bf16 operands, fp32 accumulation (the fused kernel may walk the k-blocks of every second tile backwards,
    so a bf16 rounding of a hidden activation can flip): outputs agree to about one bf16 ulp of their magnitude
    (ragged tile tails included)."""
    _tc_or_skip()
    sd, means, covs, venc = _inputs(S, R)
    a = _field(sd, means, covs, venc, S, True, normals)
    b = _field(sd, means, covs, venc, S, False, normals)
    assert_close(a["raw_rgb"], b["raw_rgb"], 1.5e-2, "raw_rgb", floor=float(b["raw_rgb"].abs().mean()))
    assert_close(a["raw_den"], b["raw_den"], 1.5e-2, "raw_den", floor=float(b["raw_den"].abs().mean()))
    if normals:
        na, nb = a["n_raw"].reshape(-1, 3), b["n_raw"].reshape(-1, 3)
        assert torch.isfinite(na).all()
        cos = torch.nn.functional.cosine_similarity(na.double(), nb.double(), dim=-1)
        big = nb.norm(dim=-1) > 1e-3 * float(nb.norm(dim=-1).mean())
        assert float((cos[big] > 0.999).float().mean()) > 0.99, float((cos[big] > 0.999).float().mean())
        assert_close(na.norm(dim=-1).mean(), nb.norm(dim=-1).mean(), 1e-2, "normal magnitude")


@pytest.mark.parametrize("need_means", [False, True])
def test_fused_forward_feeds_the_backward(need_means):
    """Activations / Jacobian rows written by the fused kernel's TMA stores drive the hand-written backward: every
    parameter gradient (incl. the adjoint of the Jacobian sweep) agrees with the layered path."""
    _tc_or_skip()
    S = 64
    sd, means, covs, venc = _inputs(S)
    R = means.shape[0]
    gen = torch.Generator().manual_seed(3)
    grads = (torch.randn(R, S, 3, generator=gen).to(DEV), torch.randn(R, S, 5, generator=gen).to(DEV),
             (torch.randn(R, S, 3, generator=gen) * 1e-2).to(DEV))
    a = _field(sd, means, covs, venc, S, True, True, grads, need_means)
    b = _field(sd, means, covs, venc, S, False, True, grads, need_means)
    for k in b["grads"]:
        ga, gb = a["grads"][k], b["grads"][k]
        cos = float((ga @ gb) / (ga.norm() * gb.norm() + 1e-30))
        assert cos > 0.9995, (k, cos)
        assert abs(float(ga.norm()) - float(gb.norm())) <= 0.01 * float(gb.norm()), k
    if need_means:
        ga, gb = a["d_means"], b["d_means"]
        assert float((ga @ gb) / (ga.norm() * gb.norm() + 1e-30)) > 0.999


def test_fused_kernel_is_deterministic():
    _tc_or_skip()
    sd, means, covs, venc = _inputs(64)
    a = _field(sd, means, covs, venc, 64, True, True)
    b = _field(sd, means, covs, venc, 64, True, True)
    for k in ("raw_rgb", "raw_den", "n_raw"):
        assert torch.equal(a[k], b[k]), k


def test_backward_kernel_small_batches_do_not_deadlock():
    """Regression: the epilogue-only first op of the backward program (no MMA) used to arrive on the activation
    barrier un-gated; with few, fast CTAs the barrier ran two phases ahead of the MMA thread (parity aliasing) and the
    kernel hung.  Exercised at the sizes that reproduced it (16 CTAs x one tile pair, all tiles full)."""
    _tc_or_skip()
    from panonerf_b200 import field, ops
    sd = golden_state_dict(load_golden("panonerf_w256.npz"))
    names = field.param_names(8, 1)
    params = [sd[n].to(DEV).contiguous() for n in names]
    pack = field.fused_pack(names, params)
    for M in (4096, 256, 21):
        g = torch.Generator(device=DEV).manual_seed(M)
        means = torch.rand(M, 3, device=DEV, generator=g) * 4 - 2
        covs = torch.rand(M, 3, device=DEV, generator=g) * 1e-3
        enc = torch.empty(M, 96, device=DEV, dtype=torch.bfloat16)
        ops.ipe_into(means, covs, 0, 16, enc)
        vb = torch.randn((M + 63) // 64, 128, device=DEV, generator=g)
        masks = field.fused_masks(M, torch.device(DEV, 0), True)
        acts = torch.empty(18, M, 256, device=DEV, dtype=torch.bfloat16)
        field.fused_forward(enc, vb, 64, 5, pack, acts, None, masks, True)
        d_rgb = torch.randn(M, 3, device=DEV, generator=g)
        d_den = torch.randn(M, 5, device=DEV, generator=g)
        for _ in range(100):
            dz = field.fused_backward(M, 5, pack, d_rgb, d_den, masks, None)
        torch.cuda.synchronize()
        # (plane 0 = dz of the 128-wide view layer: its upper 128 columns are never written)
        assert torch.isfinite(dz[1:].float()).all() and torch.isfinite(dz[0][:, :128].float()).all()


def test_wgrad_batch_matches_matmul():
    """Batched weight gradients + fused bias column sums against fp32 matmuls of the same bf16 operands, including two
    jobs that accumulate into one destination and a narrow (K=16) job."""
    _tc_or_skip()
    from panonerf_b200 import field
    M = 5000
    g = torch.Generator(device=DEV).manual_seed(1)
    planes = (torch.randn(4, M, 256, device=DEV, generator=g)).bfloat16()
    x96 = torch.randn(M, 96, device=DEV, generator=g).bfloat16()
    wb = field.WgradBatch(M, torch.device(DEV, 0))
    G0 = torch.zeros(256, 256, device=DEV)
    G1 = torch.zeros(128, 352, device=DEV)
    G2 = torch.zeros(256, 16, device=DEV)
    b0, b2 = torch.zeros(256, device=DEV), torch.zeros(256, device=DEV)
    wb.add(planes, 0, 256, planes, 1, 256, G0, b0)
    wb.add(planes, 2, 256, planes, 3, 256, G0)                     # same destination
    wb.add(planes, 1, 128, planes, 2, 256, G1[:, :256])
    wb.add(planes, 1, 128, x96, 0, 96, G1[:, 256:])
    wb.add(planes, 3, 256, x96, 0, 16, G2, b2)
    wb.launch()
    torch.cuda.synchronize()
    P = planes.float()
    ref0 = P[0].t() @ P[1] + P[2].t() @ P[3]
    assert_close(G0, ref0, 2e-5, "dW (two jobs, one destination)")
    assert_close(G1[:, :256], P[1][:, :128].t() @ P[2], 2e-5, "dW (Nw=128)")
    assert_close(G1[:, 256:], P[1][:, :128].t() @ x96.float(), 2e-5, "dW (Kw=96, strided destination)")
    assert_close(G2, P[3].t() @ x96.float()[:, :16], 2e-5, "dW (Kw=16)")
    assert_close(b0, P[0].sum(0), 2e-5, "bias column sums")
    assert_close(b2, P[3].sum(0), 2e-5, "bias column sums (narrow job)")


def test_head_grad_padding_and_group_sum():
    _tc_or_skip()
    from panonerf_b200 import field
    M, C, S = 6400, 5, 64
    g = torch.Generator(device=DEV).manual_seed(2)
    x = torch.randn(M, C, device=DEV, generator=g)
    cs = torch.zeros(C, device=DEV)
    pad = field._pad_head_grad(x, cs)
    assert torch.equal(pad[:, :C].float(), x.bfloat16().float()) and float(pad[:, C:].float().abs().max()) == 0.0
    assert_close(cs, x.sum(0), 1e-5, "head bias gradient")
    z = torch.randn(M, 256, device=DEV, generator=g).bfloat16()
    out = field._group_sum(z[:, :128], S)
    ref = z[:, :128].float().view(M // S, S, 128).sum(1)
    assert_close(out, ref, 1e-6, "group_sum (bf16x2 path)")


@pytest.mark.parametrize("S,R,with_normals", [(64, 16, False), (64, 16, True), (10, 13, False), (64, 700, True)])
def test_in_kernel_ipe_is_bit_identical_to_the_two_kernel_path(S, R, with_normals):
    """Inference forward with PNB_FUSED_IPE=1 (opt-in, see field.py): the encoder warps of the fused kernel evaluate
    the IPE in the kernel (no [M,96] array);
    the raw outputs and the density-gradient normals must equal those of pnb_ipe_fwd + pnb_mlp_fused_fwd bit for bit
    (same arithmetic, same bf16 rounding of the features).  Sizes cover a partial tile, an odd number of tiles (the
    phantom tile of the last pair) and several pairs per CTA-free grid."""
    _tc_or_skip()
    if R > 16:
        g = load_golden("panonerf_c2s.npz")
        sd = golden_state_dict(g)
        rays, _ = golden_rays(g, DEV)
        from panonerf_b200 import ops
        rays = type(rays)(*[x[:R].contiguous() for x in rays])
        _, means, covs = ops.sample_cast(rays.origins, rays.directions, rays.radii, rays.near, rays.far, S)
        venc = ops.pos_enc(rays.viewdirs, 4)
    else:
        sd, means, covs, venc = _inputs(S, R)
    outs = []
    for in_kernel in (False, True):
        if in_kernel:
            os.environ["PNB_FUSED_IPE"] = "1"
        try:
            outs.append(_field(sd, means, covs, venc, S, True, with_normals))
        finally:
            os.environ.pop("PNB_FUSED_IPE", None)
    a, b = outs
    assert torch.equal(a["raw_rgb"], b["raw_rgb"]) and torch.equal(a["raw_den"], b["raw_den"])
    if with_normals:
        assert torch.equal(a["n_raw"], b["n_raw"])
    assert torch.isfinite(b["raw_rgb"]).all()


def test_env_view_term_indexed_modulo_directions():
    """Env rays share their D directions: passing the D view encodings with venc_mod = D must equal passing the
    expanded [R*D, 27] encodings (models/pano_mip_nerf.py:337-341), in inference (in-kernel IPE) and in training."""
    _tc_or_skip()
    from panonerf_b200 import field
    sd, means, covs, venc = _inputs(10, 16)            # 16 "env rays" of 10 samples: 8 points x D = 2 directions
    D = 2
    venc_d = venc[:D].contiguous()
    venc_full = venc_d[None].expand(8, D, -1).reshape(16, -1).contiguous()
    params = {k: v.clone().to(DEV) for k, v in sd.items()}
    kw = dict(precision="bf16", samples_per_ray=10, min_deg=0, max_deg=16, density_bias=-1.0, skip=4, with_normals=False)
    with torch.no_grad():
        a = field.radiance_field(means, covs, venc_full, params, **kw)
        b = field.radiance_field(means, covs, venc_d, params, venc_mod=D, **kw)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    gen = torch.Generator().manual_seed(0)
    g1 = torch.randn(16, 10, 3, generator=gen).to(DEV)
    grads = []
    for ve, mod in ((venc_full, 0), (venc_d, D)):
        ps = {k: v.clone().to(DEV).requires_grad_() for k, v in sd.items()}
        rgb, den, _ = field.radiance_field(means, covs, ve, ps, venc_mod=mod, **kw)
        (rgb * g1).sum().backward()
        grads.append({k: p.grad.clone() for k, p in ps.items()})
    for k in grads[0]:      # (not bit-equal run to run: the head-bias sums use float atomics)
        assert_close(grads[1][k], grads[0][k], 1e-5, k, floor=max(float(grads[0][k].abs().max()), 1e-12))
