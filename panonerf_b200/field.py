"""The radiance/density field: IPE -> 8x256 trunk (+skip) -> density / extra / view / colour heads
(models/pano_mip_nerf.py:78-114 + 235-280), with the density-gradient normals (pano_mip_nerf.py:299-302) computed
by an explicit Jacobian sweep instead of vmap(jacrev).  One autograd Function owns every activation buffer and
implements the backward by hand (dgrad + wgrad per layer, plus the adjoint of the Jacobian sweep, i.e. the
"double backward" the reference gets from functorch).

Two GEMM back-ends share the orchestration:
  * `tc`   bf16 operands on tcgen05 tensor cores (pnb_linear_tc / pnb_wgrad_tc), fp32 accumulation in TMEM;
  * `f32`  fp32 FFMA (pnb_gemm_f32) - the parity mode that matches the fp32 reference to ~1e-6.
"""
from __future__ import annotations

import ctypes
import os
import weakref
from typing import Dict, List, Optional

import torch

from . import _lib
from ._lib import EPI_ACCUM, EPI_BIAS, EPI_MASK, EPI_RELU, PNB_BF16, PNB_F32, check
from . import ops
from .ops import _p, _stream, _amp_fwd, _amp_bwd

PARAM_ORDER_TRUNK = "layers.{}.0"


def param_names(depth: int, depth_cond: int) -> List[str]:
    """State-dict order of models/pano_mip_nerf.py:54-76."""
    names = []
    for i in range(depth):
        names += [f"layers.{i}.0.weight", f"layers.{i}.0.bias"]
    names += ["density_layer.weight", "density_layer.bias", "extra_layer.weight", "extra_layer.bias"]
    for i in range(depth_cond):
        names += [f"view_layers.{i}.0.weight", f"view_layers.{i}.0.bias"]
    names += ["color_layer.weight", "color_layer.bias"]
    return names


# ----------------------------------------------------------------------------------------------------------------
# GEMM back-ends
# ----------------------------------------------------------------------------------------------------------------
def _ld(t: torch.Tensor) -> int:
    assert t.dim() == 2 and t.stride(1) == 1, "matrix operands must be row-major views"
    return t.stride(0)


class _F32Backend:
    """fp32 FFMA GEMMs. Weights are used in place (views of the fp32 parameters)."""
    dtype = torch.float32
    code = PNB_F32
    name = "f32"

    def __init__(self, params: Dict[str, torch.Tensor]):
        self.params = params

    def linear(self, a, w, out, bias=None, relu=False, mask=None, row_bias=None, group=0, accum=False):
        flags = (EPI_BIAS if bias is not None else 0) | (EPI_RELU if relu else 0) | \
                (EPI_MASK if mask is not None else 0) | (EPI_ACCUM if accum else 0)
        m, k = a.shape
        n = w.shape[0]
        with torch.cuda.device(a.device):
            check(_lib.lib().pnb_gemm_f32(0, m, n, k, _p(a), _ld(a), _p(w), _ld(w), _p(out), _ld(out), _p(bias),
                                          _p(row_bias), group, _p(mask), _ld(mask) if mask is not None else 0, flags,
                                          _stream()), "gemm_f32(NT)")

    def dgrad(self, dz, w, out, mask=None, accum=False, colsum_out=None):
        """out[M,K] = dz[M,N] @ w[N,K]   (colsum_out += column sums of out: the upstream layer's bias gradient)"""
        flags = (EPI_MASK if mask is not None else 0) | (EPI_ACCUM if accum else 0)
        m, n = dz.shape
        k = w.shape[1]
        with torch.cuda.device(dz.device):
            check(_lib.lib().pnb_gemm_f32(1, m, k, n, _p(dz), _ld(dz), _p(w), _ld(w), _p(out), _ld(out), None, None, 0,
                                          _p(mask), _ld(mask) if mask is not None else 0, flags, _stream()),
                  "gemm_f32(NN)")
        if colsum_out is not None:
            _colsum(out, colsum_out)

    dgrad_small = dgrad

    def wgrad(self, dz, x, dw):
        """dw[N,K] += dz[M,N]^T @ x[M,K]"""
        m, n = dz.shape
        k = x.shape[1]
        with torch.cuda.device(dz.device):
            check(_lib.lib().pnb_gemm_f32(2, n, k, m, _p(dz), _ld(dz), _p(x), _ld(x), _p(dw), _ld(dw), None, None, 0,
                                          None, 0, 0, _stream()), "gemm_f32(TN)")

    wgrad_small = wgrad

    def w(self, name, c0=None, c1=None):
        t = self.params[name]
        return t if c0 is None else t[:, c0:c1]


# bench.py sets PROFILE = {} around its timed region: every tensor-core GEMM launch is then bracketed by CUDA events
# on the launching stream and its algorithmic bytes / FLOPs are tallied per kernel (roofline evidence).
PROFILE = None


class _prof:
    def __init__(self, name, nbytes, flops):
        self.on = PROFILE is not None
        if self.on:
            self.rec = PROFILE.setdefault(name, dict(events=[], bytes=0, flops=0))
            self.rec["bytes"] += nbytes
            self.rec["flops"] += flops

    def __enter__(self):
        if self.on:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *a):
        if self.on:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            self.rec["events"].append((self.e0, e1))


_pack_cache: Dict[int, dict] = {}
_ws_cache: Dict[str, torch.Tensor] = {}
_pack_epoch = [0]


def invalidate_packs():
    """Call after parameters were updated through raw pointers (the fused Adam kernel does not bump
    torch's version counters)."""
    _pack_epoch[0] += 1


class _PackedWeight:
    """bf16 copies of one fp32 weight (or a column slice of it): `fwd` [N,K] for A@W^T, `t` [K,N] for dZ@W.
    Inner dimensions are zero-padded to a multiple of 64 so that every TMA row is 16-byte aligned."""

    def __init__(self, w32: torch.Tensor):
        n, k = w32.shape
        self.shape = (n, k)
        kp, np_ = (k + 63) // 64 * 64, (n + 63) // 64 * 64
        self.fwd = torch.zeros(n, kp, device=w32.device, dtype=torch.bfloat16)
        self.fwd[:, :k] = w32
        self.t = torch.zeros(k, np_, device=w32.device, dtype=torch.bfloat16)
        self.t[:, :n] = w32.t()


class _TCBackend:
    """bf16 tcgen05 GEMMs.  Weight packs are cached per parameter version (re-packed after each optimiser step)."""
    dtype = torch.bfloat16
    code = PNB_BF16
    name = "tc"

    def __init__(self, params: Dict[str, torch.Tensor]):
        self.params = params
        self._packs: Dict[tuple, _PackedWeight] = {}
        w0, wv = params["layers.0.0.weight"], params["view_layers.0.0.weight"]
        if w0.shape[0] != 256 or wv.shape[0] != 128:
            raise NotImplementedError("the tcgen05 path is specialised for net_width=256 / net_width_condition=128 "
                                      "(configs/*.yaml); use precision='fp32' for other widths")

    def w(self, name, c0=None, c1=None):
        key = (name, c0, c1)
        if key not in self._packs:
            p = self.params[name]
            ck = (p.data_ptr(), p._version, _pack_epoch[0], c0, c1)
            slot = _pack_cache.get(id(p))
            if slot is not None and slot["ref"]() is not p:      # id() re-used by a different (newer) tensor
                slot = None
            if slot is None:
                slot = _pack_cache[id(p)] = {"ref": weakref.ref(p), "packs": {}}
                if len(_pack_cache) > 512:                       # drop entries of dead tensors
                    for k in [k for k, v in _pack_cache.items() if v["ref"]() is None]:
                        del _pack_cache[k]
            hit = slot["packs"].get((c0, c1))
            if hit is not None and hit[0] == ck:
                self._packs[key] = hit[1]
            else:
                with torch.no_grad():
                    pk = _PackedWeight(p.detach() if c0 is None else p.detach()[:, c0:c1])
                slot["packs"][(c0, c1)] = (ck, pk)
                self._packs[key] = pk
        return self._packs[key]

    @staticmethod
    def _workspace(dev) -> torch.Tensor:
        key = str(dev)
        if key not in _ws_cache:
            nbytes = int(_lib.lib().pnb_wgrad_tc_workspace(256, 256))
            _ws_cache[key] = torch.empty(nbytes // 4, device=dev, dtype=torch.float32)
        return _ws_cache[key]

    def linear(self, a, w: _PackedWeight, out, bias=None, relu=False, mask=None, row_bias=None, group=0, accum=False):
        flags = (EPI_BIAS if bias is not None else 0) | (EPI_RELU if relu else 0) | \
                (EPI_MASK if mask is not None else 0) | (EPI_ACCUM if accum else 0)
        m, k = a.shape
        n = w.shape[0]
        assert k == w.shape[1], (a.shape, w.shape)
        k16 = (k + 15) // 16 * 16
        nbytes = m * k * 2 + m * n * out.element_size() * (2 if accum else 1) + (m * n * 2 if mask is not None else 0)
        with torch.cuda.device(a.device), _prof("linear_tc", nbytes, 2 * m * n * k):
            check(_lib.lib().pnb_linear_tc(m, n, k16, _p(a), _ld(a), _p(w.fwd), _ld(w.fwd), _p(out), _ld(out),
                                           ops.dt_code(out.dtype), _p(bias), _p(row_bias), group, _p(mask),
                                           _ld(mask) if mask is not None else 0, flags, None, _stream()), "linear_tc")

    def dgrad(self, dz, w: _PackedWeight, out, mask=None, accum=False, colsum_out=None):
        """out[M,K] = dz[M,N] @ W[N,K]  ==  linear with the pre-transposed pack W^T[K,N].  colsum_out (fp32 [K]) is
        incremented by the column sums of `out` inside the epilogue (bias gradient of the upstream layer)."""
        flags = (EPI_MASK if mask is not None else 0) | (EPI_ACCUM if accum else 0)
        m, n = dz.shape
        k = w.shape[1]
        n16 = (n + 15) // 16 * 16
        fused = colsum_out is not None and out.dtype == torch.bfloat16 and k % 64 == 0
        nbytes = m * n * 2 + m * k * out.element_size() * (2 if accum else 1) + (m * k * 2 if mask is not None else 0)
        with torch.cuda.device(dz.device), _prof("linear_tc", nbytes, 2 * m * n * k):
            check(_lib.lib().pnb_linear_tc(m, k, n16, _p(dz), _ld(dz), _p(w.t), _ld(w.t), _p(out), _ld(out),
                                           ops.dt_code(out.dtype), None, None, 0, _p(mask),
                                           _ld(mask) if mask is not None else 0, flags,
                                           _p(colsum_out) if fused else None, _stream()), "linear_tc(dgrad)")
        if colsum_out is not None and not fused:
            _colsum(out, colsum_out)

    def _pad64(self, x32: torch.Tensor) -> torch.Tensor:
        """fp32 [M,C] head gradient -> bf16 [M,64] zero-padded (TMA rows must be 16-byte aligned)."""
        m, c = x32.shape
        buf = torch.zeros(m, 64, device=x32.device, dtype=torch.bfloat16)
        with torch.cuda.device(x32.device):
            check(_lib.lib().pnb_convert(m, c, _p(x32), _ld(x32), PNB_F32, _p(buf), 64, PNB_BF16, _stream()), "convert")
        return buf

    def dgrad_small(self, dz32, w: _PackedWeight, out, mask=None, accum=False, colsum_out=None):
        pad = self._pad64(dz32)
        # K=16 slice of the padded buffer (ld stays 64)
        self.dgrad(pad[:, :16], w, out, mask=mask, accum=accum, colsum_out=colsum_out)
        return pad

    def wgrad(self, dz, x, dw):
        m, n = dz.shape
        k = x.shape[1]
        with torch.cuda.device(dz.device), _prof("wgrad_tc", m * (n + k) * 2, 2 * m * n * k):
            check(_lib.lib().pnb_wgrad_tc(m, n, (k + 15) // 16 * 16, _p(dz), _ld(dz), _p(x), _ld(x), _p(dw), _ld(dw),
                                          _p(self._workspace(dz.device)), _stream()), "wgrad_tc")

    def wgrad_small(self, dz32, x, dw, pad=None):
        """dw[C,K] += dz32[M,C]^T x[M,K] with C < 16: run the transposed product x^T dz (x plays the 128/256-wide
        operand) into a scratch [K,64] and add its first C columns, transposed."""
        pad = self._pad64(dz32) if pad is None else pad
        c, k = dw.shape
        tmp = torch.zeros(k, 64, device=x.device, dtype=torch.float32)
        self.wgrad(x, pad, tmp)
        dw.add_(tmp[:, :c].t())


class _SimtBf16Backend(_TCBackend):
    """CUDA-core twin of the tensor-core back-end: identical bf16 buffers, identical weight packs, identical
    epilogues - only the GEMM engine differs (FFMA instead of tcgen05).  Exists so that the tests can check the
    tcgen05 forward/backward on the same bf16-rounded data (precision='bf16_simt')."""
    name = "simt_bf16"

    def linear(self, a, w: _PackedWeight, out, bias=None, relu=False, mask=None, row_bias=None, group=0, accum=False):
        flags = (EPI_BIAS if bias is not None else 0) | (EPI_RELU if relu else 0) | \
                (EPI_MASK if mask is not None else 0) | (EPI_ACCUM if accum else 0)
        m, k = a.shape
        n = w.shape[0]
        with torch.cuda.device(a.device):
            check(_lib.lib().pnb_gemm_bf16_simt(0, m, n, k, _p(a), _ld(a), _p(w.fwd), _ld(w.fwd), _p(out), _ld(out),
                                                ops.dt_code(out.dtype), _p(bias), _p(row_bias), group, _p(mask),
                                                _ld(mask) if mask is not None else 0, flags, _stream()),
                  "gemm_bf16_simt(NT)")

    def dgrad(self, dz, w: _PackedWeight, out, mask=None, accum=False, colsum_out=None):
        flags = (EPI_MASK if mask is not None else 0) | (EPI_ACCUM if accum else 0)
        m, n = dz.shape
        k = w.shape[1]
        with torch.cuda.device(dz.device):
            check(_lib.lib().pnb_gemm_bf16_simt(0, m, k, n, _p(dz), _ld(dz), _p(w.t), _ld(w.t), _p(out), _ld(out),
                                                ops.dt_code(out.dtype), None, None, 0, _p(mask),
                                                _ld(mask) if mask is not None else 0, flags, _stream()),
                  "gemm_bf16_simt(dgrad)")
        if colsum_out is not None:
            _colsum(out, colsum_out)

    def wgrad(self, dz, x, dw):
        m, n = dz.shape
        k = x.shape[1]
        with torch.cuda.device(dz.device):
            check(_lib.lib().pnb_gemm_bf16_simt(2, n, k, m, _p(dz), _ld(dz), _p(x), _ld(x), _p(dw), _ld(dw), PNB_F32,
                                                None, None, 0, None, 0, 0, _stream()), "gemm_bf16_simt(TN)")



# ----------------------------------------------------------------------------------------------------------------
# fused MLP (csrc/mlp_fused.cu): weight blob cache and launcher
# ----------------------------------------------------------------------------------------------------------------
FUSED_TOPOLOGY = dict(depth=8, skip=4, width=256, xyz_dim=96, cond_width=128, view_in=283)
_fused_cache: Dict[tuple, dict] = {}


def fused_enabled() -> bool:
    return not os.environ.get("PNB_NO_FUSED")


def fused_supported(cfg, P) -> bool:
    t = FUSED_TOPOLOGY
    wv = P["view_layers.0.0.weight"]
    return (cfg["depth"] == t["depth"] and cfg["skip"] == t["skip"] and cfg["width"] == t["width"]
            and cfg["xyz_dim"] == t["xyz_dim"] and tuple(wv.shape) == (t["cond_width"], t["view_in"])
            and P["density_layer.weight"].shape[0] <= 16 and P["color_layer.weight"].shape[0] == 3
            and len([n for n in cfg["names"] if n.startswith("view_layers.") and n.endswith(".weight")]) == 1)


def fused_pack(names: List[str], params) -> dict:
    """bf16 tile blob + fp32 bias blob of the current parameter values (re-packed when any parameter changed)."""
    key = tuple(id(p) for p in params)
    ck = tuple((p.data_ptr(), p._version) for p in params) + (_pack_epoch[0],)
    slot = _fused_cache.get(key)
    if slot is not None and slot["ck"] == ck and all(r() is p for r, p in zip(slot["refs"], params)):
        return slot
    dev = params[0].device
    lib = _lib.lib()
    if slot is None or slot["wblob"].device != dev:
        if len(_fused_cache) > 64:
            _fused_cache.clear()
        slot = dict(wblob=torch.empty(int(lib.pnb_mlp_fused_wblob_bytes()), device=dev, dtype=torch.uint8),
                    bblob=torch.empty(int(lib.pnb_mlp_fused_bblob_floats()), device=dev, dtype=torch.float32))
        _fused_cache[key] = slot
    order = param_names(8, 1)
    by_name = dict(zip(names, params))
    ptrs = (ctypes.c_void_p * 24)(*[by_name[n].data_ptr() for n in order])
    for n in order:
        t = by_name[n]
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError(f"panonerf_b200: parameter {n} must be contiguous fp32")
    C = by_name["density_layer.weight"].shape[0]
    with torch.cuda.device(dev):
        check(lib.pnb_mlp_fused_pack(ptrs, C, _p(slot["wblob"]), _p(slot["bblob"]), _stream()), "mlp_fused_pack")
    slot["ck"] = ck
    slot["refs"] = [weakref.ref(p) for p in params]
    return slot


_mask_scratch: Dict[str, torch.Tensor] = {}


def fused_masks(M: int, dev, per_tile: bool) -> torch.Tensor:
    """Sign bit-plane buffer of the fused kernels: one record per 128-sample tile when the backward kernels will
    read it, otherwise a per-device scratch that only the in-kernel Jacobian sweep reads back."""
    words = int(_lib.lib().pnb_mlp_fused_mask_words(M, 1 if per_tile else 0))
    if per_tile:
        return torch.empty(words, device=dev, dtype=torch.int32)
    key = str(dev)
    if key not in _mask_scratch:
        _mask_scratch[key] = torch.empty(words, device=dev, dtype=torch.int32)
    return _mask_scratch[key]


def fused_forward(enc, vb, S, C, pack, acts, g_enc, masks=None, masks_per_tile=False, vb_mod=0):
    """Launch the fused kernel on bf16 IPE features `enc` [M,96]; returns (raw_den, raw_rgb)."""
    M = enc.shape[0]
    dev = enc.device
    raw_den = torch.empty(M, C, device=dev, dtype=torch.float32)
    raw_rgb = torch.empty(M, 3, device=dev, dtype=torch.float32)
    if g_enc is not None and masks is None:
        masks = fused_masks(M, dev, False)
    flops = M * (2 * (96 * 256 + 6 * 256 * 256 + 352 * 256 + 256 * C + 256 * 256 + 256 * 128 + 128 * 3)
                 + (2 * (6 * 256 * 256 + 2 * 96 * 256) if g_enc is not None else 0))
    nbytes = M * (192 + 4 * C + 12 + (384 if g_enc is not None else 0)
                  + (512 * (10 + (8 if g_enc is not None else 0)) if acts is not None else 0))
    with torch.cuda.device(dev), _prof("mlp_fused", nbytes, flops):
        check(_lib.lib().pnb_mlp_fused_fwd(M, S, C, _p(enc), _ld(enc), _p(pack["wblob"]), _p(pack["bblob"]), _p(vb),
                                           int(vb_mod), _p(raw_den), _p(raw_rgb), _p(acts), _p(g_enc), _p(masks),
                                           1 if masks_per_tile else 0, _stream()), "mlp_fused_fwd")
    return raw_den, raw_rgb


_enc_scratch: Dict[str, torch.Tensor] = {}


def fused_forward_ipe(means, covs, min_deg, vb, vb_mod, S, C, pack, g_enc):
    """Inference forward with the IPE computed inside the fused kernel (csrc/mlp_fused.cu, encoder warps): no [M,96]
    encoding array, no pnb_ipe_fwd launch.  Returns (raw_den, raw_rgb)."""
    M = means.shape[0]
    dev = means.device
    raw_den = torch.empty(M, C, device=dev, dtype=torch.float32)
    raw_rgb = torch.empty(M, 3, device=dev, dtype=torch.float32)
    masks = fused_masks(M, dev, False) if g_enc is not None else None
    key = str(dev)
    if key not in _enc_scratch:
        _enc_scratch[key] = torch.empty(int(_lib.lib().pnb_mlp_fused_scratch_bytes()), device=dev, dtype=torch.uint8)
    flops = M * (2 * (96 * 256 + 6 * 256 * 256 + 352 * 256 + 256 * C + 256 * 256 + 256 * 128 + 128 * 3)
                 + (2 * (6 * 256 * 256 + 2 * 96 * 256) if g_enc is not None else 0))
    nbytes = M * (24 + 4 * C + 12 + (384 if g_enc is not None else 0))
    with torch.cuda.device(dev), _prof("mlp_fused", nbytes, flops):
        check(_lib.lib().pnb_mlp_fused_fwd_ipe(M, S, C, _p(means), _p(covs), min_deg, _p(pack["wblob"]),
                                               _p(pack["bblob"]), _p(vb), vb_mod, _p(raw_den), _p(raw_rgb), _p(g_enc),
                                               _p(masks), _p(_enc_scratch[key]), _stream()), "mlp_fused_fwd_ipe")
    return raw_den, raw_rgb


def fused_backward(M, C, pack, d_rgb, d_den, masks, d_enc):
    """dgrad chain of the backward pass in one kernel -> dz planes bf16 [10, M, 256] (include/panonerf_b200.h)."""
    dev = d_rgb.device
    planes = int(_lib.lib().pnb_mlp_fused_bwd_planes())
    dz = torch.empty(planes, M, 256, device=dev, dtype=torch.bfloat16)
    flops = M * 2 * (128 * 256 + 256 * 256 + 7 * 256 * 256 + 2 * 96 * 256)
    nbytes = M * (12 + 4 * C + 288 + 512 * planes + (768 if d_enc is not None else 0))
    with torch.cuda.device(dev), _prof("mlp_fused_bwd", nbytes, flops):
        check(_lib.lib().pnb_mlp_fused_bwd(M, C, _p(pack["wblob"]), _p(pack["bblob"]), _p(d_rgb), _p(d_den), _p(masks),
                                           _p(dz), _p(d_enc), _stream()), "mlp_fused_bwd")
    return dz


def fused_jadj(u, pack, masks):
    """Adjoint (forward-mode) sweep of the density Jacobian in one kernel -> q planes bf16 [8, M, 256]."""
    M = u.shape[0]
    dev = u.device
    planes = int(_lib.lib().pnb_mlp_fused_adj_planes())
    q = torch.empty(planes, M, 256, device=dev, dtype=torch.bfloat16)
    flops = M * 2 * (96 * 256 + 6 * 256 * 256 + 352 * 256)
    nbytes = M * (192 + 256 + 512 * planes)
    with torch.cuda.device(dev), _prof("mlp_fused_jadj", nbytes, flops):
        check(_lib.lib().pnb_mlp_fused_jadj(M, _p(u), _ld(u), _p(pack["wblob"]), _p(masks), _p(q), _stream()),
              "mlp_fused_jadj")
    return q


class WgradBatch:
    """Collects the weight-gradient GEMMs (and bias column sums) of one backward pass and runs them as ONE launch
    of pnb_wgrad_batch (csrc/wgrad_batch.cu)."""

    def __init__(self, M: int, dev):
        self.M, self.dev = M, dev
        self.maps, self.map_ids, self.jobs, self.keep = [], {}, [], []

    def _map(self, t: torch.Tensor) -> int:
        assert t.dtype == torch.bfloat16 and t.stride(-1) == 1 and t.shape[-2] == self.M
        key = (t.data_ptr(), tuple(t.shape), t.stride(-2))
        if key not in self.map_ids:
            planes = t.shape[0] if t.dim() == 3 else 1
            if t.dim() == 3:
                assert t.stride(0) == t.shape[1] * t.stride(1)
            self.map_ids[key] = len(self.maps)
            self.maps.append((t, planes, t.stride(-2), t.shape[-1]))
        return self.map_ids[key]

    def add(self, z, zplane, nw, x, xplane, kw, dW, db=None):
        """dW[nw,kw] += z[zplane][:, :nw]^T x[xplane][:, :kw];  db[nw] += column sums of z[zplane][:, :nw]."""
        assert dW.dtype == torch.float32 and dW.stride(1) == 1 and tuple(dW.shape) == (nw, kw)
        self.jobs.append((self._map(z), zplane, self._map(x), xplane, nw, kw, dW.stride(0), dW, db))

    def launch(self):
        if not self.jobs:
            return
        lib = _lib.lib()
        key = "wb" + str(self.dev)
        if key not in _ws_cache:
            _ws_cache[key] = torch.empty(int(lib.pnb_wgrad_batch_workspace()) // 4, device=self.dev,
                                         dtype=torch.float32)
        nm, nj = len(self.maps), len(self.jobs)
        base = (ctypes.c_void_p * nm)(*[m[0].data_ptr() for m in self.maps])
        desc = (ctypes.c_longlong * (3 * nm))(*[v for m in self.maps for v in (m[1], m[2], m[3])])
        jobs = (ctypes.c_longlong * (8 * nj))(*[v for j in self.jobs
                                                for v in (j[0], j[1], j[2], j[3], j[4], j[5], j[6],
                                                          1 if j[8] is not None else 0)])
        dW = (ctypes.c_void_p * nj)(*[j[7].data_ptr() for j in self.jobs])
        db = (ctypes.c_void_p * nj)(*[(j[8].data_ptr() if j[8] is not None else None) for j in self.jobs])
        nbytes = sum(self.M * 2 * (j[4] + j[5]) for j in self.jobs)
        flops = sum(2 * self.M * j[4] * j[5] for j in self.jobs)
        with torch.cuda.device(self.dev), _prof("wgrad_batch", nbytes, flops):
            check(lib.pnb_wgrad_batch(self.M, nm, base, desc, nj, jobs, dW, db, _p(_ws_cache[key]), _stream()),
                  "wgrad_batch")


def make_backend(precision: str, params: Dict[str, torch.Tensor]):
    if precision == "fp32":
        return _F32Backend(params)
    if precision == "bf16":
        return _TCBackend(params)
    if precision == "bf16_simt":
        return _SimtBf16Backend(params)
    raise ValueError(f"precision must be 'bf16', 'bf16_simt' or 'fp32', got {precision!r}")


def _mask_scale(src, w_row, out):
    m, n = src.shape
    with torch.cuda.device(src.device):
        check(_lib.lib().pnb_mask_scale(m, n, _p(src), _ld(src), _p(w_row), None, _p(out), _ld(out),
                                        ops.dt_code(out.dtype), _stream()), "mask_scale")


def _pad_head_grad(x32, colsum_out):
    """fp32 [M,C] head gradient -> bf16 [M,64] zero-padded tensor-core operand; colsum_out += its column sums."""
    m, c = x32.shape
    buf = torch.empty(m, 64, device=x32.device, dtype=torch.bfloat16)
    with torch.cuda.device(x32.device):
        check(_lib.lib().pnb_pad_head_grad(m, c, _p(x32), _p(buf), _p(colsum_out), _stream()), "pad_head_grad")
    return buf


def _colsum(x, out):
    m, n = x.shape
    with torch.cuda.device(x.device):
        check(_lib.lib().pnb_colsum(m, n, _p(x), _ld(x), ops.dt_code(x.dtype), _p(out), _stream()), "colsum")


def _group_sum(x, group):
    m, n = x.shape
    out = torch.empty(m // group, n, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        check(_lib.lib().pnb_group_sum(m, n, group, _p(x), _ld(x), ops.dt_code(x.dtype), _p(out), _stream()), "group_sum")
    return out


# ----------------------------------------------------------------------------------------------------------------
# the autograd Function
# ----------------------------------------------------------------------------------------------------------------
def _use_fused(cfg, P) -> bool:
    return (cfg["precision"] == "bf16" and fused_enabled() and fused_supported(cfg, P)
            and not (cfg["with_normals"] and cfg.get("jac_precision") == "fp32"))


def _gaussian_grads(ctx, means, covs, cfg, d_enc, d_v):
    """Gradients w.r.t. the Gaussians from dL/d enc (`d_enc`, [M,6L]): the means' (env branch: surface point ->
    distance; stop_resample_grad=False: the resampled fence-posts) and, only with stop_resample_grad=False, the
    variances'.  When the normals feed the loss (`d_v` = dL/d v given) their explicit dependence on the Gaussian
    through the encoding's Jacobian is added (pnb_ipe_cov_hess; `ctx.g_keep` = d sigma / d enc of the forward)."""
    if d_enc is None:
        return None, None
    d_means = ops.ipe_vjp(means, covs, cfg["min_deg"], cfg["max_deg"], d_enc)
    d_covs = None
    if ctx.need_covs:
        d_covs = torch.empty_like(d_means)
        ops.ipe_cov_hess(means, covs, cfg["min_deg"], cfg["max_deg"], d_enc=d_enc, d_covs=d_covs)
    if ctx.g_keep is not None and d_v is not None:
        h = ctx.g_keep if ctx.g_keep.dtype == torch.float32 else ctx.g_keep.float()
        ops.ipe_cov_hess(means, covs, cfg["min_deg"], cfg["max_deg"], h_enc=h, d_v=d_v, d_means=d_means,
                         d_covs=d_covs, accumulate=True)
    ctx.g_keep = None
    return (d_means.view(ctx.means_shape) if ctx.need_means else None,
            d_covs.view(ctx.means_shape) if d_covs is not None else None)


class _Field(torch.autograd.Function):
    """(means, covs, venc, *params) -> (raw_rgb [M,3], raw_den [M,C], n_raw [M,3] | None)."""

    @staticmethod
    @_amp_fwd
    def forward(ctx, means, covs, venc, cfg, anchor, *params):
        names = cfg["names"]
        ctx.n_tape_params = len(params)
        if cfg.get("direct_params") is not None:     # FlatAdam-managed parameters (see radiance_field)
            params = tuple(cfg["direct_params"])
        P = dict(zip(names, params))
        be = make_backend(cfg["precision"], P)
        depth, skip = cfg["depth"], cfg["skip"]
        width, xyz = cfg["width"], cfg["xyz_dim"]
        S = cfg["samples_per_ray"]
        M = means.numel() // 3
        dev = means.device
        dt = be.dtype
        means2, covs2 = means.reshape(M, 3), covs.reshape(M, 3)
        f32 = torch.float32

        fused = be.name == "tc" and _use_fused(cfg, P)
        if fused:
            # (inside forward() grad mode is always off: radiance_field() records the caller's mode in cfg)
            need_bwd = cfg.get("grad_enabled", True) and any(ctx.needs_input_grad)
            C = P["density_layer.weight"].shape[0]
            vmod = cfg.get("venc_mod", 0)
            wv = P["view_layers.0.0.weight"]
            # per-ray view-direction term of the view layer (env rays: one row per env direction, indexed modulo D)
            vb = torch.empty(venc.shape[0], wv.shape[0], device=dev, dtype=f32)
            _F32Backend(P).linear(venc, wv[:, width:], vb, bias=P["view_layers.0.0.bias"])
            pack = fused_pack(names, params)
            g_enc = torch.empty(M, xyz, device=dev, dtype=f32) if cfg["with_normals"] else None
            # In-kernel IPE (encoder warps of the fused kernel, bit-identical to the two-kernel path) is OPT-IN: measured
            # on B200 it is slower - 1.61 vs 1.14 ms per 1.2 M samples, 208 vs 184 ms per 1024x512 panorama - because
            # the inference kernel is bound by its epilogue warps' issue slots and the 2 encoder warps (96 sin/cos/exp
            # features per sample with the reference-exact argument reduction) take slots from them; with the
            # hand-shakes alone (no arithmetic) the kernel is 5 % FASTER than with HBM-resident encodings
            # (profiles/r02_in_kernel_ipe_experiment.md).
            in_kernel_ipe = (not need_bwd and xyz == 96 and 0 <= cfg["min_deg"] and cfg["max_deg"] <= 31
                             and bool(os.environ.get("PNB_FUSED_IPE")))
            if in_kernel_ipe:
                enc = acts = masks = None
                raw_den, raw_rgb = fused_forward_ipe(means2, covs2, cfg["min_deg"], vb, vmod, S, C, pack, g_enc)
            else:
                enc = torch.empty(M, xyz, device=dev, dtype=dt)
                ops.ipe_into(means2, covs2, cfg["min_deg"], cfg["max_deg"], enc)
                planes = int(_lib.lib().pnb_mlp_fused_act_planes())
                acts = torch.empty(planes, M, width, device=dev, dtype=dt) if need_bwd else None
                masks = fused_masks(M, dev, True) if need_bwd else None
                raw_den, raw_rgb = fused_forward(enc, vb, S, C, pack, acts, g_enc, masks, need_bwd, vb_mod=vmod)
            n_raw, jac = None, None
            g_keep = None
            if cfg["with_normals"]:
                v = ops.ipe_vjp(means2, covs2, cfg["min_deg"], cfg["max_deg"], g_enc)
                # stop_resample_grad=False: the Gaussians are on the tape and the normals depend on them explicitly
                # through the encoding's Jacobian; d sigma / d enc is then needed again in the backward
                g_keep = g_enc if (need_bwd and (means.requires_grad or covs.requires_grad)) else None
                del g_enc
                n_raw = torch.empty(M, 3, device=dev, dtype=f32)
                with torch.cuda.device(dev):
                    check(_lib.lib().pnb_density_grad_fwd(M, C, _p(raw_den), float(cfg["density_bias"]), _p(v),
                                                          _p(n_raw), _stream()), "density_grad_fwd")
                if need_bwd:
                    jac = ([acts[10 + i] for i in range(depth)], v)
            ctx.cfg = cfg
            ctx.n_params = len(params)
            ctx.need_means = means.requires_grad
            ctx.need_covs = covs.requires_grad
            ctx.g_keep = g_keep
            ctx.means_shape = means.shape
            if need_bwd:
                if vmod:                         # the backward's view-direction weight gradient is per ray
                    venc = venc[None].expand(M // (S * vmod), vmod, venc.shape[1]).reshape(-1, venc.shape[1]).contiguous()
                ctx.bufs = dict(means=means2, covs=covs2, venc=venc, enc=enc, hs=[acts[i] for i in range(depth)],
                                bott=acts[8], hv=acts[9][:, :wv.shape[0]], vb_rows=venc.shape[0], raw_den=raw_den,
                                jac=jac, masks=masks, pack=pack, acts=acts,
                                fused=not os.environ.get("PNB_NO_FUSED_BWD"))
            else:
                ctx.bufs = None
            ctx.params = params
            if n_raw is None:
                return raw_rgb, raw_den, None
            return raw_rgb, raw_den, n_raw

        if cfg.get("venc_mod", 0):               # layered / parity paths take one view-direction row per ray
            vmod = cfg["venc_mod"]
            venc = venc[None].expand(M // (S * vmod), vmod, venc.shape[1]).reshape(-1, venc.shape[1]).contiguous()
        # cat = [h_skip | enc] (the input of the layer after the skip connection); enc lives only there
        cat = torch.empty(M, width + xyz, device=dev, dtype=dt)
        enc = cat[:, width:]
        ops.ipe_into(means2, covs2, cfg["min_deg"], cfg["max_deg"], enc)
        hs = []
        x = enc
        for i in range(depth):
            to_cat = (i % skip == 0 and i > 0)
            out = cat[:, :width] if to_cat else torch.empty(M, width, device=dev, dtype=dt)
            be.linear(x, be.w(f"layers.{i}.0.weight"), out, bias=P[f"layers.{i}.0.bias"], relu=True)
            hs.append(out)
            x = cat if to_cat else out
        trunk_out = x
        C = P["density_layer.weight"].shape[0]
        raw_den = torch.empty(M, C, device=dev, dtype=f32)
        be.linear(trunk_out, be.w("density_layer.weight"), raw_den, bias=P["density_layer.bias"])
        bott = torch.empty(M, width, device=dev, dtype=dt)
        be.linear(trunk_out, be.w("extra_layer.weight"), bott, bias=P["extra_layer.bias"])
        # view layer: the 27 view-direction inputs are constant along a ray -> per-ray addend (fp32, tiny)
        wv = P["view_layers.0.0.weight"]
        wc_ = wv.shape[0]
        vb = torch.empty(venc.shape[0], wc_, device=dev, dtype=f32)
        _F32Backend(P).linear(venc, wv[:, width:], vb, bias=P["view_layers.0.0.bias"])
        hv = torch.empty(M, wc_, device=dev, dtype=dt)
        be.linear(bott, be.w("view_layers.0.0.weight", 0, width), hv, relu=True, row_bias=vb, group=S)
        raw_rgb = torch.empty(M, P["color_layer.weight"].shape[0], device=dev, dtype=f32)
        be.linear(hv, be.w("color_layer.weight"), raw_rgb, bias=P["color_layer.bias"])

        n_raw = None
        jac = None
        if cfg["with_normals"]:
            # Jacobian sweep: a_i = relu'(h_i) * (a_{i+1} W_{i+1}), seeded with the sigma row of the density head.
            # `jac_precision='fp32'` runs this sweep (and its adjoint) on the fp32 FFMA path: the normals are the
            # numerically delicate part of the model ("precision of the second-order path").
            jb = _F32Backend(P) if cfg.get("jac_precision") == "fp32" and be.name != "f32" else be
            jdt = jb.dtype
            hs_j = [h.float() for h in hs] if jb is not be else hs
            a = [None] * depth
            a[depth - 1] = torch.empty(M, width, device=dev, dtype=jdt)
            _mask_scale(hs_j[depth - 1], P["density_layer.weight"][0].contiguous(), a[depth - 1])
            g_enc = torch.empty(M, xyz, device=dev, dtype=f32)
            skip_layers = [i for i in range(1, depth) if (i - 1) % skip == 0 and i > 1]
            first = True
            for i in range(depth - 1, 0, -1):
                a[i - 1] = torch.empty(M, width, device=dev, dtype=jdt)
                if i in skip_layers:
                    jb.dgrad(a[i], jb.w(f"layers.{i}.0.weight", 0, width), a[i - 1], mask=hs_j[i - 1])
                    jb.dgrad(a[i], jb.w(f"layers.{i}.0.weight", width, width + xyz), g_enc, accum=not first)
                    first = False
                else:
                    jb.dgrad(a[i], jb.w(f"layers.{i}.0.weight"), a[i - 1], mask=hs_j[i - 1])
            jb.dgrad(a[0], jb.w("layers.0.0.weight"), g_enc, accum=not first)
            del hs_j
            v = ops.ipe_vjp(means2, covs2, cfg["min_deg"], cfg["max_deg"], g_enc)
            g_keep = g_enc if (means.requires_grad or covs.requires_grad) else None
            del g_enc
            n_raw = torch.empty(M, 3, device=dev, dtype=f32)
            with torch.cuda.device(dev):
                check(_lib.lib().pnb_density_grad_fwd(M, C, _p(raw_den), float(cfg["density_bias"]), _p(v), _p(n_raw),
                                                      _stream()), "density_grad_fwd")
            jac = (a, v)

        ctx.cfg = cfg
        ctx.n_params = len(params)
        ctx.need_means = means.requires_grad
        ctx.need_covs = covs.requires_grad
        ctx.g_keep = g_keep if cfg["with_normals"] else None
        ctx.means_shape = means.shape
        # raw buffers are kept on ctx (not save_for_backward): they are private to this Function and never
        # modified in place afterwards
        ctx.bufs = dict(means=means2, covs=covs2, venc=venc, enc=enc, hs=hs, bott=bott, hv=hv, vb_rows=venc.shape[0],
                        raw_den=raw_den, jac=jac)
        ctx.params = params
        if n_raw is None:
            return raw_rgb, raw_den, None
        return raw_rgb, raw_den, n_raw

    @staticmethod
    def _backward_fused(ctx, d_raw_rgb, d_raw_den, d_n_raw):
        """Backward of the fused forward: the dgrad chain and the adjoint Jacobian sweep each run as one fused kernel
        (csrc/mlp_fused.cu, programs P_BWD / P_JADJ) that reads the ReLU sign bit-planes and writes the dz / q planes;
        the weight gradients are then plain reductions over the samples (pnb_wgrad_tc)."""
        cfg, B = ctx.cfg, ctx.bufs
        names = cfg["names"]
        P = dict(zip(names, ctx.params))
        be = _TCBackend(P)
        f32be = _F32Backend(P)
        depth, width, xyz = cfg["depth"], cfg["width"], cfg["xyz_dim"]
        S = cfg["samples_per_ray"]
        enc, hs, bott, hv, raw_den = B["enc"], B["hs"], B["bott"], B["hv"], B["raw_den"]
        means, covs, venc, masks, pack, acts = B["means"], B["covs"], B["venc"], B["masks"], B["pack"], B["acts"]
        M = means.shape[0]
        dev, f32 = means.device, torch.float32
        wb = WgradBatch(M, dev)
        C = raw_den.shape[1]
        # Gradient destinations.  Parameters managed by FlatAdam carry `_pnb_direct_grad`: their .grad tensors are views
        # of the optimiser's flat buffer and every kernel below accumulates (+=), so the gradients go straight there
        # and autograd receives None for them - no zero-fill of a temporary, no 24 add kernels per level.  Anything
        # else (torch.autograd.grad, plain optimisers) gets freshly computed tensors as usual.
        direct = all(getattr(p, "_pnb_direct_grad", False) and p.grad is not None and p.grad.is_contiguous()
                     and p.grad.dtype == f32 for p in ctx.params)
        G = {}
        if direct:
            for nme, p in zip(names, ctx.params):
                G[nme] = p.grad
        else:
            sizes = [p.numel() for p in ctx.params]
            flat = torch.zeros(sum(sizes), device=dev, dtype=f32)
            off = 0
            for nme, p, sz in zip(names, ctx.params, sizes):
                G[nme] = flat[off:off + sz].view_as(p)
                off += sz
        W = lambda i: f"layers.{i}.0.weight"
        d_raw_den = torch.zeros(M, C, device=dev, dtype=f32) if d_raw_den is None else d_raw_den.contiguous().clone()
        d_raw_rgb = torch.zeros(M, 3, device=dev, dtype=f32) if d_raw_rgb is None else d_raw_rgb.contiguous()

        # ---- adjoint of the Jacobian sweep (second-order terms of the normals) -------------------------------
        d_v = None
        if B["jac"] is not None and d_n_raw is not None:
            a, v = B["jac"]
            d_n_raw = d_n_raw.contiguous()
            d_raw0 = torch.empty(M, device=dev, dtype=f32)
            d_v = torch.empty(M, 3, device=dev, dtype=f32)
            with torch.cuda.device(dev):
                check(_lib.lib().pnb_density_grad_bwd(M, C, _p(raw_den), float(cfg["density_bias"]), _p(v), _p(d_n_raw),
                                                      _p(d_raw0), _p(d_v), _stream()), "density_grad_bwd")
            d_raw_den[:, 0] += d_raw0
            u = torch.empty(M, xyz, device=dev, dtype=torch.bfloat16)
            ops.ipe_jvp_into(means, covs, cfg["min_deg"], cfg["max_deg"], d_v, u)
            q = fused_jadj(u, pack, masks)
            # dW_i += a_i^T q_{i-1}  (q_{-1} = u; layer 5 also sees u through the skip connection)
            wb.add(acts, 10, 256, u, 0, xyz, G[W(0)])
            for i in range(1, depth):
                if i == 5:
                    wb.add(acts, 10 + i, 256, q, i - 1, width, G[W(i)][:, :width])
                    wb.add(acts, 10 + i, 256, u, 0, xyz, G[W(i)][:, width:])
                else:
                    wb.add(acts, 10 + i, 256, q, i - 1, width, G[W(i)])
            # d w_sigma += column sums of q_7 (a_7 = relu'(h_7) * w_sigma): rides on a minimal dummy product
            scratch = torch.zeros(256, 16, device=dev, dtype=f32)
            wb.add(q, depth - 1, 256, u, 0, 16, scratch, G["density_layer.weight"][0])

        # ---- dgrad chain ------------------------------------------------------------------------------------
        need_enc = ctx.need_means or ctx.need_covs
        d_enc = torch.empty(M, xyz, device=dev, dtype=f32) if need_enc else None
        dz = fused_backward(M, C, pack, d_raw_rgb, d_raw_den, masks, d_enc)
        dzv = dz[0][:, :hv.shape[1]]
        # ---- weight / bias gradients: one batched launch -----------------------------------------------------
        wc = hv.shape[1]
        pad_rgb = _pad_head_grad(d_raw_rgb, G["color_layer.bias"])
        pad_den = _pad_head_grad(d_raw_den, G["density_layer.bias"])
        tmp_c = torch.zeros(wc, 64, device=dev, dtype=f32)
        tmp_d = torch.zeros(width, 64, device=dev, dtype=f32)
        wb.add(acts, 9, wc, pad_rgb, 0, 64, tmp_c)                      # (hv^T d_rgb): colour head, transposed
        wb.add(dz, 0, wc, acts, 8, width, G["view_layers.0.0.weight"][:, :width], G["view_layers.0.0.bias"])
        wb.add(dz, 1, 256, acts, 7, width, G["extra_layer.weight"], G["extra_layer.bias"])
        wb.add(acts, 7, 256, pad_den, 0, 64, tmp_d)                     # (h7^T d_den): density head, transposed
        for i in range(depth - 1, 0, -1):
            if i == 5:
                wb.add(dz, 9 - i, 256, acts, i - 1, width, G[W(i)][:, :width], G[f"layers.{i}.0.bias"])
                wb.add(dz, 9 - i, 256, enc, 0, xyz, G[W(i)][:, width:])
            else:
                wb.add(dz, 9 - i, 256, acts, i - 1, width, G[W(i)], G[f"layers.{i}.0.bias"])
        wb.add(dz, 9, 256, enc, 0, xyz, G[W(0)], G["layers.0.0.bias"])
        wb.launch()
        G["color_layer.weight"].add_(tmp_c[:, :3].t())
        G["density_layer.weight"].add_(tmp_d[:, :C].t())
        dvb = _group_sum(dzv, S)
        f32be.wgrad(dvb, venc, G["view_layers.0.0.weight"][:, width:])
        d_means, d_covs = _gaussian_grads(ctx, means, covs, cfg, d_enc if need_enc else None, d_v)
        ctx.bufs = None
        if direct:
            return (d_means, d_covs, None, None, None) + (None,) * ctx.n_tape_params
        if ctx.n_tape_params == 0:                   # off-tape parameters lost their flat views mid-step: fold
            for nme, p_ in zip(names, ctx.params):
                p_.grad = G[nme] if p_.grad is None else p_.grad + G[nme]
            return (d_means, d_covs, None, None, None)
        return (d_means, d_covs, None, None, None) + tuple(G[n] for n in names)

    @staticmethod
    @_amp_bwd
    def backward(ctx, d_raw_rgb, d_raw_den, d_n_raw):
        cfg, B = ctx.cfg, ctx.bufs
        if B.get("fused"):
            return _Field._backward_fused(ctx, d_raw_rgb, d_raw_den, d_n_raw)
        names = cfg["names"]
        P = dict(zip(names, ctx.params))
        be = make_backend(cfg["precision"], P)
        f32be = _F32Backend(P)
        depth, skip, width, xyz = cfg["depth"], cfg["skip"], cfg["width"], cfg["xyz_dim"]
        S = cfg["samples_per_ray"]
        enc, hs, bott, hv, raw_den = B["enc"], B["hs"], B["bott"], B["hv"], B["raw_den"]
        means, covs, venc = B["means"], B["covs"], B["venc"]
        M = means.shape[0]
        dev, dt, f32 = means.device, be.dtype, torch.float32
        C = raw_den.shape[1]
        skip_layers = [i for i in range(1, depth) if (i - 1) % skip == 0 and i > 1]

        # one flat, zero-initialised gradient buffer; per-parameter views are what autograd receives
        sizes = [p.numel() for p in ctx.params]
        flat = torch.zeros(sum(sizes), device=dev, dtype=f32)
        G, off = {}, 0
        for nme, p, sz in zip(names, ctx.params, sizes):
            G[nme] = flat[off:off + sz].view_as(p)
            off += sz

        d_raw_den = torch.zeros(M, C, device=dev, dtype=f32) if d_raw_den is None else d_raw_den.contiguous().clone()
        d_raw_rgb = torch.zeros(M, 3, device=dev, dtype=f32) if d_raw_rgb is None else d_raw_rgb.contiguous()

        # ---- adjoint of the Jacobian sweep (second-order terms of the normals) -------------------------------
        d_v = None
        if B["jac"] is not None and d_n_raw is not None:
            a, v = B["jac"]
            d_n_raw = d_n_raw.contiguous()
            d_raw0 = torch.empty(M, device=dev, dtype=f32)
            d_v = torch.empty(M, 3, device=dev, dtype=f32)
            with torch.cuda.device(dev):
                check(_lib.lib().pnb_density_grad_bwd(M, C, _p(raw_den), float(cfg["density_bias"]), _p(v), _p(d_n_raw),
                                                      _p(d_raw0), _p(d_v), _stream()), "density_grad_bwd")
            d_raw_den[:, 0] += d_raw0
            # u = d L / d g_enc = J_ipe d_v ; forward-mode sweep q_i = relu'(h_i) * (q_{i-1} W_i^T)
            jb = _F32Backend(P) if cfg.get("jac_precision") == "fp32" and be.name != "f32" else be
            jdt = jb.dtype
            hs_j = [h.float() for h in hs] if jb is not be else hs
            ucat = torch.empty(M, width + xyz, device=dev, dtype=jdt)
            u = ucat[:, width:]
            ops.ipe_jvp_into(means, covs, cfg["min_deg"], cfg["max_deg"], d_v, u)
            q_prev = u
            for i in range(depth):
                wname = f"layers.{i}.0.weight"
                # dW_i += a_i^T q_{i-1}
                if i in skip_layers:
                    jb.wgrad(a[i], ucat[:, :width], G[wname][:, :width])
                    jb.wgrad(a[i], u, G[wname][:, width:])
                else:
                    jb.wgrad(a[i], q_prev, G[wname])
                to_cat = (i % skip == 0 and i > 0)
                q = ucat[:, :width] if to_cat else torch.empty(M, width, device=dev, dtype=jdt)
                jb.linear(q_prev, jb.w(wname), q, mask=hs_j[i])
                a[i] = None
                q_prev = ucat if to_cat else q
            # d wd[0,:] += colsum(q_last)   (a_last = relu'(h_last) * wd[0])
            last = q_prev if q_prev.shape[1] == width else q_prev[:, :width]
            _colsum(last, G["density_layer.weight"][0])
            del hs_j
            del ucat, q_prev

        # ---- heads ------------------------------------------------------------------------------------------
        if (depth - 1) % skip == 0 and depth - 1 > 0:
            raise NotImplementedError("a skip connection feeding the heads is not supported")
        trunk_out = hs[depth - 1]
        # colour head
        _colsum(d_raw_rgb, G["color_layer.bias"])
        dzv = torch.empty(M, hv.shape[1], device=dev, dtype=dt)
        pad = be.dgrad_small(d_raw_rgb, be.w("color_layer.weight"), dzv, mask=hv)
        if isinstance(be, _TCBackend):
            be.wgrad_small(d_raw_rgb, hv, G["color_layer.weight"], pad=pad)
        else:
            be.wgrad(d_raw_rgb, hv, G["color_layer.weight"])
        del pad
        # view layer: bottleneck part on the GEMM path, per-ray view-direction part in fp32
        be.wgrad(dzv, bott, G["view_layers.0.0.weight"][:, :width])
        dvb = _group_sum(dzv, S)
        _colsum(dvb, G["view_layers.0.0.bias"])
        f32be.wgrad(dvb, venc, G["view_layers.0.0.weight"][:, width:])
        d_bott = torch.empty(M, width, device=dev, dtype=dt)
        be.dgrad(dzv, be.w("view_layers.0.0.weight", 0, width), d_bott, colsum_out=G["extra_layer.bias"])
        del dzv
        # extra + density heads -> d trunk_out (masked by the last ReLU)
        be.wgrad(d_bott, trunk_out if trunk_out.shape[1] == width else trunk_out, G["extra_layer.weight"])
        _colsum(d_raw_den, G["density_layer.bias"])
        dz = torch.empty(M, width, device=dev, dtype=dt)
        h_last = hs[depth - 1]
        be.dgrad(d_bott, be.w("extra_layer.weight"), dz)
        pad = be.dgrad_small(d_raw_den, be.w("density_layer.weight"), dz, mask=h_last, accum=True,
                             colsum_out=G[f"layers.{depth - 1}.0.bias"])
        if isinstance(be, _TCBackend):
            be.wgrad_small(d_raw_den, h_last, G["density_layer.weight"], pad=pad)
        else:
            be.wgrad(d_raw_den, h_last, G["density_layer.weight"])
        del pad, d_bott

        # ---- trunk -------------------------------------------------------------------------------------------
        need_enc = ctx.need_means or ctx.need_covs
        d_enc = torch.zeros(M, xyz, device=dev, dtype=f32) if need_enc else None
        for i in range(depth - 1, -1, -1):
            wname = f"layers.{i}.0.weight"
            prev_bias = G[f"layers.{i - 1}.0.bias"] if i > 0 else None   # filled by this layer's dgrad epilogue
            if i == 0:
                be.wgrad(dz, enc, G[wname])
                if need_enc:
                    be.dgrad(dz, be.w(wname), d_enc, accum=True)
                break
            x_prev = hs[i - 1]
            if i in skip_layers:
                be.wgrad(dz, hs[i - 1], G[wname][:, :width])
                be.wgrad(dz, enc, G[wname][:, width:])
                if need_enc:
                    be.dgrad(dz, be.w(wname, width, width + xyz), d_enc, accum=True)
                nxt = torch.empty(M, width, device=dev, dtype=dt)
                be.dgrad(dz, be.w(wname, 0, width), nxt, mask=x_prev, colsum_out=prev_bias)
            else:
                be.wgrad(dz, x_prev, G[wname])
                nxt = torch.empty(M, width, device=dev, dtype=dt)
                be.dgrad(dz, be.w(wname), nxt, mask=x_prev, colsum_out=prev_bias)
            dz = nxt
        d_means, d_covs = _gaussian_grads(ctx, means, covs, cfg, d_enc if need_enc else None, d_v)
        ctx.bufs = None
        return (d_means, d_covs, None, None, None) + tuple(G[n] for n in names)


def radiance_field(means, covs, venc, params: Dict[str, torch.Tensor], *, precision: str, samples_per_ray: int,
                   min_deg: int, max_deg: int, density_bias: float, skip: int, with_normals: bool,
                   jac_precision: Optional[str] = None, venc_mod: int = 0):
    """Evaluate the MLP on [R,S,3] Gaussians. Returns raw_rgb [R,S,3], raw_den [R,S,C], n_raw [R,S,3] | None.
    `venc` is one view encoding per ray, or - with `venc_mod` = D - one per env direction, ray r using row r % D
    (models/pano_mip_nerf.py:337-341 broadcasts the D env directions over the surface points)."""
    names = list(params.keys())
    depth = len([n for n in names if n.startswith("layers.") and n.endswith(".weight")])
    w0 = params["layers.0.0.weight"]
    cfg = dict(names=names, precision=precision, depth=depth, skip=skip, width=w0.shape[0], xyz_dim=w0.shape[1],
               samples_per_ray=samples_per_ray, min_deg=min_deg, max_deg=max_deg, density_bias=density_bias,
               with_normals=with_normals, jac_precision=jac_precision, grad_enabled=torch.is_grad_enabled(),
               venc_mod=int(venc_mod))
    if venc_mod and (venc.shape[0] != venc_mod or means.shape[0] % venc_mod):
        raise RuntimeError("venc_mod: expected one view encoding per env direction and rays = points x directions")
    if w0.shape[1] != 6 * (max_deg - min_deg):
        raise RuntimeError("IPE width does not match the first layer")
    R = means.shape[0]
    if means.numel() == 0:                       # an empty batch: empty outputs of the right widths, nothing to launch
        C = params["density_layer.weight"].shape[0]
        tie = sum(p.sum() for p in params.values()) * 0.0      # keeps the outputs on the tape (zero gradients)
        e = lambda ch: torch.zeros(R, samples_per_ray, ch, device=means.device, dtype=torch.float32) + tie
        return e(params["color_layer.weight"].shape[0]), e(C), (e(3) if with_normals else None)
    plist = [params[n] for n in names]
    anchor = None
    if (cfg["grad_enabled"] and _use_fused(cfg, params)
            and all(getattr(p, "_pnb_direct_grad", False) and p.grad is not None for p in plist)):
        # FlatAdam-managed parameters: the fused backward accumulates straight into their .grad views, so autograd
        # does not need edges to the 24 leaves at all.  They are handed over outside the tape and a fresh, empty
        # `anchor` leaf keeps the Function differentiable.  Besides saving 24 no-op AccumulateGrad calls per level
        # this keeps CUDA-graph capture independent of earlier eager steps: a parameter's cached AccumulateGrad node
        # is pinned to the stream of the step that created it (e.g. the default stream) as long as any old loss
        # tensor keeps that tape alive, and the engine would then sync the capture stream with it
        # ("dependency created on uncaptured work in another stream").
        cfg["direct_params"] = plist
        anchor = torch.empty(0, device=means.device, dtype=torch.float32, requires_grad=True)
        plist = []
    raw_rgb, raw_den, n_raw = _Field.apply(means, covs, venc, cfg, anchor, *plist)
    raw_rgb = raw_rgb.view(R, samples_per_ray, raw_rgb.shape[-1])   # (explicit widths: R may be 0)
    raw_den = raw_den.view(R, samples_per_ray, raw_den.shape[-1])
    if n_raw is not None:
        n_raw = n_raw.view(R, samples_per_ray, 3)
    return raw_rgb, raw_den, n_raw
