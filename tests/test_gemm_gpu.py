"""GPU parity of the MLP GEMM entry points: fp32 FFMA path (1e-5) and the tcgen05 bf16 path (bf16 inputs are
exact in fp32, so against an fp32 product of the same bf16-rounded operands the error is accumulation order only)."""
import ctypes

import pytest
import torch

from util import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _be(kind, params=None):
    from panonerf_b200 import field
    return field._F32Backend(params or {}) if kind == "f32" else field._TCBackend.__new__(field._TCBackend)


@pytest.mark.parametrize("m,n,k", [(1000, 256, 96), (4096, 256, 256), (777, 256, 352), (300, 5, 256), (64, 3, 128),
                                   (129, 128, 283)])
def test_gemm_f32_modes(m, n, k):
    from panonerf_b200 import field
    gen = torch.Generator().manual_seed(m + n + k)
    a = torch.randn(m, k, generator=gen).to(DEV)
    w = (torch.randn(n, k, generator=gen) / k ** 0.5).to(DEV)
    b = torch.randn(n, generator=gen).to(DEV)
    be = field._F32Backend({})
    out = torch.empty(m, n, device=DEV)
    be.linear(a, w, out, bias=b, relu=True)
    ref = torch.relu(a.double() @ w.double().t() + b.double())
    assert_close(out, ref.float(), 1e-5, "linear", floor=float(ref.abs().mean()))
    dz = torch.randn(m, n, generator=gen).to(DEV)
    dx = torch.empty(m, k, device=DEV)
    be.dgrad(dz, w, dx, mask=a)
    ref = (dz.double() @ w.double()) * (a > 0)
    assert_close(dx, ref.float(), 1e-5, "dgrad", floor=float(ref.abs().mean()))
    dw = torch.zeros(n, k, device=DEV)
    be.wgrad(dz, a, dw)
    ref = dz.double().t() @ a.double()
    assert_close(dw, ref.float(), 2e-5, "wgrad", floor=float(ref.abs().mean()))
    # strided views (column slices of wider buffers)
    wide = torch.randn(m, k + 40, generator=gen).to(DEV)
    out2 = torch.zeros(m, n + 8, device=DEV)
    be.linear(wide[:, 8:8 + k], w, out2[:, 4:4 + n])
    ref = wide[:, 8:8 + k].double() @ w.double().t()
    assert_close(out2[:, 4:4 + n], ref.float(), 1e-5, "strided", floor=float(ref.abs().mean()))
    assert float(out2[:, :4].abs().max()) == 0 and float(out2[:, 4 + n:].abs().max()) == 0


def _tc():
    from panonerf_b200 import _lib
    if not _lib.lib().pnb_tc_available():
        pytest.skip("not an sm_100 device")


def _linear_tc(a, w_fwd, out, n, k, bias=None, relu=False, mask=None, row_bias=None, group=0, accum=False, colsum=None):
    from panonerf_b200 import _lib, ops
    from panonerf_b200._lib import EPI_ACCUM, EPI_BIAS, EPI_MASK, EPI_RELU
    flags = (EPI_BIAS if bias is not None else 0) | (EPI_RELU if relu else 0) | (EPI_MASK if mask is not None else 0) | \
            (EPI_ACCUM if accum else 0)
    p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
    _lib.check(_lib.lib().pnb_linear_tc(a.shape[0], n, k, p(a), a.stride(0), p(w_fwd), w_fwd.stride(0), p(out),
                                        out.stride(0), ops.dt_code(out.dtype), p(bias), p(row_bias), group, p(mask),
                                        mask.stride(0) if mask is not None else 0, flags, p(colsum),
                                        ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "linear_tc")


@pytest.mark.parametrize("m,n,k", [(128, 256, 64), (128, 256, 256), (1000, 256, 96), (70000, 256, 256),
                                   (4096, 256, 352), (333, 128, 256), (4096, 96, 256), (500, 5, 256), (500, 3, 128),
                                   (2048, 256, 16), (2048, 64, 128), (100000, 128, 128)])
def test_linear_tc(m, n, k):
    _tc()
    gen = torch.Generator().manual_seed(m * 7 + n * 3 + k)
    a = torch.randn(m, k, generator=gen).to(DEV).to(torch.bfloat16)
    kp = (k + 63) // 64 * 64
    w = torch.zeros(n, kp, device=DEV, dtype=torch.bfloat16)
    w[:, :k] = (torch.randn(n, k, generator=gen) / k ** 0.5).to(DEV)
    b = torch.randn(n, generator=gen).to(DEV)
    ref = a.double() @ w[:, :k].double().t() + b.double()
    out = torch.full((m, n), 7.0, device=DEV, dtype=torch.float32)
    _linear_tc(a, w, out, n, k, bias=b)
    torch.cuda.synchronize()
    assert_close(out, ref.float(), 2e-5, "fp32 out", floor=float(ref.abs().mean()))
    outb = torch.empty(m, n, device=DEV, dtype=torch.bfloat16)
    _linear_tc(a, w, outb, n, k, bias=b, relu=True)
    assert_close(outb.float(), torch.relu(ref).float(), 5e-3, "bf16 relu out", floor=float(ref.abs().mean()))


def test_linear_tc_epilogues_and_strides():
    _tc()
    gen = torch.Generator().manual_seed(11)
    m, S = 64 * 50, 64
    cat = torch.zeros(m, 352, device=DEV, dtype=torch.bfloat16)
    cat[:, 256:] = torch.randn(m, 96, generator=gen).to(DEV)
    w0 = (torch.randn(256, 96, generator=gen) / 10).to(DEV)
    w0p = torch.zeros(256, 128, device=DEV, dtype=torch.bfloat16)
    w0p[:, :96] = w0
    # layer-0 style: A is the tail of the cat buffer, output goes to the head of it
    _linear_tc(cat[:, 256:], w0p, cat[:, :256], 256, 96, relu=True)
    ref = torch.relu(cat[:, 256:].double() @ w0p[:, :96].double().t())
    assert_close(cat[:, :256].float(), ref.float(), 5e-3, "strided in/out", floor=float(ref.abs().mean()))
    # mask + per-ray row bias + accumulate
    a = torch.randn(m, 256, generator=gen).to(DEV).to(torch.bfloat16)
    w = (torch.randn(128, 256, generator=gen) / 16).to(DEV).to(torch.bfloat16)
    rb = torch.randn(m // S, 128, generator=gen).to(DEV)
    mask = torch.randn(m, 128, generator=gen).to(DEV).to(torch.bfloat16)
    out = torch.empty(m, 128, device=DEV, dtype=torch.bfloat16)
    _linear_tc(a, w, out, 128, 256, relu=True, row_bias=rb, group=S)
    ref = torch.relu(a.double() @ w.double().t() + rb.double().repeat_interleave(S, 0))
    assert_close(out.float(), ref.float(), 5e-3, "row_bias", floor=float(ref.abs().mean()))
    # fused column sums (bias gradient) on a ragged M, bf16 mask+accum epilogue
    mr = m - 37
    prevb = torch.randn(mr, 128, generator=gen).to(DEV).to(torch.bfloat16)
    outb = prevb.clone()
    cs = torch.ones(128, device=DEV)
    _linear_tc(a[:mr], w, outb, 128, 256, mask=mask[:mr], accum=True, colsum=cs)
    ref = (a[:mr].double() @ w.double().t() + prevb.double()) * (mask[:mr].double() > 0)
    assert_close(outb.float(), ref.float(), 5e-3, "bf16 mask+accum", floor=float(ref.abs().mean()))
    assert_close(cs, (outb.double().sum(0) + 1.0).float(), 1e-4, "fused colsum", floor=float(outb.double().sum(0).abs().mean()))
    prev = torch.randn(m, 128, generator=gen).to(DEV)
    out32 = prev.clone()
    _linear_tc(a, w, out32, 128, 256, mask=mask, accum=True)
    ref = (a.double() @ w.double().t() + prev.double()) * (mask.double() > 0)
    assert_close(out32, ref.float(), 2e-5, "mask+accum", floor=float(ref.abs().mean()))


@pytest.mark.parametrize("m,nw,kw", [(64, 256, 256), (4096, 256, 256), (100000, 256, 256), (5000, 128, 256),
                                     (3000, 256, 96), (3000, 256, 64), (777, 128, 64)])
def test_wgrad_tc(m, nw, kw):
    _tc()
    from panonerf_b200 import _lib
    gen = torch.Generator().manual_seed(m + nw + kw)
    dz = torch.randn(m, nw, generator=gen).to(DEV).to(torch.bfloat16)
    x = torch.randn(m, kw, generator=gen).to(DEV).to(torch.bfloat16)
    dw = torch.ones(nw, kw, device=DEV)
    ws = torch.empty(int(_lib.lib().pnb_wgrad_tc_workspace(256, 256)) // 4, device=DEV)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    _lib.check(_lib.lib().pnb_wgrad_tc(m, nw, kw, p(dz), dz.stride(0), p(x), x.stride(0), p(dw), dw.stride(0), p(ws),
                                       ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "wgrad_tc")
    ref = dz.double().t() @ x.double() + 1.0
    assert_close(dw, ref.float(), 2e-5, "wgrad_tc", floor=float(ref.abs().mean()))
