"""Micro-benchmark of the tcgen05 GEMM entry points (CUDA events, L2-exceeding operands)."""
import ctypes, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from panonerf_b200 import _lib, field, ops
DEV = "cuda"
M = int(os.environ.get("M", 524288))
iters = int(os.environ.get("ITERS", 10))

def timeit(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

P = {"layers.0.0.weight": torch.randn(256, 96, device=DEV) / 10, "view_layers.0.0.weight": torch.randn(128, 283, device=DEV) / 16,
     "w": torch.randn(256, 256, device=DEV) / 16, "w5": torch.randn(256, 352, device=DEV) / 18}
be = field._TCBackend(P)
a = torch.randn(M, 256, device=DEV).to(torch.bfloat16)
cat = torch.randn(M, 352, device=DEV).to(torch.bfloat16)
out = torch.empty(M, 256, device=DEV, dtype=torch.bfloat16)
out2 = torch.empty(M, 256, device=DEV, dtype=torch.bfloat16)
bias = torch.randn(256, device=DEV)
dw = torch.zeros(256, 256, device=DEV)
res = {}
res["linear_256x256_relu"] = timeit(lambda: be.linear(a, be.w("w"), out, bias=bias, relu=True))
res["linear_256x256_nobias"] = timeit(lambda: be.linear(a, be.w("w"), out))
res["dgrad_256x256_mask"] = timeit(lambda: be.dgrad(a, be.w("w"), out2, mask=out))
res["linear_352"] = timeit(lambda: be.linear(cat, be.w("w5"), out, bias=bias, relu=True))
res["linear_96"] = timeit(lambda: be.linear(cat[:, 256:], be.w("layers.0.0.weight"), out, bias=bias, relu=True))
res["wgrad_256x256"] = timeit(lambda: be.wgrad(a, out, dw))
gb = lambda nbytes, ms: nbytes / (ms / 1e3) / 1e9
print(json.dumps({k: dict(ms=v) for k, v in res.items()}))
print("linear relu GB/s", gb(M * 256 * 2 * 2, res["linear_256x256_relu"]), "TFLOP/s", 2 * M * 256 * 256 / (res["linear_256x256_relu"] / 1e3) / 1e12)
print("dgrad mask GB/s", gb(M * 256 * 2 * 3, res["dgrad_256x256_mask"]))
print("wgrad GB/s", gb(M * 256 * 2 * 2, res["wgrad_256x256"]))
