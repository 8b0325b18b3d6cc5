"""HBM bandwidth probes (CUDA events): pure write (fill), pure read (sum), copy - denominators for the fused kernels
whose traffic is mostly one-directional (training forward: writes; weight gradients: reads)."""
import json
import torch

dev = torch.device("cuda", 0)
n = 1 << 30                                         # 4 GiB of fp32
a = torch.empty(n, device=dev, dtype=torch.float32)
b = torch.empty(n, device=dev, dtype=torch.float32)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


out = {}
ms = timeit(lambda: a.fill_(1.0))
out["fill_write_GBps"] = 4 * n / ms / 1e6
ms = timeit(lambda: a.zero_())
out["memset_write_GBps"] = 4 * n / ms / 1e6
ms = timeit(lambda: a.sum())
out["sum_read_GBps"] = 4 * n / ms / 1e6
ms = timeit(lambda: b.copy_(a))
out["copy_read_plus_write_GBps"] = 8 * n / ms / 1e6
a16 = a.view(torch.bfloat16)
ms = timeit(lambda: torch.cudaMemsetAsync if False else a16.fill_(0.5))
out["fill_bf16_write_GBps"] = 4 * n / ms / 1e6
print(json.dumps(out))
